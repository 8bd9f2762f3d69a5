"""Persistence of tree planes behind ``save_to`` / ``load_from`` (plenvdb/lib/vdb/plenvdb.h:126-148, 211-240).

The reference writes OpenVDB ``.vdb`` files through libopenvdb, which cannot be linked here (no TBB /
Boost / Blosc).  Until the native ``.vdb`` codec lands, grids are stored in
a self-describing container at the path the caller gives: an ``.npz`` payload (written without the
extension being appended) holding the active-voxel coordinates and their values.  ``load_from`` applies
the reference's ``pruneGrid()`` semantics with tolerance 0 at leaf level: leaves whose stored voxels are
all inactive are dropped; the resolution is the active bounding box (``evalActiveVoxelDim``).
"""
import io

import numpy as np
import torch

from .tree import Topology

MAGIC = "plenvdb_b200.planes.v1"


def _active_coords(topo):
    """int32 [n_active,3] coords, plus (leaf, off) index arrays, of active voxels in leaf order."""
    masks = topo.h_leaf_mask[: topo.n_leaf]
    bits = np.unpackbits(masks.view(np.uint8).reshape(topo.n_leaf, 64), axis=1, bitorder="little").astype(bool)
    leaf, off = np.nonzero(bits)
    org = topo.h_leaf_origin[: topo.n_leaf][leaf]
    xyz = np.stack([org[:, 0] + (off >> 6), org[:, 1] + ((off >> 3) & 7), org[:, 2] + (off & 7)], 1).astype(np.int32)
    return xyz, leaf, off


def save_planes(path, topo, plane, reso, names):
    xyz, leaf, off = _active_coords(topo)
    vals = plane.detach().cpu().numpy()[leaf, off]      # [n_active, C]
    buf = io.BytesIO()
    np.savez_compressed(buf, magic=np.array(MAGIC), reso=np.asarray(reso, np.int32), xyz=xyz, values=vals.astype(np.float32),
                        names=np.array(names))
    with open(path, "wb") as f:
        f.write(buf.getvalue())


def _read(path):
    with open(path, "rb") as f:
        z = np.load(io.BytesIO(f.read()), allow_pickle=False)
    if str(z["magic"]) != MAGIC:
        raise ValueError("%s is not a plenvdb_b200 plane container" % path)
    return z


def load_planes(path, channels, device):
    """Returns (Topology, plane, reso). Topology = leaves containing at least one stored active voxel."""
    z = _read(path)
    xyz, vals = z["xyz"], z["values"]
    if vals.shape[1] != channels:
        raise ValueError("%s holds %d channels, expected %d" % (path, vals.shape[1], channels))
    if xyz.shape[0] == 0:
        reso = (1, 1, 1)
        active = np.zeros(reso, np.uint8)
    else:
        reso = tuple(int(v) + 1 for v in xyz.max(0))   # evalActiveVoxelDim of a grid anchored at 0
        active = np.zeros(reso, np.uint8)
        active[xyz[:, 0], xyz[:, 1], xyz[:, 2]] = 1
    topo = Topology.from_mask(active, device=device)
    plane = _scatter(topo, xyz, vals, channels, device)
    return topo, plane, reso


def load_plane_onto(path, topo, channels, reso, device):
    """Load values onto an existing topology (optimizer moments stay congruent with their parameter)."""
    z = _read(path)
    return _scatter(topo, z["xyz"], z["values"], channels, device)


def _scatter(topo, xyz, vals, channels, device):
    plane = np.zeros((max(topo.n_leaf, 1), 512, channels), np.float32)
    if xyz.shape[0]:
        # leaf lookup through the host tables (single root tile anchored at the origin)
        u = ((xyz[:, 0] >> 7) << 10) | ((xyz[:, 1] >> 7) << 5) | (xyz[:, 2] >> 7)
        low = topo.h_upper[u]
        l = (((xyz[:, 0] & 127) >> 3) << 8) | (((xyz[:, 1] & 127) >> 3) << 4) | ((xyz[:, 2] & 127) >> 3)
        ok = low >= 0
        leaf = np.full(xyz.shape[0], -1, np.int64)
        leaf[ok] = topo.h_lower[low[ok].astype(np.int64) * 4096 + l[ok]]
        ok &= leaf >= 0
        off = ((xyz[:, 0] & 7) << 6) | ((xyz[:, 1] & 7) << 3) | (xyz[:, 2] & 7)
        plane[leaf[ok], off[ok]] = vals[ok]
    return torch.from_numpy(plane).to(device)

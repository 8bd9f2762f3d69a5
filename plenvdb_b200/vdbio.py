"""Persistence of tree planes behind ``save_to`` / ``load_from`` (plenvdb/lib/vdb/plenvdb.h:126-148, 211-240).

The reference writes OpenVDB ``.vdb`` files through libopenvdb, which cannot be linked here (no TBB / Boost / Blosc;
SURVEY.md §8c).  ``save_planes`` writes real ``.vdb`` files with the native codec of ``openvdb_io``:
a DensityVDB becomes one FloatGrid "density", a ColorVDB four Vec3SGrids "color0".."color3" in one file, as the reference
does.  ``load_planes`` reads ``.vdb`` files (and still reads the round-1 interim ``.npz`` container) and applies the
reference's ``pruneGrid()`` semantics with tolerance 0 at leaf level: leaves whose stored voxels are all inactive are
dropped; the resolution is the active bounding box (``evalActiveVoxelDim``).
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


def save_planes(path, topo, plane, reso, names, container="vdb"):
    """names: ["density"] for a 1-channel plane, ["color0", ...] for 3 channels per name."""
    if container == "vdb":
        from . import openvdb_io
        p = plane.detach().cpu().numpy()
        per = p.shape[-1] // len(names)
        openvdb_io.write_vdb(path, topo, [(n, p[:, :, i * per:(i + 1) * per]) for i, n in enumerate(names)])
        return
    xyz, leaf, off = _active_coords(topo)
    vals = plane.detach().cpu().numpy()[leaf, off]      # [n_active, C]
    buf = io.BytesIO()
    np.savez_compressed(buf, magic=np.array(MAGIC), reso=np.asarray(reso, np.int32), xyz=xyz, values=vals.astype(np.float32),
                        names=np.array(names))
    with open(path, "wb") as f:
        f.write(buf.getvalue())


def _read(path):
    """-> dict(xyz int32 [n,3], values float32 [n,C]) of the active voxels stored in `path` (.vdb or the interim container)."""
    with open(path, "rb") as f:
        data = f.read()
    from . import openvdb_io
    if openvdb_io.is_vdb(data):
        grids = openvdb_io.decode_grids(data)
        if not grids:
            raise ValueError("%s holds no float / vec3s grid" % path)
        if len(grids) == 1:
            return dict(xyz=grids[0]["coords"], values=grids[0]["values"])
        # several grids (color0..3): channels side by side over the union of their active voxels
        allxyz = np.concatenate([g["coords"] for g in grids])
        xyz, inv = np.unique(allxyz, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        vals = np.zeros((xyz.shape[0], sum(g["components"] for g in grids)), np.float32)
        c0 = r0 = 0
        for g in grids:
            n = g["coords"].shape[0]
            vals[inv[r0:r0 + n], c0:c0 + g["components"]] = g["values"]
            r0 += n
            c0 += g["components"]
        return dict(xyz=xyz.astype(np.int32), values=vals)
    z = np.load(io.BytesIO(data), allow_pickle=False)
    if str(z["magic"]) != MAGIC:
        raise ValueError("%s is neither a .vdb file nor a plenvdb_b200 plane container" % path)
    return dict(xyz=z["xyz"], values=z["values"])


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


def save_dense_as_vdb(path, dense, name="density"):
    """FloatGrid with active voxels where `dense` != 0 (pyopenvdb `copyFromArray` semantics with background 0, as
    plenvdb/vdb_compression.py:49-55 stores the merged index grid)."""
    from . import openvdb_io
    a = np.ascontiguousarray(np.asarray(dense, np.float32))
    topo = Topology.from_mask(a != 0, device="cpu")
    xyz, leaf, off = _active_coords(topo)
    plane = np.zeros((max(topo.n_leaf, 1), 512, 1), np.float32)
    plane[leaf, off, 0] = a[xyz[:, 0], xyz[:, 1], xyz[:, 2]]
    openvdb_io.write_vdb(path, topo, [(name, plane)])


def load_vdb_as_dense(path, shape=None):
    """First grid of a .vdb file as a dense float32 array anchored at the origin (background 0 elsewhere)."""
    z = _read(path)
    xyz, vals = z["xyz"], z["values"]
    if shape is None:
        shape = tuple(int(v) + 1 for v in xyz.max(0)) if xyz.shape[0] else (1, 1, 1)
    out = np.zeros(tuple(shape), np.float32)
    ok = ((xyz >= 0) & (xyz < np.asarray(shape))).all(1)
    out[xyz[ok, 0], xyz[ok, 1], xyz[ok, 2]] = vals[ok, 0]
    return out

"""Grid maintenance around the hot loop, on the device.

The reference does each of these through dense host arrays (`get_dense_grid` -> torch op -> `copyFromDense`, ~200 MB per
call at 160^3): `VDBGrid.scale_volume_grid` (plenvdb/lib/grid.py:91-101), `DirectVoxGO.update_occupancy_cache`
(plenvdb/lib/dvgo.py:201-210), the (disabled) `total_variation_add_grad` (grid.py:103-106,
lib/cuda/total_variation_kernel.cu:14-35), and pruning on `load_from` (plenvdb.h:126-142).  Here they run on the sparse
planes through the C-ABI (csrc/maintenance.cu)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .tree import Topology


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def scale_volume_grid(vdb, new_world_size):
    """-> a new grid of the same class at `new_world_size` (dense-fill topology, like `DensityVDB(new_world_size)` +
    `copyFromDense(F.interpolate(dense, size, mode='trilinear', align_corners=True))` in grid.py:91-101)."""
    new = [int(x) for x in new_world_size]
    out = type(vdb).__new__(type(vdb))
    out.num, out.reso, out.ndim, out.device, out.timer = vdb.num, new, vdb.ndim, vdb.device, 0.0
    out._set_topology(Topology.dense(new, device=vdb.device))
    _lib.call("pvdb_resample_trilinear", vdb.topo.ref, _lib.ptr(vdb.grid), vdb.ndim, vdb.reso[0], vdb.reso[1], vdb.reso[2],
              out.topo.ref, _lib.ptr(out.grid), new[0], new[1], new[2], _lib.current_stream())
    return out


def resparsify(vdbs, keep_mask, extra_planes=()):
    """Drop the leaves in which `keep_mask` (bool [reso]) has no voxel.  `vdbs`: grids sharing one topology; they are
    switched to the pruned topology in place (values carried over, gradients cleared).  `extra_planes`: further planes on the
    old topology (optimiser moments); their remapped versions are returned in the same order."""
    old = vdbs[0].topo
    m = keep_mask.cpu().numpy() if torch.is_tensor(keep_mask) else np.asarray(keep_mask)
    new = Topology.from_mask(m.astype(bool), device=vdbs[0].device)
    st = _lib.current_stream()
    moved = []
    for p in extra_planes:
        q = new.new_plane(p.shape[-1])
        _lib.call("pvdb_plane_remap", old.ref, _lib.ptr(p), new.ref, _lib.ptr(q), p.shape[-1], st)
        moved.append(q)
    for v in vdbs:
        assert v.topo is old, "grids must share the topology that is being pruned"
        src = v.grid
        v._set_topology(new)
        _lib.call("pvdb_plane_remap", old.ref, _lib.ptr(src), new.ref, _lib.ptr(v.grid), v.ndim, st)
    torch.cuda.current_stream().synchronize()   # `old` and its planes may be released by the caller now
    return new, moved


def update_occupancy_cache(density, mask, params):
    """mask (uint8 / bool cuda tensor [mx,my,mz]) &= maxpool3(alpha(density at the mask voxel centres)) > fast_color_thres,
    in place (dvgo.py:201-210).  `params`: xyz_min, xyz_max, act_shift, interval, fast_color_thres."""
    assert mask.is_cuda and mask.is_contiguous()
    m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    tmp = torch.empty(m8.shape, dtype=torch.float32, device=m8.device)
    _lib.call("pvdb_occupancy_update", density.topo.ref, _lib.ptr(density.grid), density.reso[0], density.reso[1], density.reso[2],
              _f3(params["xyz_min"]), _f3(params["xyz_max"]), float(params["act_shift"]), float(params["interval"]),
              float(params["fast_color_thres"]), _lib.ptr(m8), m8.shape[0], m8.shape[1], m8.shape[2], _lib.ptr(tmp), _lib.current_stream())
    return mask


def total_variation_add_grad(vdb, wx, wy, wz, dense_mode=True):
    """grad += TV gradient (total_variation_kernel.cu:14-35) with the dense kernel's semantics on the sparse planes."""
    _lib.call("pvdb_total_variation_add_grad", vdb.topo.ref, _lib.ptr(vdb.grid), _lib.ptr(vdb.grad), vdb.ndim, vdb.reso[0], vdb.reso[1],
              vdb.reso[2], float(wx), float(wy), float(wz), int(bool(dense_mode)), _lib.current_stream())


@torch.no_grad()
def voxel_count_views(world_size, xyz_min, xyz_max, rays_o_tr, rays_d_tr, imsz, near, far, stepsize, voxel_size, downrate=1,
                      irregular_shape=False, device="cuda", ray_chunk=65536):
    """DirectVoxGO.voxel_count_views (plenvdb/lib/dvgo.py:212-243) on the device: for every training view, scatter the
    trilinear weights of all its sample points into a grid of ones' gradient and count the voxels whose accumulated weight
    exceeds 1.  The reference builds a DenseGrid and back-propagates through F.grid_sample per 10 000 rays; here the sample
    points go through the sparse gradient scatter (`pvdb_sample_backward`, C = 1) on a dense-fill topology.
    Returns the count as a float tensor [1, 1, rx, ry, rz] like the reference."""
    from .plenvdb import DensityVDB
    dev = torch.device(device)
    ws = [int(v) for v in world_size]
    lo = torch.as_tensor(np.asarray(xyz_min, np.float32)).to(dev)
    hi = torch.as_tensor(np.asarray(xyz_max, np.float32)).to(dev)
    ones = DensityVDB(ws, 1, device=device)
    far = 1e9                                                    # dvgo.py:214
    n_samples = int(np.linalg.norm(np.array(ws) + 1) / stepsize) + 1
    rng = torch.arange(n_samples, dtype=torch.float32, device=dev)[None]
    count = torch.zeros(ws, dtype=torch.float32, device=dev)
    scale = torch.tensor([w - 1 for w in ws], dtype=torch.float32, device=dev)
    if not irregular_shape and torch.is_tensor(rays_o_tr) and rays_o_tr.dim() == 4:
        views = zip(rays_o_tr, rays_d_tr)                        # [n_views, H, W, 3]
    else:
        views = zip(rays_o_tr.split(imsz), rays_d_tr.split(imsz))
    for ro_v, rd_v in views:
        if irregular_shape:
            ro_v, rd_v = ro_v.to(dev).reshape(-1, 3), rd_v.to(dev).reshape(-1, 3)
        else:
            ro_v = ro_v[::downrate, ::downrate].to(dev).flatten(0, -2)
            rd_v = rd_v[::downrate, ::downrate].to(dev).flatten(0, -2)
        ones.grad.zero_()
        for ro, rd in zip(ro_v.split(ray_chunk), rd_v.split(ray_chunk)):
            vec = torch.where(rd == 0, torch.full_like(rd, 1e-6), rd)
            rate_a, rate_b = (hi - ro) / vec, (lo - ro) / vec
            t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
            step = stepsize * voxel_size * rng
            interpx = t_min[..., None] + step / rd.norm(dim=-1, keepdim=True)
            pts = ro[..., None, :] + rd[..., None, :] * interpx[..., None]          # [n, S, 3] world
            idx = ((pts - lo) / (hi - lo) * scale).reshape(-1, 3).t().contiguous()    # DenseGrid: grid_sample, align_corners=True
            ones.backward_torch(idx, torch.ones(idx.shape[1], dtype=torch.float32, device=dev))
        count += (ones.get_dense_grid_torch(ones.grad)[..., 0] > 1).float()
    return count[None, None]

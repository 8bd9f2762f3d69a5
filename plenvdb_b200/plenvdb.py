"""Host-side mirror of the reference's pybind11 module ``plenvdb``
(plenvdb/lib/vdb/plenvdb.cpp:3-172, classes in plenvdb/lib/vdb/plenvdb.h).

Same class names, method names, argument meaning and numpy in/out contract, so the reference's callers
(plenvdb/lib/grid.py, masked_adam.py, run.py) can ``import plenvdb`` from here unchanged
(see INTEGRATION.md).  Every numeric method goes through the C-ABI into the sm_100a kernels; nothing is
computed on the CPU.  ``*_torch`` methods are the zero-copy additions (CUDA tensors in/out, no host
round trip, no synchronisation).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import vdbio
from .tree import Topology


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


class _BaseVDB:
    """BaseVDB<SorVGrid> (plenvdb.h:389-429): a value grid and a congruent gradient grid."""

    def __init__(self, resolution, n, device="cuda"):
        self.reso = [int(resolution[0]), int(resolution[1]), int(resolution[2])]
        self.ndim = int(n)
        self.device = torch.device(device)
        self.timer = 0.0
        self._set_topology(Topology.dense(self.reso, device=self.device))

    def _set_topology(self, topo, grid=None, grad=None):
        """The ONE place where a grid changes its tree or its planes.  `topo_version` counts the changes: objects that hold
        state congruent with the old planes (optimiser moments, the fused trainer's raw pointers) compare it before every use
        and rebuild instead of indexing the new leaf order with old planes (the reference has this as a latent bug on
        load_from)."""
        self.topo = topo
        self.grid = topo.new_plane(self.ndim) if grid is None else grid   # value plane  [n_leaf,512,ndim]
        self.grad = topo.new_plane(self.ndim) if grad is None else grad   # gradient plane
        assert self.grid.shape[0] == max(topo.n_leaf, 1) and self.grad.shape == self.grid.shape
        self.topo_version = getattr(self, "topo_version", 0) + 1

    # ---- maintenance on the device (the reference composes these from dense host round trips)
    # The reference's VDB classes do not implement the TV regulariser: the method prints "Not Supported Now..." and returns
    # (plenvdb.h:497-501, 574-578; grid.py:103-106 returns before calling it).  A drop-in must not silently change what a run
    # with non-zero TV weights computes, so that stays the default; set `enable_total_variation = True` (class or instance)
    # to run the dense kernel's semantics (total_variation_kernel.cu:14-35) on the sparse planes instead.
    enable_total_variation = False

    def total_variation_add_grad(self, wx, wy, wz, dense_mode=True):
        if not self.enable_total_variation:
            print("Not Supported Now...")
            return
        from . import maintenance
        maintenance.total_variation_add_grad(self, wx, wy, wz, dense_mode)

    def scale_volume_grid(self, new_world_size):
        """In-place form of VDBGrid.scale_volume_grid (grid.py:91-101): this object becomes the resampled dense-fill grid."""
        from . import maintenance
        new = maintenance.scale_volume_grid(self, new_world_size)
        self.reso = new.reso
        self._set_topology(new.topo, new.grid, new.grad)

    # ---- info / timers (plenvdb.h:397-399, 423-428)
    def getndim(self):
        return self.ndim

    def resetTimer(self):
        self.timer = 0.0

    def getTimer(self):
        return self.timer

    def setReso(self, rx, ry, rz):
        self.reso = [int(rx), int(ry), int(rz)]

    def getinfo(self):
        print("Resolution: (%d,%d,%d)" % tuple(self.reso))
        print("Num: %d" % self.num)
        print("Dim: %d" % self.ndim)

    # ---- zero-copy device API
    def forward_torch(self, pts, corner_out=None):
        """pts: float32 CUDA tensor [3,N] of index-space coords (as QueryVerticalInVDB builds them,
        plenvdb/lib/grid.py:82) -> [N, ndim].  corner_out=(leaf[N,8], off[N,8]) int32 for parity tests."""
        assert pts.is_cuda and pts.dtype == torch.float32 and pts.dim() == 2 and pts.shape[0] == 3
        pts = pts.contiguous()
        n = pts.shape[1]
        out = torch.empty((n, self.ndim), dtype=torch.float32, device=pts.device)
        cl, co = corner_out if corner_out is not None else (None, None)
        _lib.call("pvdb_sample_forward", self.topo.ref, _lib.ptr(self.grid), self.ndim, _lib.ptr(pts[0]), _lib.ptr(pts[1]),
                  _lib.ptr(pts[2]), n, _lib.ptr(out), _lib.ptr(cl), _lib.ptr(co), _lib.current_stream())
        return out

    def backward_torch(self, pts, grad_out):
        assert pts.is_cuda and grad_out.is_cuda
        pts = pts.contiguous()
        g = grad_out.contiguous().reshape(-1)
        n = pts.shape[1]
        assert g.numel() == n * self.ndim
        _lib.call("pvdb_sample_backward", self.topo.ref, _lib.ptr(self.grad), self.ndim, _lib.ptr(pts[0]), _lib.ptr(pts[1]),
                  _lib.ptr(pts[2]), _lib.ptr(g), n, _lib.current_stream())

    # ---- reference API: host numpy buffers (plenvdb.cpp:10-31, 47-67)
    def forward(self, x, y, z):
        x, y, z = _f32(x), _f32(y), _f32(z)
        n = x.size
        res = np.empty(n * self.ndim, dtype=np.float32)
        _lib.call("pvdb_sample_forward_host", self.topo.ref, _lib.ptr(self.grid), self.ndim, x.ctypes.data, y.ctypes.data,
                  z.ctypes.data, n, res.ctypes.data, _lib.current_stream())
        return res

    def backward(self, x, y, z, g):
        x, y, z, g = _f32(x), _f32(y), _f32(z), _f32(g)
        n = x.size
        assert g.size == n * self.ndim
        _lib.call("pvdb_sample_backward_host", self.topo.ref, _lib.ptr(self.grad), self.ndim, x.ctypes.data, y.ctypes.data,
                  z.ctypes.data, g.ctypes.data, n, _lib.current_stream())

    def forward_single(self, i, j, k):
        """forward_single (plenvdb.h:503-523): nearest value at integer coords."""
        ii = torch.as_tensor(np.ascontiguousarray(i, dtype=np.int32), device=self.device)
        jj = torch.as_tensor(np.ascontiguousarray(j, dtype=np.int32), device=self.device)
        kk = torch.as_tensor(np.ascontiguousarray(k, dtype=np.int32), device=self.device)
        out = torch.empty((ii.numel(), self.ndim), dtype=torch.float32, device=self.device)
        _lib.call("pvdb_sample_nearest", self.topo.ref, _lib.ptr(self.grid), self.ndim, _lib.ptr(ii), _lib.ptr(jj),
                  _lib.ptr(kk), ii.numel(), _lib.ptr(out), _lib.current_stream())
        return out.reshape(-1).cpu().numpy()

    # ---- dense <-> sparse (plenvdb.h:408-422, 149-167, 241-271)
    def copyFromDense_torch(self, dense, plane=None):
        dense = dense.contiguous().to(self.device, torch.float32)
        assert dense.numel() == self.reso[0] * self.reso[1] * self.reso[2] * self.ndim
        plane = self.grid if plane is None else plane
        _lib.call("pvdb_copy_from_dense", self.topo.ref, _lib.ptr(plane), plane.shape[-1], _lib.ptr(dense), self.reso[0],
                  self.reso[1], self.reso[2], _lib.current_stream())

    def copyFromDense(self, arr):
        arr = _f32(arr)
        assert arr.size == self.reso[0] * self.reso[1] * self.reso[2] * self.ndim
        self.copyFromDense_torch(torch.from_numpy(arr))

    def get_dense_grid_torch(self, plane=None):
        plane = self.grid if plane is None else plane
        ch = plane.shape[-1]
        dense = torch.empty((self.reso[0], self.reso[1], self.reso[2], ch), dtype=torch.float32, device=self.device)
        _lib.call("pvdb_copy_to_dense", self.topo.ref, _lib.ptr(plane), ch, _lib.ptr(dense), self.reso[0], self.reso[1],
                  self.reso[2], _lib.current_stream())
        return dense

    def get_dense_grid(self):
        return self.get_dense_grid_torch().reshape(-1).cpu().numpy()

    def copyToDense(self):
        return self.get_dense_grid()

    # ---- persistence (plenvdb.h:126-148, 211-240, 272-279)
    def save_to(self, path):
        vdbio.save_planes(path, self.topo, self.grid, self.reso, self._grid_names())

    def load_from(self, path):
        topo, plane, reso = vdbio.load_planes(path, self.ndim, self.device)
        # the reference re-creates only `grid`; we keep grad congruent by construction
        self._set_topology(topo, plane, None)
        self.reso = list(reso)


class DensityVDB(_BaseVDB):
    """DensityVDB (plenvdb.h:432-524)."""

    def __init__(self, resolution, n=1, device="cuda"):
        assert n == 1
        self.num = 1
        super().__init__(resolution, 1, device)

    def _grid_names(self):
        return ["density"]

    def setValuesOn_bymask_torch(self, mask, val):
        m = mask.contiguous().to(self.device).to(torch.uint8)
        assert m.numel() == self.reso[0] * self.reso[1] * self.reso[2]
        _lib.call("pvdb_set_values_on_by_mask", self.topo.ref, _lib.ptr(self.grid), _lib.ptr(m), float(val), self.reso[0],
                  self.reso[1], self.reso[2], _lib.current_stream())

    def setValuesOn_bymask(self, mask, val):
        self.setValuesOn_bymask_torch(torch.from_numpy(np.ascontiguousarray(mask).astype(np.uint8)), val)


class ColorVDB(_BaseVDB):
    """ColorVDB (plenvdb.h:527-602): n channels = n/3 Vec3f grids in the reference, one [.,512,n] plane here."""

    def __init__(self, resolution, n=12, device="cuda"):
        assert n % 3 == 0 and n > 0
        self.num = n // 3
        super().__init__(resolution, n, device)

    def _grid_names(self):
        return ["color%d" % d for d in range(self.num)]


class _BaseOptimizer:
    """BaseOptimizer (plenvdb.h:686-744): Adam moments congruent with the parameter grid."""

    def __init__(self, pvdb, lr, eps, beta0, beta1):
        self.params = pvdb
        self.lr, self.eps, self.beta0, self.beta1 = float(lr), float(eps), float(beta0), float(beta1)
        self.step_count = 0
        self.has_per_lr = False
        self.per_lr = None
        self._bind()

    def _bind(self):
        """(Re)allocate the moments for the parameter grid's CURRENT topology.  The reference builds the optimiser before
        model.load_from (utils.py:64-76; run.py:386-388) and --no_reload_optimizer skips opt.load_from: after the grid got
        another tree its moments must not be indexed with the old leaf order, so they restart from zero (what a fresh optimiser
        has); a per-voxel lr plane is dropped for the same reason and must be set again."""
        p = self.params
        self.exp_avg = p.topo.new_plane(p.ndim)
        self.exp_avg_sq = p.topo.new_plane(p.ndim)
        if self.per_lr is not None:
            self.per_lr, self.has_per_lr = None, False
        self._bound = (p.topo_version, p.topo)

    def _check_bound(self):
        p = self.params
        if self._bound[0] != p.topo_version or self._bound[1] is not p.topo:
            self._bind()

    # scalars (plenvdb.h:697-706)
    def getStep(self): return self.step_count
    def getLr(self): return self.lr
    def getEps(self): return self.eps
    def getBeta0(self): return self.beta0
    def getBeta1(self): return self.beta1
    def setStep(self, x): self.step_count = int(x)
    def setLr(self, x): self.lr = float(x)
    def setEps(self, x): self.eps = float(x)
    def setBeta0(self, x): self.beta0 = float(x)
    def setBeta1(self, x): self.beta1 = float(x)

    def update_lr(self, factor):
        # float multiply like `lr *= factor` on a C++ float (plenvdb.h:712)
        self.lr = float(np.float32(self.lr) * np.float32(factor))

    def set_grad(self, arr):
        self.params.copyFromDense_torch(torch.from_numpy(_f32(arr)), plane=self.params.grad)

    def set_pervoxel_lr(self, count):
        self._check_bound()
        p = self.params
        arr = _f32(count)
        assert arr.size == p.reso[0] * p.reso[1] * p.reso[2]
        self.per_lr = p.topo.new_plane(1)
        p.copyFromDense_torch(torch.from_numpy(arr), plane=self.per_lr)
        self.has_per_lr = True

    def zero_grad(self):
        p = self.params
        _lib.call("pvdb_zero_grad", p.topo.ref, _lib.ptr(p.grad), p.ndim, _lib.current_stream())

    def stepsize(self):
        return float(_lib.lib.pvdb_adam_stepsize(self.lr, self.beta0, self.beta1, self.step_count))

    def step(self, stepmode):
        """step_optimizer (plenvdb.h:751-767, 774-789)."""
        self._check_bound()
        p = self.params
        if stepmode == 2 and self.per_lr is None:
            raise RuntimeError("stepmode 2 needs set_pervoxel_lr() on the grid's current topology")
        self.step_count += 1
        _lib.call("pvdb_adam_step", p.topo.ref, _lib.ptr(p.grid), _lib.ptr(p.grad), _lib.ptr(self.exp_avg),
                  _lib.ptr(self.exp_avg_sq), p.ndim, int(stepmode), self.stepsize(), self.eps, self.beta0, self.beta1,
                  _lib.ptr(self.per_lr) if stepmode == 2 else None, _lib.current_stream())

    def save_to(self, prefix):
        p = self.params
        vdbio.save_planes(prefix + "exp_avg.vdb", p.topo, self.exp_avg, p.reso, p._grid_names())
        vdbio.save_planes(prefix + "exp_avg_sq.vdb", p.topo, self.exp_avg_sq, p.reso, p._grid_names())

    def load_from(self, prefix):
        self._check_bound()
        p = self.params
        self.exp_avg = vdbio.load_plane_onto(prefix + "exp_avg.vdb", p.topo, p.ndim, p.reso, p.device)
        self.exp_avg_sq = vdbio.load_plane_onto(prefix + "exp_avg_sq.vdb", p.topo, p.ndim, p.reso, p.device)

    def getinfo(self):
        p = self.params
        print("Resolution: (%d,%d,%d)" % tuple(p.reso))
        print("Num: %d" % p.num)
        print("Dim: %d" % p.ndim)
        print("Step: %d" % self.step_count)
        print("lr: %g" % self.lr)
        print("eps: %g" % self.eps)
        print("beta: %g,%g" % (self.beta0, self.beta1))


class DensityOpt(_BaseOptimizer):
    """DensityOpt (plenvdb.h:747-768)."""


class ColorOpt(_BaseOptimizer):
    """ColorOpt (plenvdb.h:770-790)."""




from .renderer import MGRenderer  # noqa: E402,F401  (plenvdb.h:933-1068)

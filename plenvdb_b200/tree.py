"""Sparse-tree topology on the device (pvdb_tree) and its torch-owned storage.

The topology is built once on the host by the library (csrc/topology.cpp), uploaded into torch
tensors, and shared by every payload plane that must stay congruent (value, grad, exp_avg,
exp_avg_sq, per-voxel lr) — the reference keeps four independent NanoVDB grids per parameter
(plenvdb/lib/vdb/plenvdb.h:438, 533, 707-710) and relies on them staying congruent.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Topology:
    """Device-resident 5-4-3 tree topology. ``c`` is the ``pvdb_tree`` passed to every kernel."""

    def __init__(self, handle, reso, device):
        self.reso = tuple(int(r) for r in reso)
        nu, nl, nf = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.call("pvdb_topo_counts", handle, C.byref(nu), C.byref(nl), C.byref(nf))
        self.n_upper, self.n_lower, self.n_leaf = nu.value, nl.value, nf.value
        root_keys = np.zeros(max(self.n_upper, 1), dtype=np.uint64)
        upper = np.full(max(self.n_upper, 1) * 32768, -1, dtype=np.int32)
        lower = np.full(max(self.n_lower, 1) * 4096, -1, dtype=np.int32)
        origin = np.zeros(max(self.n_leaf, 1) * 3, dtype=np.int32)
        mask = np.zeros(max(self.n_leaf, 1) * 8, dtype=np.uint64)
        _lib.call("pvdb_topo_export", handle, root_keys.ctypes.data, upper.ctypes.data, lower.ctypes.data,
                  origin.ctypes.data, mask.ctypes.data)
        _lib.lib.pvdb_topo_destroy(handle)
        self.device = torch.device(device)
        # host copies (tests, file IO); device copies (kernels)
        self.h_root_keys, self.h_upper, self.h_lower = root_keys, upper, lower
        self.h_leaf_origin = origin.reshape(-1, 3)
        self.h_leaf_mask = mask.reshape(-1, 8)
        up = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(self.device)
        self.d_root_keys, self.d_upper, self.d_lower = up(root_keys), up(upper), up(lower)
        self.d_leaf_origin, self.d_leaf_mask = up(origin), up(mask)
        self.c = _lib.pvdb_tree(
            n_upper=self.n_upper, n_lower=self.n_lower, n_leaf=self.n_leaf, reserved=0,
            root_key0=int(root_keys[0]) if self.n_upper else 0xFFFFFFFFFFFFFFFF,
            root_keys=self.d_root_keys.data_ptr(), upper_child=self.d_upper.data_ptr(),
            lower_child=self.d_lower.data_ptr(), leaf_origin=self.d_leaf_origin.data_ptr(),
            leaf_mask=self.d_leaf_mask.data_ptr())
        self.ref = C.byref(self.c)

    @classmethod
    def dense(cls, reso, device="cuda"):
        """denseFill(bbox, 0, active) topology (plenvdb.h:117-125)."""
        h = _lib.lib.pvdb_topo_create_dense(int(reso[0]), int(reso[1]), int(reso[2]))
        if not h:
            raise _lib.PvdbError(_lib.last_error())
        return cls(h, reso, device)

    @classmethod
    def from_mask(cls, active, device="cuda"):
        """Leaves wherever ``active`` (bool [rx,ry,rz]) has a set voxel; value mask = ``active``."""
        a = np.ascontiguousarray(np.asarray(active).astype(np.uint8))
        assert a.ndim == 3
        h = _lib.lib.pvdb_topo_create_from_mask(a.ctypes.data, a.shape[0], a.shape[1], a.shape[2])
        if not h:
            raise _lib.PvdbError(_lib.last_error())
        return cls(h, a.shape, device)

    def new_plane(self, channels=1):
        """Zero payload plane [n_leaf, 512, channels] (fp32) congruent with this topology."""
        return torch.zeros((max(self.n_leaf, 1), 512, channels), dtype=torch.float32, device=self.device)

    @property
    def n_voxel_slots(self):
        return self.n_leaf * 512

    def active_voxel_count(self):
        return int(sum(bin(int(w)).count("1") for w in self.h_leaf_mask[: self.n_leaf].reshape(-1)))

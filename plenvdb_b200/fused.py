"""Fused fine-stage trainer: the opt-in fast path behind the reference's call sequence
``model(rays) -> zero_grad -> loss.backward() -> optimizer.step() / vdbopt.step()`` (plenvdb/run.py:541-588).

``FusedTrainer`` owns (through torch) every device buffer the C-ABI entry point ``pvdb_train_step`` needs and
enqueues one whole iteration on the current stream with no host synchronisation, so the step can be captured
in a CUDA graph.  Grids are plain ``DensityVDB`` / ``ColorVDB`` objects from ``plenvdb_b200.plenvdb`` sharing
one topology, i.e. the same objects the drop-in API exposes.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .plenvdb import ColorVDB, DensityVDB
from .synth import mask_scale_shift  # noqa: F401  (numpy-only helper; re-exported for callers of this module)
from .tree import Topology

PHASE_FORWARD, PHASE_BACKWARD, PHASE_UPDATE, PHASE_LISTS_READY, PHASE_ACCUMULATE = 1, 2, 4, 8, 16
NET_N = 22019


def get_rays_of_a_view(H, W, K, c2w, inverse_y=False, flip_x=False, flip_y=False, device=None):
    """dvgo.get_rays_of_a_view (dvgo.py:470-499, 518-526) with mode 'center', no NDC: rays_o, rays_d, viewdirs as float32
    [H, W, 3] torch tensors on `device`.  Same torch expressions as the reference, so the rays carry the same bits."""
    c2w = torch.as_tensor(np.asarray(c2w.cpu() if torch.is_tensor(c2w) else c2w, np.float32)).reshape(-1)[:16].reshape(4, 4)
    K = np.asarray(K.cpu() if torch.is_tensor(K) else K, np.float32).reshape(3, 3)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    c2w = c2w.to(dev)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev), indexing="ij")
    i, j = i.t().float() + 0.5, j.t().float() + 0.5
    if flip_x:
        i = i.flip((1,))
    if flip_y:
        j = j.flip((0,))
    if inverse_y:
        dirs = torch.stack([(i - K[0][2]) / K[0][0], (j - K[1][2]) / K[1][1], torch.ones_like(i)], -1)
    else:
        dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    return rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()


class FusedTrainer:
    OVERFLOW_CHECK_EVERY = 64     # step_from_host reads the overflow flag of the sample lists every that many calls (and on the first)

    def __init__(self, params, density, k0, mask, net, n_rays, device="cuda", use_tensor_cores=True,
                 cap_alpha_per_ray=96, cap_keep_per_ray=64, parity_counts=False, n_rays_global=None,
                 scratch_per_ray=128, use_graph=False):
        """params: dict from synth.scene_params (or the equivalent run.py scalars).
        density: DensityVDB, k0: ColorVDB(12) on the SAME topology (k0.topo is density.topo).
        mask: bool [reso] mask_cache.mask.  net: float32[22019] packed rgbnet parameters."""
        assert k0.topo is density.topo, "density and k0 must share one topology"
        self.P = dict(params)
        self.dev = torch.device(device)
        self.density, self.k0 = density, k0
        topo = density.topo
        self.topo = topo
        self._bound = (density.topo_version, k0.topo_version)
        self.n_rays = int(n_rays)
        self.k0_dim = int(k0.ndim)
        self.direct = self.k0_dim == 3       # coarse stage: 3 colour channels, no rgbnet (dvgo.py:344-346)
        assert self.k0_dim in (3, 12), "k0 must have 12 channels (fine stage) or 3 (coarse stage)"
        self.use_tc = bool(use_tensor_cores) and not self.direct
        self.step_count = 0
        self.launches_total = 0      # kernels of this library launched through this trainer
        f32 = dict(dtype=torch.float32, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        # optimiser state (DensityOpt / ColorOpt / MaskedAdam equivalents)
        self.den_m, self.den_v = topo.new_plane(1), topo.new_plane(1)
        self.k0_m, self.k0_v = topo.new_plane(self.k0_dim), topo.new_plane(self.k0_dim)
        self.den_perlr = None                # per-voxel lr plane of stepmode 2 (set_pervoxel_lr)
        if net is None:
            assert self.direct, "the fine stage needs the packed rgbnet parameters"
            net = np.zeros(NET_N, np.float32)
        self.net = torch.as_tensor(np.asarray(net, np.float32)).to(self.dev).contiguous()
        assert self.net.numel() == NET_N
        self.net_grad, self.net_m, self.net_v = (torch.zeros(NET_N, **f32) for _ in range(3))
        self.lr_density, self.lr_k0, self.lr_net = float(params["lr_density"]), float(params["lr_k0"]), float(params["lr_net"])
        # occupancy bits
        self.set_mask(mask)
        n = self.n_rays
        ca, ck = int(cap_alpha_per_ray) * n, int(cap_keep_per_ray) * n
        self.cap_alpha, self.cap_keep = ca, ck
        z = lambda *shape, **kw: torch.zeros(*shape, **kw)
        self.t = dict(
            t_min=z(n, **f32), t_max=z(n, **f32), n_steps=z(n, **i32), cnt_mask=z(n, **i32), cnt_alpha=z(n, **i32),
            cnt_keep=z(n, **i32), cnt_alpha_full=z(n, **i32), off_alpha=z(n + 1, **i32), off_keep=z(n + 1, **i32),
            alphainv_last=z(n, **f32), rgb_marched=z(n, 3, **f32), grad_last=z(n, **f32),
            s_ray=z(ca, **i32), s_step=z(ca, **i32), s_xyz=z(ca, 3, **f32), s_density=z(ca, **f32), s_alpha=z(ca, **f32),
            s_T=z(ca, **f32), s_weight=z(ca, **f32), s_gden=z(ca, **f32),
            k_sample=z(ck, **i32), k_ray=z(ck, **i32), k_xyz=z(ck, 3, **f32), k_feat=z(ck, 12, **f32), k_rgb=z(ck, 3, **f32),
            k_gw=z(ck, **f32),
            den_touched=z(max(topo.n_leaf, 1), **i32), k0_touched=z(max(topo.n_leaf, 1), **i32),
            den_touched_list=z(max(topo.n_leaf, 1), **i32), k0_touched_list=z(max(topo.n_leaf, 1), **i32),
            counters=z(16, **i32), loss=z(4, **f32),
        )
        # activations kept for the fp32 rgbnet backward (the tcgen05 forward still pairs with it)
        # the tcgen05 kernels write and read these tensors in whole 128-sample tiles: rows rounded up to a tile (cap_keep itself
        # is 64 * n_rays, not a multiple of 128 for an odd batch)
        ckp = (ck + 127) // 128 * 128
        if self.direct:   # no rgbnet: no activations to keep
            self.t["k_h0"] = self.t["k_h1"] = z(1, **f32)
        else:
            self.t["k_h0"] = z(ckp, 128, **f32)
            self.t["k_h1"] = z(ckp, 128, **f32)
        if self.use_tc or self.direct:
            self.t["k_corner"] = z(ck, 8, **i32)     # record ids of the eight corners, saved by the march
        if self.use_tc:   # tensor-core backward: input rows and masked activation gradients for the weight-gradient GEMM
            self.t["k_x"] = z(ckp, 40, **f32)
            self.t["k_dh0"] = z(ckp, 128, **f32)
            self.t["k_mask"] = z(ckp, 8, **i32)
            self.t["net_img"] = z(512 * 1024 // 4, **i32)
            self.t["net_partial"] = z(148, 22048, **f32)
            self.t["ray_pe"] = z(n, 28, **f32)       # view-direction embedding per ray (the forward reads it per kept sample)
        if self.use_tc:   # leaf-local alternative of the k0 gather / scatter (csrc/leaf_local.cu): used only when switched on
            nl = max(topo.n_leaf, 1)
            self.t["ll_cnt"], self.t["ll_off"], self.t["ll_cur"], self.t["ll_list"] = z(nl, **i32), z(nl + 1, **i32), z(nl, **i32), z(nl, **i32)
            self.t["ll_items"] = z(ck, **i32)
            self.t["k_dx"] = z(ck, 12, **f32)
        self.scratch_per_ray = int(scratch_per_ray)
        if self.scratch_per_ray > 0:
            self.t["march_scratch"] = z(20 * n * self.scratch_per_ray, **i32)
        self.parity_counts = bool(parity_counts)
        self.n_rays_global = int(n_rays_global) if n_rays_global else self.n_rays
        self._bufs = None
        self._graphs = {}          # staging buffer address -> captured graph of the whole iteration (+ its pinned scalar slot)
        self._dstage = None
        self.use_graph = bool(use_graph)
        self._build_structs()

    # ---- state
    def set_mask(self, mask):
        if torch.is_tensor(mask):
            m = mask.to(self.dev, torch.uint8).contiguous()
        else:
            m = torch.as_tensor(np.ascontiguousarray(np.asarray(mask).astype(np.uint8))).to(self.dev)
        self.mask_shape = tuple(m.shape)
        nb = [(s + 7) // 8 for s in self.mask_shape]
        nblk = nb[0] * nb[1] * nb[2]
        self.occ_fine = torch.zeros(nblk * 8, dtype=torch.int64, device=self.dev)
        self.occ_coarse = torch.zeros((nblk + 63) // 64, dtype=torch.int64, device=self.dev)
        _lib.call("pvdb_occ_build", _lib.ptr(m), m.shape[0], m.shape[1], m.shape[2], _lib.ptr(self.occ_fine),
                  _lib.ptr(self.occ_coarse), _lib.current_stream())
        self.mask_dev = m
        if getattr(self, "_bufs", None) is not None:
            self._build_structs()

    def update_occupancy_cache(self):
        """dvgo.py:201-210 on the device: prune the mask with the current density and rebuild the occupancy bits."""
        from . import maintenance
        maintenance.update_occupancy_cache(self.density, self.mask_dev, self.P)
        self.set_mask(self.mask_dev)

    def _build_structs(self):
        P = self.P
        c = _lib.pvdb_train_cfg()
        c.xyz_min = (C.c_float * 3)(*[float(v) for v in P["xyz_min"]])
        c.xyz_max = (C.c_float * 3)(*[float(v) for v in P["xyz_max"]])
        c.reso = (C.c_int32 * 3)(*[int(v) for v in P["reso"]])
        c.mask_reso = (C.c_int32 * 3)(*self.mask_shape)
        sc, sh = mask_scale_shift(self.mask_shape, P["xyz_min"], P["xyz_max"])
        c.mask_scale = (C.c_float * 3)(*[float(v) for v in sc])
        c.mask_shift = (C.c_float * 3)(*[float(v) for v in sh])
        for k in ("near", "far", "stepdist", "act_shift", "interval", "fast_color_thres", "bg", "weight_main",
                  "weight_entropy_last", "weight_rgbper", "eps", "beta0", "beta1", "den_mode", "k0_mode"):
            setattr(c, k, P[k])
        c.k0_dim, c.net_width = (3, 0) if self.direct else (12, 128)
        c.use_tensor_cores = int(self.use_tc)
        c.n_rays_global = self.n_rays_global
        c.parity_counts = int(self.parity_counts)
        self.cfg = c
        b = _lib.pvdb_train_bufs()
        b.tree = C.pointer(self.topo.c)
        p = lambda t: t.data_ptr()
        b.den, b.den_grad, b.den_m, b.den_v = p(self.density.grid), p(self.density.grad), p(self.den_m), p(self.den_v)
        b.k0, b.k0_grad, b.k0_m, b.k0_v = p(self.k0.grid), p(self.k0.grad), p(self.k0_m), p(self.k0_v)
        b.occ_fine, b.occ_coarse = p(self.occ_fine), p(self.occ_coarse)
        b.net, b.net_grad, b.net_m, b.net_v = p(self.net), p(self.net_grad), p(self.net_m), p(self.net_v)
        b.cap_alpha, b.cap_keep = self.cap_alpha, self.cap_keep
        for k, t in self.t.items():
            setattr(b, k, p(t))
        b.scratch_rays, b.scratch_per_ray = self.n_rays, self.scratch_per_ray
        b.step_scalars = None
        b.den_perlr = self.den_perlr.data_ptr() if self.den_perlr is not None else None
        self._bufs = b
        self._graphs = {}           # captured graphs hold the old pointers

    def _set_step_scalars(self):
        s = self.step_count
        self.cfg.den_stepsz = _lib.lib.pvdb_adam_stepsize(self.lr_density, self.P["beta0"], self.P["beta1"], s)
        self.cfg.k0_stepsz = _lib.lib.pvdb_adam_stepsize(self.lr_k0, self.P["beta0"], self.P["beta1"], s)
        self.cfg.net_lr = self.lr_net
        self.cfg.net_step = s

    def _step_graphed(self, stage):
        """One full iteration on the staging buffer `stage` [4, n, 3] as a CUDA-graph replay.  Issued one by one, the ~25 launches,
        events and the side-stream fork / join of pvdb_train_step cost the host ~95 us and leave ~1 us gaps between the kernels
        on the device; a replay costs the host ~10 us and runs the same kernels 8 % faster (measured: 183 vs 199 us per
        iteration, scratch/graph_timing.py).  The graph is captured on the second call for a staging buffer (the first runs
        directly, which also initialises what the library creates lazily).  The per-iteration scalars (Adam step sizes with their
        bias corrections, after any lr decay) are read by the update kernels from a pinned host word of this graph
        (pvdb_train_bufs.step_scalars), written just before the replay: the caller must not have more than one iteration per
        staging buffer in flight (step_from_host synchronises; step_from_host_async alternates two buffers and waits for
        iteration i - 1 before it issues i + 1; step() waits for the previous replay of its buffer)."""
        self._check_bound()
        key = (stage.data_ptr(), stage.shape[1])
        g = self._graphs.get(key)
        if g is None:                                    # first call for this buffer: run directly
            self._graphs[key] = dict(graph=None, scal=torch.zeros(4, dtype=torch.float32).pin_memory(), done=torch.cuda.Event())
            self.run(stage[0], stage[1], stage[2], stage[3], PHASE_FORWARD | PHASE_BACKWARD | PHASE_UPDATE)
            return
        self.step_count += 1
        self._set_step_scalars()
        h = g["scal"]
        h[0], h[1] = self.cfg.den_stepsz, self.cfg.k0_stepsz
        h[2] = _lib.lib.pvdb_dense_adam_stepsize_host(self.lr_net, self.P["beta0"], self.P["beta1"], self.step_count)
        if g["graph"] is None:                           # second call: capture
            graph = torch.cuda.CUDAGraph()
            self._bufs.step_scalars = h.data_ptr()       # pinned host memory, device-readable at the same address (UVA)
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    _lib.call("pvdb_train_step", C.byref(self.cfg), C.byref(self._bufs), _lib.ptr(stage[0]), _lib.ptr(stage[1]),
                              _lib.ptr(stage[2]), _lib.ptr(stage[3]), stage.shape[1], PHASE_FORWARD | PHASE_BACKWARD | PHASE_UPDATE,
                              _lib.current_stream())
            finally:
                self._bufs.step_scalars = None           # direct launches keep taking the scalars from cfg
            g["graph"], g["launches"] = graph, int(_lib.lib.pvdb_last_launch_count())
        g["graph"].replay()
        self.launches_total += g["launches"]

    def set_pervoxel_lr(self, count):
        """VDBAdam.set_pervoxel_lr (masked_adam.py:43-46): `count` is the dense view-count grid [reso]; the density step of
        stepmode 2 is scaled by count / count.max() per voxel."""
        c = torch.as_tensor(np.asarray(count.cpu() if torch.is_tensor(count) else count, np.float32)).to(self.dev)
        per = (c / c.max()).reshape(-1).contiguous()
        self.den_perlr = self.topo.new_plane(1)
        self.density.copyFromDense_torch(per.reshape(self.density.reso), plane=self.den_perlr)
        self._build_structs()

    def decay_lr(self, factor):
        """run.py:592-598: multiply every lr by `factor` (float32 for the grids, like the C++ members)."""
        self.lr_density = float(np.float32(self.lr_density) * np.float32(factor))
        self.lr_k0 = float(np.float32(self.lr_k0) * np.float32(factor))
        self.lr_net = self.lr_net * factor

    def _check_bound(self):
        """The buffer struct holds raw pointers into the grids' planes and a copy of the tree: a grid that got another topology
        or other planes since (load_from, scale_volume_grid, maintenance.resparsify) must be bound again before any launch."""
        if self._bound != (self.density.topo_version, self.k0.topo_version):
            raise RuntimeError("the density / k0 grids changed their topology or planes after this FusedTrainer was built: call "
                               "rebind() (Adam moments restart from zero unless remapped planes are passed) before stepping")

    def rebind(self, den_m=None, den_v=None, k0_m=None, k0_v=None):
        """Bind the trainer to the grids' CURRENT topology and planes.  The grid Adam moments restart from zero unless planes
        congruent with the new topology are given (maintenance.resparsify returns the remapped ones)."""
        assert self.k0.topo is self.density.topo, "density and k0 must share one topology"
        topo = self.density.topo
        self.topo = topo
        self.den_m = topo.new_plane(1) if den_m is None else den_m
        self.den_v = topo.new_plane(1) if den_v is None else den_v
        self.k0_m = topo.new_plane(self.k0_dim) if k0_m is None else k0_m
        self.k0_v = topo.new_plane(self.k0_dim) if k0_v is None else k0_v
        self.den_perlr = None      # congruent with the old tree: set_pervoxel_lr again
        i32 = dict(dtype=torch.int32, device=self.dev)
        for k in ("den_touched", "k0_touched", "den_touched_list", "k0_touched_list") + (("ll_cnt", "ll_cur", "ll_list") if "ll_cnt" in self.t else ()):
            self.t[k] = torch.zeros(max(topo.n_leaf, 1), **i32)
        if "ll_off" in self.t:
            self.t["ll_off"] = torch.zeros(max(topo.n_leaf, 1) + 1, **i32)
        self._bound = (self.density.topo_version, self.k0.topo_version)
        self._build_structs()

    # ---- execution
    def run(self, rays_o, rays_d, viewdirs, target, phases):
        self._check_bound()
        n = rays_o.shape[0]
        assert n <= self.n_rays
        for t in (rays_o, rays_d, viewdirs) + ((target,) if target is not None else ()):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        if phases & PHASE_UPDATE:
            self.step_count += 1
            self._set_step_scalars()
        if (phases & PHASE_BACKWARD) and getattr(self, "_grads_pending", False):
            phases |= PHASE_ACCUMULATE        # a backward without update came before: its leaves stay on the touched lists
        _lib.call("pvdb_train_step", C.byref(self.cfg), C.byref(self._bufs), _lib.ptr(rays_o), _lib.ptr(rays_d),
                  _lib.ptr(viewdirs), _lib.ptr(target), n, int(phases), _lib.current_stream())
        self.launches_total += int(_lib.lib.pvdb_last_launch_count())
        if phases & PHASE_UPDATE:
            self._grads_pending = False
        elif phases & PHASE_BACKWARD:
            self._grads_pending = True

    def step(self, rays_o, rays_d, viewdirs, target):
        """One full iteration: forward + backward + sparse Adam (grids) + Adam (rgbnet).  With use_graph=True the batch is
        gathered into a staging buffer by one kernel and the iteration is a CUDA-graph replay (_step_graphed); the host blocks
        only if the replay before the previous one has not finished yet.  Default off — measured on B200
        (scratch/graph_timing.py): back-to-back replays alone run at 183-194 us per iteration against 198 us for the direct
        launches (which already overlap kernel boundaries through programmatic dependent launch), but the per-iteration host work
        between replays (scalars, staging) brings a replayed loop to 203-214 us; a strictly synchronous loop gains 10 us
        (224 -> 214 us).  Per-kernel profiling always issues the kernels directly."""
        n = rays_o.shape[0]
        if not self.use_graph or _lib.PROFILING or n != self.n_rays:
            self.run(rays_o, rays_d, viewdirs, target, PHASE_FORWARD | PHASE_BACKWARD | PHASE_UPDATE)
            return
        for t in (rays_o, rays_d, viewdirs, target):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        if self._dstage is None:       # two staging buffers (two graphs): the host issues iteration i while i - 1 runs
            self._dstage = [torch.empty((4, n, 3), dtype=torch.float32, device=self.dev) for _ in range(2)]
            self._dstage_i = 0
        st = self._dstage[self._dstage_i & 1]
        self._dstage_i += 1
        g = self._graphs.get((st.data_ptr(), n))
        if g is not None:
            g["done"].synchronize()          # iteration i - 2 has finished: this graph's pinned scalar slot is free again
        _lib.call("pvdb_stage_rays", _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(viewdirs), _lib.ptr(target), n, _lib.ptr(st),
                  _lib.current_stream())
        self.launches_total += 1
        self._step_graphed(st)
        self._graphs[(st.data_ptr(), n)]["done"].record(torch.cuda.current_stream())

    def step_from_host(self, batch_host, stepper=None):
        """One iteration fed from HOST memory, the way run.py:541-588 is driven: `batch_host` is a pinned float32 tensor
        [4, n, 3] = (rays_o, rays_d, viewdirs, target).  One H2D copy, the fused step, one D2H of the four loss words;
        returns them as a pinned host tensor (valid after this call: it synchronises the stream, like `loss.item()`; it is
        overwritten by the next call).  stepper: callable(rays_o, rays_d, viewdirs, target) issuing the iteration (default
        self.step; DataParallelTrainer.step for a sharded run)."""
        n = batch_host.shape[1]
        if getattr(self, "_stage", None) is None or self._stage.shape[1] != n:
            self._stage = torch.empty((4, n, 3), dtype=torch.float32, device=self.dev)
            self._loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
        self._stage.copy_(batch_host, non_blocking=True)
        if stepper is None and self.use_graph and not _lib.PROFILING:
            self._step_graphed(self._stage)
        else:
            (stepper or self.step)(self._stage[0], self._stage[1], self._stage[2], self._stage[3])
        self._host_calls = getattr(self, "_host_calls", 0) + 1
        check = self._host_calls % self.OVERFLOW_CHECK_EVERY == 1
        if check:      # at logging cadence: the sample-list overflow flag travels with the loss (one more 64-byte D2H, same sync)
            if getattr(self, "_cnt_host", None) is None:
                self._cnt_host = torch.zeros(16, dtype=torch.int32).pin_memory()
            self._cnt_host.copy_(self.t["counters"], non_blocking=True)
        self._loss_host.copy_(self.t["loss"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if check and int(self._cnt_host[3]) != 0:
            raise RuntimeError("sample lists overflowed (M_alpha %d > %d or M_keep %d > %d): rays were truncated and gradients dropped; "
                               "build the trainer with larger cap_alpha_per_ray / cap_keep_per_ray" % (
                                   int(self._cnt_host[0]), self.cap_alpha, int(self._cnt_host[1]), self.cap_keep))
        return self._loss_host

    def step_from_host_async(self, batch_host, stepper=None):
        """Pipelined variant of step_from_host for a training loop that does not need the loss of iteration i before it issues
        iteration i + 1 (run.py only prints it every few hundred iterations).  The H2D copy of `batch_host` (pinned float32
        [4, n, 3]) goes through a copy stream into one of two staging buffers, so it overlaps the previous iteration still
        running; the step waits for it on the current stream; the loss words follow into one of two pinned buffers.  Returns
        the PREVIOUS call's loss words (None on the first call) after waiting for that iteration only — the GPU always has the
        next iteration queued behind the running one.  host_pipeline_flush() returns the last one.  `batch_host` may be refilled
        as soon as the call returns (its copy has completed); a returned loss tensor is overwritten two calls later.
        stepper: callable(rays_o, rays_d, viewdirs, target) issuing the iteration (default self.step; DataParallelTrainer.step
        for a sharded run)."""
        n = batch_host.shape[1]
        p = getattr(self, "_pipe", None)
        if p is None or p["n"] != n:
            p = dict(n=n, i=0, copy=torch.cuda.Stream(device=self.dev),
                     stage=[torch.empty((4, n, 3), dtype=torch.float32, device=self.dev) for _ in range(2)],
                     loss=[torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)],
                     copied=[torch.cuda.Event() for _ in range(2)], done=[torch.cuda.Event() for _ in range(2)], used=[False, False])
            self._pipe = p
        k = p["i"] & 1
        cur = torch.cuda.current_stream()
        if p["used"][k]:
            p["copy"].wait_event(p["done"][k])      # the iteration that read stage[k] (two calls ago) has finished
        with torch.cuda.stream(p["copy"]):
            p["stage"][k].copy_(batch_host, non_blocking=True)
            p["copied"][k].record(p["copy"])
        cur.wait_event(p["copied"][k])
        st = p["stage"][k]
        if stepper is None and self.use_graph and not _lib.PROFILING:
            self._step_graphed(st)           # iteration i - 2, the previous user of this buffer's graph, was waited for one call ago
        else:
            (stepper or self.step)(st[0], st[1], st[2], st[3])
        p["loss"][k].copy_(self.t["loss"], non_blocking=True)
        p["done"][k].record(cur)
        p["used"][k] = True
        p["i"] += 1
        p["copied"][k].synchronize()         # the caller may refill `batch_host` as soon as this returns (~10 us, under the running iteration)
        if p["used"][k ^ 1]:
            p["done"][k ^ 1].synchronize()
            return p["loss"][k ^ 1]
        return None

    def host_pipeline_flush(self):
        """Wait for the last iteration issued through step_from_host_async and return its loss words (pinned host tensor)."""
        p = getattr(self, "_pipe", None)
        if p is None or p["i"] == 0:
            return None
        k = (p["i"] - 1) & 1
        p["done"][k].synchronize()
        return p["loss"][k]

    def step_dp(self, peers, dp_step, rays_o, rays_d, viewdirs, target):
        """One data-parallel iteration (pvdb_train_step_dp): the NVLink tile exchange overlaps the weight-gradient kernel."""
        self._check_bound()
        n = rays_o.shape[0]
        assert n <= self.n_rays
        for t in (rays_o, rays_d, viewdirs, target):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        self.step_count += 1
        self._set_step_scalars()
        _lib.call("pvdb_train_step_dp", C.byref(self.cfg), C.byref(self._bufs), C.byref(peers), int(dp_step), _lib.ptr(rays_o),
                  _lib.ptr(rays_d), _lib.ptr(viewdirs), _lib.ptr(target), rays_o.shape[0], _lib.current_stream())
        self.launches_total += int(_lib.lib.pvdb_last_launch_count())

    def forward_backward(self, rays_o, rays_d, viewdirs, target):
        self.run(rays_o, rays_d, viewdirs, target, PHASE_FORWARD | PHASE_BACKWARD)

    def update(self, lists_ready=False):
        """Apply the optimisers to whatever gradients are accumulated (after a gradient all-reduce).  lists_ready: the
        touched-leaf lists were already rebuilt by pvdb_dp_exchange."""
        self._check_bound()
        dummy = self.t["t_min"]
        self.step_count += 1
        self._set_step_scalars()
        _lib.call("pvdb_train_step", C.byref(self.cfg), C.byref(self._bufs), _lib.ptr(dummy), _lib.ptr(dummy), _lib.ptr(dummy),
                  None, self.n_rays, PHASE_UPDATE | (PHASE_LISTS_READY if lists_ready else 0), _lib.current_stream())
        self.launches_total += int(_lib.lib.pvdb_last_launch_count())
        self._grads_pending = False

    def forward(self, rays_o, rays_d, viewdirs):
        """Render rays through the training model (run.py:171-189); returns rgb_marched [n,3]."""
        self.run(rays_o, rays_d, viewdirs, None, PHASE_FORWARD)
        return self.t["rgb_marched"][: rays_o.shape[0]]

    def render_view(self, H, W, K, c2w, inverse_y=False, flip_x=False, flip_y=False):
        """The non-merged render path of run.py:171-189: the rays of one view through the TRAINING grids and rgbnet (forward
        phase of the fused step), in chunks of this trainer's batch size; returns rgb_marched as a [H, W, 3] CUDA tensor."""
        ro, rd, vd = get_rays_of_a_view(H, W, K, c2w, inverse_y, flip_x, flip_y, device=self.dev)
        ro, rd, vd = ro.reshape(-1, 3), rd.reshape(-1, 3), vd.reshape(-1, 3)
        out = torch.empty((H * W, 3), dtype=torch.float32, device=self.dev)
        for a in range(0, H * W, self.n_rays):
            b = min(a + self.n_rays, H * W)
            out[a:b] = self.forward(ro[a:b].contiguous(), rd[a:b].contiguous(), vd[a:b].contiguous())
        return out.reshape(H, W, 3)

    def hit_mask(self, rays_o, rays_d):
        """hit_coarse_geo (dvgo.py:253-270): bool [n] — does the ray touch the occupancy mask?"""
        n = rays_o.shape[0]
        hit = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        _lib.call("pvdb_rays_hit_mask", C.byref(self.cfg), C.byref(self._bufs), _lib.ptr(rays_o.contiguous()),
                  _lib.ptr(rays_d.contiguous()), n, _lib.ptr(hit), _lib.current_stream())
        return hit.bool()

    def launches_last_call(self):
        return int(_lib.lib.pvdb_last_launch_count())

    def counters(self):
        c = self.t["counters"].cpu().numpy()
        return dict(M_alpha=int(c[0]), M_keep=int(c[1]), n_touched_den=int(c[2]), overflow=int(c[3]), n_touched_k0=int(c[4]))


def build_scene_grids(scene, device="cuda"):
    """DensityVDB + ColorVDB(12) on one shared topology, filled from a synth.make_scene dict."""
    reso = scene["reso"]
    den = DensityVDB(list(reso), 1, device=device)
    if scene["active"] is not None:
        den._set_topology(Topology.from_mask(scene["active"], device=device))
    k0 = ColorVDB.__new__(ColorVDB)
    k0.num, k0.reso, k0.ndim, k0.device, k0.timer = 4, list(reso), 12, torch.device(device), 0.0
    k0._set_topology(den.topo)
    den.copyFromDense_torch(torch.from_numpy(scene["density"]))
    k0.copyFromDense_torch(torch.from_numpy(scene["k0"]))
    return den, k0


def build_stress_scene(reso=512, device="cuda", bound=1.3, seed=6, half_thickness=0.055):
    """SURVEY.md §8(d) cfg 5 (S512): noisy thick shell (half_thickness 0.055 -> ~5 % of the voxels, ~10 % of the 262 144 leaf
    blocks at 512^3, as the survey specifies) on a PRUNED topology, built on the device — the
    dense host arrays synth.make_scene works with would be 6.4 GB at 512^3.  Same formulas as synth.shell_occupancy /
    make_scene (occupancy, N(6,1) density inside / -10 outside, U(-1,1) k0 on the dilated occupancy, mask = 3^3 max-pool),
    torch RNG instead of numpy's.  Returns (params, density grid, k0 grid, mask [reso^3] uint8 cuda)."""
    from . import synth
    dev = torch.device(device)
    P = synth.scene_params(reso, bound=bound)
    R = reso
    ax = torch.linspace(-bound, bound, R, dtype=torch.float64, device=dev)
    X, Y, Z = ax.view(R, 1, 1), ax.view(1, R, 1), ax.view(1, 1, R)
    ph = torch.from_numpy(np.random.default_rng(seed).uniform(0, 2 * np.pi, (4, 3))).to(dev)
    r = torch.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    noise = torch.zeros((R, R, R), dtype=torch.float64, device=dev)
    for k in range(4):
        f = (2 ** k) * 3.0
        noise += 0.5 ** k * torch.sin(f * X + ph[k, 0]) * torch.sin(f * Y + ph[k, 1]) * torch.sin(f * Z + ph[k, 2])
    occ = (r - 0.8 - 0.06 * noise).abs() <= half_thickness
    del noise, r
    mask = torch.nn.functional.max_pool3d(occ[None, None].to(torch.float16), 3, 1, 1)[0, 0] > 0
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    dens = torch.full((R, R, R), -10.0, dtype=torch.float32, device=dev)
    dens[occ] = 6.0 + torch.randn(int(occ.sum()), generator=g, device=dev)
    den = DensityVDB.__new__(DensityVDB)      # skip the dense-fill topology of the constructor (262 144 leaves at 512^3)
    den.num, den.reso, den.ndim, den.device, den.timer = 1, [R, R, R], 1, dev, 0.0
    den._set_topology(Topology.from_mask(mask.cpu().numpy(), device=device))   # pruned: leaves where the mask has voxels
    k0 = ColorVDB.__new__(ColorVDB)
    k0.num, k0.reso, k0.ndim, k0.device, k0.timer = 4, [R, R, R], 12, dev, 0.0
    k0._set_topology(den.topo)
    dens[~mask] = 0.0            # outside the tree: background
    den.copyFromDense_torch(dens)
    del dens
    g.manual_seed(2)
    k0d = torch.zeros((R, R, R, 12), dtype=torch.float32, device=dev)
    k0d[mask] = torch.rand((int(mask.sum()), 12), generator=g, device=dev) * 2 - 1
    k0.copyFromDense_torch(k0d)
    del k0d
    P.update(variant="pruned", occupied_fraction=float(occ.float().mean()), n_leaf=den.topo.n_leaf)
    return P, den, k0, mask.to(torch.uint8)

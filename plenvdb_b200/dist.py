"""Multi-GPU paths; the reference has none (single process, one GPU: run.py:31,696).

Training: data parallel, one process per GPU, replicated grids and rgbnet, ray batch sharded.  One exchange per
iteration between the backward and the update phases of the fused step:
  1. all-reduce(MAX) of the per-leaf touched flags (density and k0)            -> union of touched leaves
  2. gather the union leaves' gradient tiles into one contiguous buffer [n_union, 512, 13] + 22019 rgbnet grads
  3. all-reduce(SUM) of that buffer (NCCL over NVLink), scatter back
after which every rank runs the identical sparse Adam.  The losses are means over the GLOBAL batch
(cfg.n_rays_global), so the summed shard gradients equal the single-GPU full-batch gradients.

Rendering: replicas + row sharding.  Default on CUDA: interleaved row groups whose pixels each rank's composite kernel
stores straight into rank 0's frame buffer over NVLink peer memory (PeerFrame; no collective at all).  Fallback: contiguous
row bands + an NCCL gather.
"""
import torch
import torch.distributed as dist

from .fused import PHASE_BACKWARD, PHASE_FORWARD, FusedTrainer


def init_from_env(backend=None):
    """torchrun-style init: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n items; the union over ranks is exactly range(n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def union_touched(den_touched, k0_touched, group=None):
    """In-place MAX all-reduce of the two flag arrays (NCCL has no OR); returns the sorted union leaf ids."""
    flags = torch.maximum(den_touched, k0_touched)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    den_touched.copy_(flags)
    k0_touched.copy_(flags)
    return torch.nonzero(flags, as_tuple=False).reshape(-1)


def allreduce_sparse_grads(den_grad, k0_grad, net_grad, leaves, group=None):
    """SUM all-reduce of the gradient tiles of `leaves` (+ rgbnet grads) through one packed buffer."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return 0
    n = leaves.numel()
    nd, nk = n * 512, n * 512 * 12
    buf = torch.empty(nd + nk + net_grad.numel(), dtype=torch.float32, device=den_grad.device)
    buf[:nd] = den_grad[leaves].reshape(-1)
    buf[nd:nd + nk] = k0_grad[leaves].reshape(-1)
    buf[nd + nk:] = net_grad
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    den_grad[leaves] = buf[:nd].reshape(n, 512, 1)
    k0_grad[leaves] = buf[nd:nd + nk].reshape(n, 512, 12)
    net_grad.copy_(buf[nd + nk:])
    return buf.numel() * 4


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank cannot set up the CUDA-IPC symmetric blocks."""


class PeerExchange:
    """The NVLink peer-memory exchange (csrc/dp_exchange.cu): one symmetric block per rank, mapped by every peer through
    CUDA IPC; the handles travel over torch.distributed (plumbing), the gradients never touch NCCL."""

    def __init__(self, n_leaf, cap_leaves=None, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib, self._group = C, _lib, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        cap = int(cap_leaves) if cap_leaves else max(int(n_leaf), 1)
        nbytes = _lib.lib.pvdb_dp_symm_bytes(int(n_leaf), cap)
        self._own, self._opened = None, []
        # collective; on any rank's failure every rank raises PeerExchangeUnavailable after everything was unmapped and freed
        self._own, bases, self._opened = open_symmetric_blocks(nbytes, group)
        self.peers = _lib.pvdb_dp_peers()
        self.peers.world, self.peers.rank, self.peers.n_leaf, self.peers.cap_leaves = self.world, self.rank, int(n_leaf), cap
        for r in range(self.world):
            self.peers.base[r] = bases[r]
        self.step = 0
        self.nbytes = nbytes
        # the gradient planes inside the own block: the exchange works on the ranks' planes in place
        pd, pk = C.c_void_p(), C.c_void_p()
        _lib.call("pvdb_dp_grad_planes", C.byref(self.peers), C.byref(pd), C.byref(pk))
        dev = torch.device("cuda", torch.cuda.current_device())
        nl = max(int(n_leaf), 1)
        self.den_grad = _lib.tensor_from_ptr(pd.value, (nl, 512, 1), dev)
        self.k0_grad = _lib.tensor_from_ptr(pk.value, (nl, 512, 12), dev)

    def bind(self, trainer):
        """Make the trainer's grids accumulate their gradients in this block's planes (they are what the peers read and write)."""
        den, k0 = trainer.density, trainer.k0
        self.den_grad.copy_(den.grad)
        self.k0_grad.copy_(k0.grad)
        den._set_topology(den.topo, den.grid, self.den_grad)
        k0._set_topology(k0.topo, k0.grid, self.k0_grad)
        trainer.rebind(trainer.den_m, trainer.den_v, trainer.k0_m, trainer.k0_v)

    def unbind(self, trainer):
        """Give the grids ordinary gradient planes again (before the block is freed)."""
        den, k0 = trainer.density, trainer.k0
        if den.grad.data_ptr() == self.den_grad.data_ptr():
            den._set_topology(den.topo, den.grid, self.den_grad.clone())
            k0._set_topology(k0.topo, k0.grid, self.k0_grad.clone())
            trainer.rebind(trainer.den_m, trainer.den_v, trainer.k0_m, trainer.k0_v)

    def exchange(self, bufs):
        self._lib.call("pvdb_dp_exchange", self._C.byref(self.peers), self._C.byref(bufs), self.step, self._lib.current_stream())
        self.step += 1

    def error(self):
        e = self._C.c_int32(0)
        self._lib.call("pvdb_dp_symm_error", self._C.byref(self.peers), self._C.byref(e))
        return int(e.value)

    def close(self):
        """Collective: every rank calls it.  The stream is drained, the peers' blocks are unmapped on every rank, and only after a
        barrier does a rank free its own block (CUDA IPC leaves freeing a block that a peer still maps undefined)."""
        if self._own is None and not self._opened:
            return
        self.den_grad = self.k0_grad = None
        close_symmetric_blocks(self._own, self._opened, self._group, barrier=True)
        self._own, self._opened = None, []


class DataParallelTrainer:
    """FusedTrainer + the gradient exchange.  `n_rays` is the PER-RANK shard size.

    exchange="nvlink" (default on CUDA): pvdb_train_step_dp — the union of the ranks' touched leaves (device-side barrier) runs
    under the rgbnet forward, pack / reduce-scatter + all-gather over NVLink peer memory / unpack under the weight-gradient
    kernel, the rgbnet gradients ride on the weight-gradient reduction; no NCCL, no host sync; identical bits on every rank.
    exchange="nccl": MAX all-reduce of the touched flags -> pvdb_dp_pack -> one host read of the union size -> ONE SUM
    all-reduce of the packed tiles + rgbnet gradients -> pvdb_dp_unpack -> update."""

    def __init__(self, params, density, k0, mask, net, n_rays, world=None, group=None, exchange="nvlink", **kw):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.tr = FusedTrainer(params, density, k0, mask, net, n_rays, n_rays_global=n_rays * self.world, **kw)
        self._init_exchange(exchange)

    def _init_exchange(self, exchange):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        tr = self.tr
        n_leaf = max(tr.topo.n_leaf, 1)
        dev = tr.dev
        self.exchange = exchange
        self.peer = None
        self.last_exchange_bytes = 0
        if self.world > 1 and exchange == "nvlink":
            try:
                self.peer = PeerExchange(tr.topo.n_leaf, group=self.group)
                self.peer.bind(tr)
                return
            except PeerExchangeUnavailable as e:      # raised on all ranks together: fall back to NCCL everywhere
                import warnings
                warnings.warn("%s; using the NCCL all-reduce of packed tiles instead" % e)
                self.exchange = "nccl"
        self.flags = torch.zeros(2 * n_leaf, dtype=torch.int32, device=dev)
        self.union_list = torch.zeros(n_leaf, dtype=torch.int32, device=dev)
        self.union_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.union_count_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.cap = n_leaf * 512 * 13 + 22019
        self.buf = torch.empty(self.cap, dtype=torch.float32, device=dev)

    @classmethod
    def wrap(cls, trainer, world, group=None, exchange="nvlink"):
        """Turn an existing single-GPU FusedTrainer into the data-parallel one (its loss means become global)."""
        self = cls.__new__(cls)
        self.group, self.world, self.tr = group, world, trainer
        trainer.n_rays_global = trainer.n_rays * world
        trainer._build_structs()
        self._init_exchange(exchange)
        return self

    def step(self, rays_o, rays_d, viewdirs, target):
        tr, C, _lib = self.tr, self._C, self._lib
        if self.world == 1:
            tr.step(rays_o, rays_d, viewdirs, target)
            return
        if self.peer is not None:
            tr.step_dp(self.peer.peers, self.peer.step, rays_o, rays_d, viewdirs, target)
            self.peer.step += 1
            return
        tr.run(rays_o, rays_d, viewdirs, target, PHASE_FORWARD | PHASE_BACKWARD)
        n_leaf = tr.topo.n_leaf
        # 1. union of touched leaves: both flag arrays in one MAX all-reduce
        self.flags[:n_leaf].copy_(tr.t["den_touched"][:n_leaf])
        self.flags[n_leaf:2 * n_leaf].copy_(tr.t["k0_touched"][:n_leaf])
        dist.all_reduce(self.flags, op=dist.ReduceOp.MAX, group=self.group)
        tr.t["den_touched"][:n_leaf].copy_(self.flags[:n_leaf])
        tr.t["k0_touched"][:n_leaf].copy_(self.flags[n_leaf:2 * n_leaf])
        # 2. pack on the device, read the union size (the only host sync of the step)
        st = _lib.current_stream()
        _lib.call("pvdb_dp_pack", C.byref(tr._bufs), _lib.ptr(self.union_list), _lib.ptr(self.union_count),
                  C.c_void_p(self.union_count_host.data_ptr()), _lib.ptr(self.buf), self.cap, st)
        torch.cuda.current_stream().synchronize()
        n = int(self.union_count_host[0])
        numel = n * 512 * 13 + 22019
        # 3. one SUM all-reduce over NVLink, then unpack + identical update everywhere
        dist.all_reduce(self.buf[:numel], op=dist.ReduceOp.SUM, group=self.group)
        _lib.call("pvdb_dp_unpack", C.byref(tr._bufs), _lib.ptr(self.union_list), _lib.ptr(self.union_count), _lib.ptr(self.buf), st)
        tr.launches_total += 4
        tr.update()
        self.last_exchange_bytes = numel * 4 + self.flags.numel() * 4

    def exchange_bytes(self):
        """Bytes this rank moved over NVLink in the last exchange (peer reads, or the all-reduce payload)."""
        if self.peer is not None:
            n, w = int(self.tr.t["counters"][2].item()), self.world
            own = (n - self.peer.rank + w - 1) // w if n > self.peer.rank else 0       # union slots this rank reduces
            tiles = 2 * (w - 1) * own * 512 * 13 * 4          # peer loads of the owned leaves + peer stores of their sums
            return tiles + (w - 1) * 22019 * 4 + (w - 1) * self.tr.topo.n_leaf     # + rgbnet pushes + the peers' flag bytes
        return self.last_exchange_bytes

    def close(self):
        """Collective: releases the NVLink symmetric blocks (no-op for the NCCL exchange)."""
        if self.peer is not None:
            self.peer.unbind(self.tr)
            self.peer.close()
            self.peer = None


def open_symmetric_blocks(nbytes, group=None):
    """One cudaMalloc'ed block of `nbytes` per rank, mapped by every peer through CUDA IPC.  The handles travel over
    torch.distributed (plumbing).  Returns (own c_void_p, [base address of rank r's block in THIS process], [opened mappings]).
    Every rank takes part in both collectives whatever happens locally and all ranks agree on the outcome: on any failure
    PeerExchangeUnavailable is raised on EVERY rank after whatever was set up has been released."""
    import ctypes as C
    from . import _lib
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert world <= 8, "one NVSwitch domain (<= 8 GPUs)"
    own, handle, opened, bases, err = C.c_void_p(), C.create_string_buffer(64), [], [None] * world, None
    try:
        _lib.call("pvdb_dp_symm_alloc", int(nbytes), C.byref(own), handle)
    except _lib.PvdbError as e:
        err, own = str(e), None
    handles = [None] * world
    dist.all_gather_object(handles, None if err else bytes(handle.raw), group=group)
    if err is None and all(h is not None for h in handles):
        try:
            for r in range(world):
                if r == rank:
                    bases[r] = own.value
                else:
                    p = C.c_void_p()
                    _lib.call("pvdb_dp_symm_open", C.create_string_buffer(handles[r], 64), C.byref(p))
                    bases[r] = p.value
                    opened.append(p)
        except _lib.PvdbError as e:
            err = str(e)
    elif err is None:
        err = "a peer could not allocate its symmetric block"
    oks = [None] * world
    dist.all_gather_object(oks, err, group=group)   # also the barrier: every block is zeroed and mapped before the first signal
    bad = [(r, e) for r, e in enumerate(oks) if e is not None]
    if bad:
        close_symmetric_blocks(own, opened, group, barrier=True)     # every rank is here: unmap everywhere, then free
        raise PeerExchangeUnavailable("NVLink peer memory unavailable (rank %d: %s)" % bad[0])
    return own, bases, opened


def close_symmetric_blocks(own, opened, group=None, barrier=False):
    """Unmap the peers' blocks, then free the own one.  barrier=True (collective: every rank must call it): no rank frees its
    block while a peer still has it mapped."""
    from . import _lib
    for p in opened:
        _lib.call("pvdb_dp_symm_close", p)
    if barrier and dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier(group=group)
    if own is not None:
        _lib.call("pvdb_dp_symm_free", own)


def interleaved_rows_of(H, band_rows, rank, world):
    """Image rows of `rank` when groups of band_rows rows are dealt round-robin to the ranks (row r -> rank (r // band_rows)
    % world), ascending — the order of the local rows of pvdb_render_rows_interleaved."""
    return [r for r in range(H) if (r // band_rows) % world == rank]


class PeerFrame:
    """Frame assembly over NVLink peer memory (csrc/renderer.cu: pvdb_render_frame_sharded): every rank renders interleaved
    groups of `band_rows` rows and its composite kernel stores them straight into root's frame buffer; one release/acquire
    signal per rank replaces the gather.  No NCCL call and no host synchronisation per frame."""

    def __init__(self, H, W, band_rows=16, root=0, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.H, self.W, self.band_rows, self.root = int(H), int(W), int(band_rows), int(root)
        self._own, self._opened, self._group = None, [], group
        self._own, bases, self._opened = open_symmetric_blocks(_lib.lib.pvdb_frame_symm_bytes(self.H, self.W), group)
        self.peers = _lib.pvdb_frame_peers()
        self.peers.world, self.peers.rank, self.peers.root, self.peers.H, self.peers.W = self.world, self.rank, self.root, self.H, self.W
        for r in range(self.world):
            self.peers.base[r] = bases[r]
        self.frame_no = 0
        self._views = None

    def render(self, renderer, c2w_dev):
        """All ranks call it with the same camera.  Root gets the full [H, W, 3] frame: a VIEW of its frame buffer that stays
        valid until root's call after the next one starts (consume it on the same stream before that); the others get None."""
        f = self.frame_no
        renderer.render_frame_sharded(self.peers, c2w_dev, self.band_rows, f)
        self.frame_no += 1
        if self.rank != self.root:
            return None
        if self._views is None:
            views = []
            for k in range(2):
                p = self._C.c_void_p()
                self._lib.call("pvdb_frame_ptr", self._C.byref(self.peers), k, self._C.byref(p))
                views.append(self._lib.tensor_from_ptr(p.value, (self.H, self.W, 3), renderer.dev))
            self._views = views
        return self._views[f & 1]

    def error(self):
        e = self._C.c_int32(0)
        self._lib.call("pvdb_frame_error", self._C.byref(self.peers), self._C.byref(e))
        return int(e.value)

    def close(self):
        """Collective: every rank calls it (the owners free their blocks only after all peers have unmapped them)."""
        self._views = None
        close_symmetric_blocks(self._own, self._opened, self._group, barrier=True)
        self._own, self._opened = None, []


def render_sharded(renderer, c2w_dev, rank, world, gather=True, group=None, peer=None):
    """Rank 0 gets the full [H, W, 3] frame.  peer: a PeerFrame — interleaved row groups written straight into rank 0's frame
    over NVLink (no collective).  Otherwise: contiguous row bands + an NCCL gather (gather=False: just this rank's band)."""
    if peer is not None and world > 1:
        return peer.render(renderer, c2w_dev)
    H, W = renderer.cfg.H, renderer.cfg.W
    lo, hi = shard_range(H, rank, world)
    band = renderer.render_rows_torch(c2w_dev, lo, hi)
    if world == 1 or not gather:
        return band
    sizes = [shard_range(H, r, world) for r in range(world)]
    max_rows = max(b - a for a, b in sizes)
    if band.shape[0] < max_rows:   # NCCL gather wants equal shapes: pad the short bands
        band = torch.cat([band, band.new_zeros((max_rows - band.shape[0], W, 3))], 0)
    if rank == 0:
        parts = [torch.empty((max_rows, W, 3), dtype=torch.float32, device=band.device) for _ in sizes]
        dist.gather(band, parts, dst=0, group=group)
        return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], 0)
    dist.gather(band, None, dst=0, group=group)
    return None

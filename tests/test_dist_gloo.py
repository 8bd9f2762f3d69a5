"""world_size-2 CPU (gloo) coverage of the multi-GPU host logic: touched-leaf union, packed sparse gradient
all-reduce, shard ranges and the row-band gather layout.  The kernels themselves are covered by the -m gpu tests; here the
exchange runs on CPU tensors, exactly the code path DataParallelTrainer.step uses between backward and update."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from plenvdb_b200 import dist as pdist
    r, _, w = pdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    n_leaf = 40
    rng = np.random.default_rng(100 + rank)
    # each rank touched a different subset of leaves and holds gradients only there
    touched = np.zeros(n_leaf, np.int32)
    mine = rng.choice(n_leaf, 12, replace=False)
    touched[mine] = 1
    den_grad = torch.zeros(n_leaf, 512, 1)
    k0_grad = torch.zeros(n_leaf, 512, 12)
    den_grad[mine] = torch.from_numpy(rng.standard_normal((12, 512, 1)).astype(np.float32))
    k0_grad[mine] = torch.from_numpy(rng.standard_normal((12, 512, 12)).astype(np.float32))
    net_grad = torch.from_numpy(rng.standard_normal(22019).astype(np.float32))
    local = (den_grad.clone(), k0_grad.clone(), net_grad.clone())
    dt, kt = torch.from_numpy(touched.copy()), torch.from_numpy(touched.copy())
    leaves = pdist.union_touched(dt, kt)
    nbytes = pdist.allreduce_sparse_grads(den_grad, k0_grad, net_grad, leaves)
    # reference: dense all-reduce of the local copies
    dense = [t.clone() for t in local]
    for t in dense:
        dist.all_reduce(t)
    ok = all(torch.equal(a, b) for a, b in zip((den_grad, k0_grad, net_grad), dense))
    # ranges
    lo, hi = pdist.shard_range(8192, rank, world)
    # row-band gather layout on CPU tensors
    H, W = 10, 4
    a, b = pdist.shard_range(H, rank, world)
    band = torch.full((b - a, W, 3), float(rank))
    parts = [torch.empty((b - a, W, 3)) for _ in range(world)] if rank == 0 else None
    dist.gather(band, parts, dst=0)
    q.put((rank, ok, leaves.tolist(), int(dt.sum()), nbytes, (lo, hi), None if parts is None else torch.cat(parts)[:, 0, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sparse_gradient_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ok0, leaves0, n0, bytes0, span0, img0), (r1, ok1, leaves1, n1, bytes1, span1, _) = out
    assert ok0 and ok1, "sparse exchange differs from the dense all-reduce"
    assert leaves0 == leaves1 and n0 == n1 == len(leaves0) and 12 <= n0 <= 24
    assert bytes0 == bytes1 == (n0 * 512 * 13 + 22019) * 4
    assert span0 == (0, 4096) and span1 == (4096, 8192)
    assert img0 == [0.0] * 5 + [1.0] * 5


def _peer_setup_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from plenvdb_b200 import dist as pdist
    pdist.init_from_env(backend="gloo")
    outcomes = []
    # without a CUDA device the symmetric blocks cannot be allocated: every rank must learn that in the same collective and
    # raise the same exception — nobody is left waiting for a peer's handle (the NVLink paths then fall back to NCCL together)
    for make in (lambda: pdist.open_symmetric_blocks(1 << 20), lambda: pdist.PeerFrame(64, 64), lambda: pdist.PeerExchange(40)):
        try:
            make()
            outcomes.append("ok")
        except pdist.PeerExchangeUnavailable as e:
            outcomes.append("unavailable: %s" % str(e)[:40])
    # the rows the sharded renderer gives each rank partition the image
    rows = pdist.interleaved_rows_of(122, 16, rank, world)
    all_rows = [None] * world
    dist.all_gather_object(all_rows, rows)
    q.put((rank, outcomes, sorted(sum(all_rows, [])) == list(range(122))))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_setup_fails_together_world2():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("exercises the no-device failure path; the working path is covered by tests/test_dp_gpu.py")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_setup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, outcomes, partition in out:
        assert len(outcomes) == 3 and all(o.startswith("unavailable") for o in outcomes), outcomes
        assert partition
    assert out[0][1] == out[1][1]      # the same diagnosis (first failing rank's message) everywhere

"""Two-GPU data-parallel training step (NCCL) equals the one-GPU step on the union batch, and the
row-band sharded render equals the single-GPU frame.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, exchange="nvlink"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from plenvdb_b200 import dist as pdist
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    pdist.init_from_env()
    dev = torch.device("cuda", rank)
    scene = synth.make_scene(96, "sparse")
    net = synth.rgbnet_init()
    rays = synth.ray_batch(2048, H=200, W=200, K=synth.intrinsics(200, 200), seed=777)
    lo, hi = pdist.shard_range(2048, rank, world)
    den, k0 = build_scene_grids(scene, device=dev)
    dp = pdist.DataParallelTrainer(scene, den, k0, scene["mask"], net, hi - lo, device=dev, exchange=exchange)
    shard = [torch.from_numpy(a[lo:hi].copy()).to(dev) for a in rays]
    for _ in range(2):
        dp.step(*shard)
    torch.cuda.synchronize()
    res = dict(den=den.get_dense_grid(), k0=k0.get_dense_grid(), net=dp.tr.net.cpu().numpy(), bytes=dp.exchange_bytes(),
               err=dp.peer.error() if dp.peer is not None else 0)
    if rank == 0:   # single-GPU reference on the full batch
        den1, k01 = build_scene_grids(scene, device=dev)
        tr = FusedTrainer(scene, den1, k01, scene["mask"], net, 2048, device=dev)
        full = [torch.from_numpy(a).to(dev) for a in rays]
        for _ in range(2):
            tr.step(*full)
        torch.cuda.synchronize()
        res.update(den1=den1.get_dense_grid(), k01=k01.get_dense_grid(), net1=tr.net.cpu().numpy())
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nvlink", "nccl"])
def test_two_gpu_dp_step_matches_single_gpu(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, exchange)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = out[0], out[1]
    assert a["err"] == 0 and b["err"] == 0
    # replicas stay identical (same reduced gradients, same Adam)
    assert np.array_equal(a["den"], b["den"]) and np.array_equal(a["k0"], b["k0"]) and np.array_equal(a["net"], b["net"])
    assert a["bytes"] > 0
    for k, k1 in (("den", "den1"), ("k0", "k01"), ("net", "net1")):
        np.testing.assert_allclose(a[k], a[k1], rtol=1e-3, atol=2e-4, err_msg=k)


def _frame_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from plenvdb_b200 import dist as pdist
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import build_scene_grids
    from plenvdb_b200.plenvdb import MGRenderer
    from plenvdb_b200.renderer import merge_grids
    pdist.init_from_env()
    dev = torch.device("cuda", rank)
    scene = synth.make_scene(96, "dense")
    den, k0 = build_scene_grids(scene, device=dev)
    dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(synth.rgbnet_init())
    H, W = 122, 120
    r = MGRenderer(12, 27, 128, 3, device=dev)
    r.load_data_dense(dend, cold, idx)
    r.load_params(np.ascontiguousarray(w0.T).reshape(-1), b0, np.ascontiguousarray(w1.T).reshape(-1), b1,
                  np.ascontiguousarray(w2.T).reshape(-1), b2)
    r.setScene(list(scene["reso"]), synth.intrinsics(H, W).reshape(-1), scene["xyz_min"], scene["xyz_max"])
    r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"],
                False, H, W)
    poses = torch.from_numpy(synth.render_cameras(8).reshape(8, 16)).to(dev)
    peer = pdist.PeerFrame(H, W, band_rows=4)
    frames = []
    for i in range(6):      # six frames: both frame buffers are reused twice, with no host synchronisation in between
        img = pdist.render_sharded(r, poses[i], rank, world, peer=peer)
        if rank == 0:
            frames.append(img.clone())    # stream-ordered consumption before the buffer's next use
    torch.cuda.synchronize()
    res = dict(err=peer.error(), mismatched=[])
    if rank == 0:
        for i, img in enumerate(frames):
            full = r.render_rows_torch(poses[i], 0, H)
            if not torch.equal(img, full):
                res["mismatched"].append(i)
        res["lit"] = float((frames[2] != scene["bg"]).float().mean())
    dist.barrier()
    peer.close()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_peer_frame_matches_single_gpu():
    """pvdb_render_frame_sharded: interleaved row groups, peer stores into rank 0's frame over NVLink, signals instead of a
    gather.  The assembled frames equal the single-GPU frames bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frame_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out[0]["err"] == 0 and out[1]["err"] == 0
    assert out[0]["mismatched"] == []
    assert out[0]["lit"] > 0.02, "degenerate view"


def test_dp_pack_unpack_roundtrip_single_gpu():
    """pvdb_dp_pack / pvdb_dp_unpack on one GPU: union list = sorted set of touched leaves, packed tiles equal the gradient
    planes, and unpack writes back exactly what the buffer holds (here 2x, as a 2-rank sum of equal shards would)."""
    import ctypes as C
    from plenvdb_b200 import _lib, synth
    from plenvdb_b200 import dist as pdist
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids, PHASE_FORWARD, PHASE_BACKWARD
    dev = torch.device("cuda", 0)
    scene = synth.make_scene(96, "sparse")
    net = synth.rgbnet_init()
    rays = [torch.from_numpy(a).to(dev) for a in synth.ray_batch(1024, H=200, W=200, K=synth.intrinsics(200, 200), seed=5)]
    den, k0 = build_scene_grids(scene, device=dev)
    tr = FusedTrainer(scene, den, k0, scene["mask"], net, 1024, device=dev)
    dp = pdist.DataParallelTrainer.wrap(tr, 1)
    tr.run(*rays, PHASE_FORWARD | PHASE_BACKWARD)
    n_leaf = tr.topo.n_leaf
    flags = ((tr.t["den_touched"][:n_leaf] | tr.t["k0_touched"][:n_leaf]) != 0)
    want = torch.nonzero(flags).flatten().cpu().numpy()
    dg, kg, ng = tr.density.grad.clone(), tr.k0.grad.clone(), tr.net_grad.clone()
    st = _lib.current_stream()
    _lib.call("pvdb_dp_pack", C.byref(tr._bufs), _lib.ptr(dp.union_list), _lib.ptr(dp.union_count),
              C.c_void_p(dp.union_count_host.data_ptr()), _lib.ptr(dp.buf), dp.cap, st)
    torch.cuda.synchronize()
    n = int(dp.union_count_host[0])
    assert n == len(want) and n > 0
    got = dp.union_list[:n].cpu().numpy()
    assert np.array_equal(got, want)   # ascending: identical slot order on every rank
    tiles = dp.buf[:n * 6656].reshape(n, 6656)
    gl = torch.from_numpy(got.astype(np.int64)).to(dev)
    assert torch.equal(tiles[:, :512], dg.reshape(-1, 512)[gl])
    assert torch.equal(tiles[:, 512:], kg.reshape(-1, 512 * 12)[gl])
    assert torch.equal(dp.buf[n * 6656:n * 6656 + 22019], ng)
    dp.buf[:n * 6656 + 22019] *= 2
    _lib.call("pvdb_dp_unpack", C.byref(tr._bufs), _lib.ptr(dp.union_list), _lib.ptr(dp.union_count), _lib.ptr(dp.buf), st)
    torch.cuda.synchronize()
    assert torch.equal(tr.density.grad, dg * 2) and torch.equal(tr.k0.grad, kg * 2)
    assert torch.equal(tr.net_grad, ng * 2)

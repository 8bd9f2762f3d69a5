"""GPU parity of the fused fine-stage training step (SURVEY.md §8a rows S1..N3 composed as in dvgo.py:296-388 +
run.py:541-588) against the CPU oracle on the same seeded scene, rays and rgbnet.
Integers (sample counts, segment offsets, ray/step ids, voxel/leaf indices) bit-exact; values within 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle_state(scene):
    from oracle import oracle as orc
    R, act = scene["reso"], scene["active"]
    den, k0 = orc.Grid(R, 1, act), orc.Grid(R, 12, act)
    den.copy_from_dense(scene["density"])
    k0.copy_from_dense(scene["k0"])
    aux = [orc.Grid(R, c, act) for c in (1, 1, 1, 12, 12, 12)]   # den_grad, den_m, den_v, k0_grad, k0_m, k0_v
    return den, k0, aux


def _oracle_cfg(P, step=1, do_update=1, n_rays_global=0, threads=8):
    keys = ["xyz_min", "xyz_max", "reso", "near", "far", "stepdist", "act_shift", "interval", "fast_color_thres", "bg",
            "weight_main", "weight_entropy_last", "weight_rgbper", "lr_density", "lr_k0", "lr_net", "eps", "beta0", "beta1",
            "den_mode", "k0_mode"]
    c = {k: P[k] for k in keys}
    c.update(step=step, do_update=do_update, n_rays_global=n_rays_global, threads=threads)
    return c


def _run_oracle(scene, net, rays, step=1, do_update=1, cap_keep=0, n_rays_global=0):
    from oracle import oracle as orc
    den, k0, aux = _oracle_state(scene)
    net = net.copy()
    nm, nv = np.zeros_like(net), np.zeros_like(net)
    out = orc.train_step(_oracle_cfg(scene, step, do_update, n_rays_global), den, aux[0], aux[1], aux[2], k0, aux[3], aux[4],
                         aux[5], scene["mask"], net, nm, nv, *rays, cap_keep=cap_keep)
    return out, den, k0, aux, net, nm, nv


def _trainer(scene, net, n_rays, **kw):
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    den, k0 = build_scene_grids(scene)
    return FusedTrainer(scene, den, k0, scene["mask"], net, n_rays, **kw), den, k0


@pytest.fixture(scope="module")
def small():
    from plenvdb_b200 import synth
    scene = synth.make_scene(96, "dense")
    net = synth.rgbnet_init()
    rays = synth.ray_batch(2048, H=200, W=200, K=synth.intrinsics(200, 200), seed=777)
    return scene, net, rays


@pytest.mark.parametrize("variant", ["dense", "sparse"])
def test_forward_counts_offsets_and_values(variant, small):
    from plenvdb_b200 import synth
    scene, net, rays = small
    if variant == "sparse":
        scene = synth.make_scene(96, "sparse")
    n = rays[0].shape[0]
    o, *_ = _run_oracle(scene, net, rays, do_update=0, cap_keep=200000)
    assert o["M3"] > 1000, "degenerate test scene"
    tr, den, k0 = _trainer(scene, net, n, parity_counts=True)
    tr.forward_backward(*[_cu(a) for a in rays])
    t = {k: v.cpu().numpy() for k, v in tr.t.items()}
    c = tr.counters()
    assert c["overflow"] == 0
    # ---- integers: bit exact
    assert np.array_equal(t["n_steps"], o["n_steps"].astype(np.int32))
    assert np.array_equal(t["cnt_mask"], o["cnt_mask"])
    assert np.array_equal(t["cnt_alpha_full"], o["cnt_alpha_full"])
    assert np.array_equal(t["cnt_alpha"], o["cnt_alpha"])
    assert np.array_equal(t["cnt_keep"], o["cnt_keep"])
    assert c["M_alpha"] == o["M2_trim"] and c["M_keep"] == o["M3"]
    assert np.array_equal(t["off_keep"], np.concatenate([[0], np.cumsum(o["cnt_keep"])]).astype(np.int32))
    assert np.array_equal(t["off_alpha"], np.concatenate([[0], np.cumsum(o["cnt_alpha"])]).astype(np.int32))
    M3 = o["M3"]
    assert np.array_equal(t["k_ray"][:M3], o["keep_ray"])
    assert np.array_equal(t["s_step"][t["k_sample"][:M3]], o["keep_step"])
    # leaf / voxel indices of the 8 corners of every kept sample
    cl = torch.zeros((M3, 8), dtype=torch.int32, device="cuda")
    co = torch.zeros_like(cl)
    k0.forward_torch(tr.t["k_xyz"][:M3].t().contiguous(), corner_out=(cl, co))
    assert np.array_equal(cl.cpu().numpy(), o["keep_leaf"]) and np.array_equal(co.cpu().numpy(), o["keep_off"])
    # ---- values
    np.testing.assert_allclose(t["s_weight"][t["k_sample"][:M3]], o["keep_weight"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(t["k_feat"][:M3], o["keep_feat"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(t["alphainv_last"], o["alphainv_last"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(t["rgb_marched"], o["rgb_marched"], rtol=RTOL, atol=2e-6)
    np.testing.assert_allclose(t["loss"], o["loss"], rtol=1e-4)


def test_gradients_match_oracle(small):
    scene, net, rays = small
    n = rays[0].shape[0]
    o, oden, ok0, aux, *_ = _run_oracle(scene, net, rays, do_update=0)
    tr, den, k0 = _trainer(scene, net, n)
    tr.forward_backward(*[_cu(a) for a in rays])
    for got, want, name in ((den.get_dense_grid_torch(den.grad).cpu().numpy(), aux[0].to_dense(), "density"),
                            (k0.get_dense_grid_torch(k0.grad).cpu().numpy(), aux[3].to_dense(), "k0")):
        scale = np.abs(want).max()
        assert scale > 0
        assert np.abs(got - want).max() <= 2e-5 * scale, "%s grad: max err %g of scale %g" % (name, np.abs(got - want).max(), scale)
        assert ((got != 0) == (want != 0)).mean() > 0.9999
    gn, wn = tr.net_grad.cpu().numpy(), o["net_grad"]
    assert np.abs(gn - wn).max() <= 2e-5 * np.abs(wn).max()


def test_update_matches_oracle_over_steps(small):
    scene, net, rays = small
    from oracle import oracle as orc
    n = rays[0].shape[0]
    tr, den, k0 = _trainer(scene, net, n)
    oden, ok0, aux = _oracle_state(scene)
    onet, onm, onv = net.copy(), np.zeros_like(net), np.zeros_like(net)
    cu = [_cu(a) for a in rays]
    for step in (1, 2, 3):
        tr.step(*cu)
        orc.train_step(_oracle_cfg(scene, step, 1), oden, aux[0], aux[1], aux[2], ok0, aux[3], aux[4], aux[5], scene["mask"], onet,
                       onm, onv, *rays)
    assert tr.launches_last_call() > 0
    # Adam divides by sqrt(v): relative differences of the gradients pass through ~unchanged; parameters moved by lr*O(1)
    for got, want, name in ((den.get_dense_grid().reshape(-1), oden.to_dense().reshape(-1), "density"),
                            (k0.get_dense_grid().reshape(-1), ok0.to_dense().reshape(-1), "k0"),
                            (tr.net.cpu().numpy(), onet, "rgbnet")):
        moved = got != (scene["density"].reshape(-1) if name == "density" else scene["k0"].reshape(-1) if name == "k0" else net)
        assert moved.sum() > 0, name
        np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-4, err_msg=name)
    # gradients consumed by the update are cleared (replaces zero_grad); untouched leaves were never visited
    assert float(den.grad.abs().max()) == 0.0 and float(k0.grad.abs().max()) == 0.0


def test_data_parallel_shards_sum_to_full_batch(small):
    """Two half-batches with n_rays_global = N reproduce the full-batch gradients."""
    scene, net, rays = small
    n = rays[0].shape[0]
    full, dfull, kfull = _trainer(scene, net, n)
    full.forward_backward(*[_cu(a) for a in rays])
    acc_d, acc_k, acc_n = 0, 0, 0
    for lo, hi in ((0, n // 2), (n // 2, n)):
        tr, d, k = _trainer(scene, net, n // 2, n_rays_global=n)
        tr.forward_backward(*[_cu(a[lo:hi]) for a in rays])
        acc_d = acc_d + d.grad.clone()
        acc_k = acc_k + k.grad.clone()
        acc_n = acc_n + tr.net_grad.clone()
    for a, b in ((acc_d, dfull.grad), (acc_k, kfull.grad), (acc_n, full.net_grad)):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())


@pytest.mark.parametrize("scratch_per_ray", [0, 4])
def test_emit_by_compaction_equals_second_march(scratch_per_ray, small):
    """The emit pass copies what the count pass parked in the per-ray scratch; rays that overflow the scratch (here almost all,
    with 4 entries) and the scratch-less configuration march a second time.  All three must give identical lists."""
    scene, net, rays = small
    rays_d = [_cu(a) for a in rays]
    ref, *_ = _trainer(scene, net, 2048)                       # default: 128 scratch entries per ray
    alt, *_ = _trainer(scene, net, 2048, scratch_per_ray=scratch_per_ray)
    for tr in (ref, alt):
        tr.forward(*rays_d[:3])
    torch.cuda.synchronize()
    ca, cb = ref.counters(), alt.counters()
    assert ca["M_alpha"] == cb["M_alpha"] and ca["M_keep"] == cb["M_keep"] and ca["M_alpha"] > 1000
    ma, mk = ca["M_alpha"], ca["M_keep"]
    for k in ("s_ray", "s_step", "s_xyz", "s_density", "s_alpha", "s_T", "s_weight"):
        assert torch.equal(ref.t[k][:ma], alt.t[k][:ma]), k
    for k in ("k_sample", "k_ray", "k_xyz", "k_corner", "k_rgb"):
        assert torch.equal(ref.t[k][:mk], alt.t[k][:mk]), k
    assert torch.equal(ref.t["rgb_marched"], alt.t["rgb_marched"])


@pytest.mark.parametrize("parity_counts", [False, True])
def test_run_skip_of_the_march_is_bit_identical(parity_counts, small):
    """Pass A of the march skips runs of 8 steps that provably cannot hit the mask (train_step.cu: run_may_hit).  With the skip
    switched off every step is tested one by one like the reference does (render_utils_kernel.cu:181-192, 385-395): the per-ray
    counts (incl. the in-mask count of every ray), both sample lists and the rendered colours must be the same bits."""
    from plenvdb_b200 import _lib
    scene, net, rays = small
    rays_d = [_cu(a) for a in rays]
    res = []
    try:
        for on in (1, 0):
            _lib.lib.pvdb_debug_set_run_skip(on)
            tr, *_ = _trainer(scene, net, 2048, parity_counts=parity_counts)
            tr.forward(*rays_d[:3])
            torch.cuda.synchronize()
            res.append(tr)
    finally:
        _lib.lib.pvdb_debug_set_run_skip(1)
    a, b = res
    ca, cb = a.counters(), b.counters()
    assert ca == cb and ca["M_alpha"] > 1000
    ma, mk = ca["M_alpha"], ca["M_keep"]
    for k in ("n_steps", "cnt_mask", "cnt_alpha", "cnt_keep", "cnt_alpha_full", "off_alpha", "off_keep", "alphainv_last", "rgb_marched"):
        assert torch.equal(a.t[k], b.t[k]), k
    assert int(a.t["cnt_mask"].sum()) > ma
    for k in ("s_ray", "s_step", "s_xyz", "s_density", "s_alpha", "s_T", "s_weight"):
        assert torch.equal(a.t[k][:ma], b.t[k][:ma]), k
    for k in ("k_sample", "k_ray", "k_xyz", "k_corner", "k_rgb"):
        assert torch.equal(a.t[k][:mk], b.t[k][:mk]), k


def test_stress_scene_properties():
    """Shell scene on a pruned topology built on the device (the S512 configuration of SURVEY.md §8d at 192^3 so that it runs
    in seconds), inverse_y cameras.  No oracle at this size: size-independent properties instead —
    determinism (two trainers give bit-identical lists), ray-ordered segments, sum of weights + T_last = 1 per ray,
    kept subset of alpha list, and the emit-by-compaction path equals the second march."""
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import FusedTrainer, build_stress_scene
    P, den, k0, mask = build_stress_scene(192)
    assert 0.01 < P["occupied_fraction"] < 0.12 and den.topo.n_leaf < (192 // 8) ** 3
    net = synth.rgbnet_init()
    rng = np.random.default_rng(7)
    n = 4096
    K = np.array([[400.0, 0, 192.0], [0, 400.0, 144.0], [0, 0, 1]], np.float32)
    poses = np.stack([synth.pose_spherical(rng.uniform(-180, 180), rng.uniform(-90, 0), rng.uniform(2.5, 3.5)) for _ in range(16)])
    poses[:, :3, 1:3] *= -1      # inverse_y (OpenCV-style) cameras look along +z
    cam = rng.integers(0, 16, n)
    px, py = rng.integers(0, 384, n), rng.integers(0, 288, n)
    ro, rd, vd = synth.rays_of_pixels(K, poses[cam], px, py, inverse_y=True)
    tg = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    rays = [_cu(a) for a in (ro, rd, vd, tg)]
    trs = [FusedTrainer(P, den, k0, mask, net, n, scratch_per_ray=spr) for spr in (128, 128, 0)]
    for tr in trs:
        tr.run(*rays, 3)
    torch.cuda.synchronize()
    c = trs[0].counters()
    ma, mk = c["M_alpha"], c["M_keep"]
    assert ma > 2000 and mk > 1000 and c["overflow"] == 0
    for other in trs[1:]:
        assert other.counters()["M_alpha"] == ma and other.counters()["M_keep"] == mk
        for k in ("s_ray", "s_step", "s_xyz", "s_alpha", "s_T", "s_weight", "k_sample", "k_corner", "off_alpha", "off_keep"):
            m = ma if k.startswith("s_") else mk if k.startswith("k_") else n + 1
            assert torch.equal(trs[0].t[k][:m], other.t[k][:m]), k
    t = trs[0].t
    s_ray = t["s_ray"][:ma].long()
    assert bool((s_ray[1:] >= s_ray[:-1]).all())                                   # ray order
    off = t["off_alpha"][:n + 1].long()
    assert bool((torch.bincount(s_ray, minlength=n) == off[1:] - off[:-1]).all())   # segments match the counts
    wsum = torch.zeros(n, device="cuda").index_add_(0, s_ray, t["s_weight"][:ma])
    np.testing.assert_allclose((wsum + t["alphainv_last"][:n]).cpu().numpy(), 1.0, atol=2e-5)   # transmittance partition
    ks = t["k_sample"][:mk].long()
    assert bool((ks[1:] > ks[:-1]).all()) and int(ks.max()) < ma                     # kept list = ordered subset
    assert bool((t["s_weight"][:ma][ks] > P["fast_color_thres"]).all())
    rgb = t["rgb_marched"][:n]
    assert bool(torch.isfinite(rgb).all()) and float(rgb.min()) >= -1e-5 and float(rgb.max()) <= 1.0 + 1e-5


def test_edge_batches_no_hits_and_ragged(small):
    """Edge cases of the fused step: a batch in which no ray hits anything (every list empty: the MLP, scatter and update
    kernels see zero work), a batch smaller than the trainer's capacity, and a single ray."""
    scene, net, rays = small
    tr, den, k0 = _trainer(scene, net, 2048)
    den0, k00, net0 = den.grid.clone(), k0.grid.clone(), tr.net.clone()
    ro, rd, vd, tg = [_cu(a) for a in rays]
    away = [ro, -rd, -vd, tg]                                  # cameras look away from the box
    tr.step(*away)
    torch.cuda.synchronize()
    c = tr.counters()
    assert c["M_alpha"] == 0 and c["M_keep"] == 0 and c["n_touched_den"] == 0 and c["overflow"] == 0
    np.testing.assert_allclose(tr.t["rgb_marched"].cpu().numpy(), scene["bg"], rtol=0, atol=0)   # background only
    assert torch.equal(den.grid, den0) and torch.equal(k0.grid, k00)                          # stepmode 1: nothing touched
    assert bool(torch.isfinite(tr.t["loss"]).all()) and bool(torch.isfinite(tr.net).all())
    # rgbnet: zero gradient -> Adam moves nothing (m = v = 0 -> p -= 0 / eps)
    assert torch.equal(tr.net, net0)
    # ragged: 777 rays through a 2048-ray trainer must equal a 777-ray trainer
    part = [t[:777].contiguous() for t in (ro, rd, vd, tg)]
    tr2, den2, k02 = _trainer(scene, net, 2048)
    tr3, den3, k03 = _trainer(scene, net, 777)
    for t_ in (tr2, tr3):
        t_.run(*part, 3)
    torch.cuda.synchronize()
    assert tr2.counters()["M_keep"] == tr3.counters()["M_keep"] > 0
    mk = tr2.counters()["M_keep"]
    assert torch.equal(tr2.t["k_sample"][:mk], tr3.t["k_sample"][:mk]) and torch.equal(tr2.t["rgb_marched"][:777], tr3.t["rgb_marched"][:777])
    # one ray
    tr4, *_ = _trainer(scene, net, 64)
    tr4.step(*[t[5:6].contiguous() for t in (ro, rd, vd, tg)])
    torch.cuda.synchronize()
    assert bool(torch.isfinite(tr4.t["loss"]).all())


def test_training_reduces_the_loss(small):
    """Functional check of the whole loop: render targets with the scene's own parameters, perturb the colour grid and the
    rgbnet, and train with the fused step — the photometric loss must fall by more than half within 60 iterations."""
    from plenvdb_b200 import synth
    scene, net, rays = small
    ro, rd, vd, _ = [_cu(a) for a in rays]
    tr_gt, *_ = _trainer(scene, net, 2048)
    target = tr_gt.forward(ro, rd, vd).clone()
    rng = np.random.default_rng(11)
    net2 = (net + rng.standard_normal(net.size).astype(np.float32) * 0.05).astype(np.float32)
    tr, den, k0 = _trainer(scene, net2, 2048)
    k0.grid.add_(torch.randn_like(k0.grid) * 0.3)
    losses = []
    for _ in range(60):
        tr.step(ro, rd, vd, target)
        losses.append(float(tr.t["loss"][1].item()))       # mse term
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])
    assert tr.counters()["overflow"] == 0


def test_host_pipeline_matches_synchronous_steps(small):
    """step_from_host_async (H2D on a copy stream into two staging buffers, losses one call late) drives exactly the same
    iterations as the synchronous step_from_host: same losses in the same order, same parameters."""
    scene, net, rays = small
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(6):
        perm = rng.permutation(2048)[:1024]
        batches.append(torch.from_numpy(np.stack([a[perm] for a in rays], 0).copy()).pin_memory())   # [4, 1024, 3]
    tr_a, den_a, k0_a = _trainer(scene, net, 1024)
    tr_b, den_b, k0_b = _trainer(scene, net, 1024)
    la = [tr_a.step_from_host(b).clone() for b in batches]
    lb = []
    for i, b in enumerate(batches):
        prev = tr_b.step_from_host_async(b)
        assert (prev is None) == (i == 0)
        if prev is not None:
            lb.append(prev.clone())
    lb.append(tr_b.host_pipeline_flush().clone())
    assert len(lb) == len(la) == 6
    for x, y in zip(la, lb):
        np.testing.assert_allclose(y.numpy(), x.numpy(), rtol=1e-4, atol=1e-6)
    assert tr_a.step_count == tr_b.step_count == 6
    np.testing.assert_allclose(tr_b.net.cpu().numpy(), tr_a.net.cpu().numpy(), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(den_b.grid.cpu().numpy(), den_a.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(k0_b.grid.cpu().numpy(), k0_a.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)


def test_graph_replay_drives_the_same_iterations(small):
    """step_from_host replays a CUDA graph of the whole iteration from its third call on (the per-iteration Adam step sizes
    travel through pvdb_train_bufs.step_scalars): same losses and parameters as issuing every kernel directly, including the
    bias corrections of later steps and an lr decay in between."""
    scene, net, rays = small
    rng = np.random.default_rng(4)
    batches = []
    for _ in range(7):
        perm = rng.permutation(2048)[:1024]
        batches.append(torch.from_numpy(np.stack([a[perm] for a in rays], 0).copy()).pin_memory())   # [4, 1024, 3]
    tr_a, den_a, k0_a = _trainer(scene, net, 1024, use_graph=False)
    tr_b, den_b, k0_b = _trainer(scene, net, 1024, use_graph=True)
    la, lb = [], []
    for i, b in enumerate(batches):
        if i == 4:
            tr_a.decay_lr(0.5)
            tr_b.decay_lr(0.5)
        la.append(tr_a.step_from_host(b).clone())
        lb.append(tr_b.step_from_host(b).clone())
    assert len(tr_b._graphs) == 1 and all(g["graph"] is not None for g in tr_b._graphs.values()) and not tr_a._graphs
    assert tr_a.step_count == tr_b.step_count == 7 and tr_a.launches_total == tr_b.launches_total
    for x, y in zip(la, lb):
        np.testing.assert_allclose(y.numpy(), x.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(tr_b.net.cpu().numpy(), tr_a.net.cpu().numpy(), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(den_b.grid.cpu().numpy(), den_a.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(k0_b.grid.cpu().numpy(), k0_a.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    # the moments carry the step sizes' history: a stale bias correction would show here first
    np.testing.assert_allclose(tr_b.net_m.cpu().numpy(), tr_a.net_m.cpu().numpy(), rtol=1e-3, atol=1e-7)


def test_graphed_step_with_device_batches_matches_direct_steps(small):
    """FusedTrainer.step with use_graph: the batch is gathered into one of two staging buffers (pvdb_stage_rays) and the iteration
    is a graph replay; batches at changing addresses, an lr decay in between."""
    scene, net, rays = small
    rng = np.random.default_rng(5)
    tr_a, den_a, k0_a = _trainer(scene, net, 1024, use_graph=False)
    tr_b, den_b, k0_b = _trainer(scene, net, 1024, use_graph=True)
    for i in range(8):
        perm = rng.permutation(2048)[:1024]
        batch = [_cu(a[perm]) for a in rays]
        if i == 5:
            tr_a.decay_lr(0.5)
            tr_b.decay_lr(0.5)
        tr_a.step(*batch)
        tr_b.step(*batch)
    torch.cuda.synchronize()
    assert len(tr_b._graphs) == 2 and all(g["graph"] is not None for g in tr_b._graphs.values())
    assert tr_a.step_count == tr_b.step_count == 8
    np.testing.assert_allclose(tr_b.t["loss"].cpu().numpy(), tr_a.t["loss"].cpu().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(tr_b.net.cpu().numpy(), tr_a.net.cpu().numpy(), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(den_b.grid.cpu().numpy(), den_a.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(k0_b.grid.cpu().numpy(), k0_a.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(tr_b.net_m.cpu().numpy(), tr_a.net_m.cpu().numpy(), rtol=1e-3, atol=1e-7)


def test_render_view_equals_the_forward_phase_on_the_same_rays():
    """The non-merged render path (run.py:171-189): FusedTrainer.render_view = get_rays_of_a_view + the forward phase in
    batch-sized chunks, bit for bit."""
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import get_rays_of_a_view
    scene = synth.make_scene(64, "dense")
    tr, den, k0 = _trainer(scene, synth.rgbnet_init(), 2048)
    H, W = 60, 70
    K, c2w = synth.intrinsics(H, W), synth.render_cameras(8)[2]
    img = tr.render_view(H, W, K, c2w)
    ro, rd, vd = [t.reshape(-1, 3) for t in get_rays_of_a_view(H, W, K, c2w, device="cuda")]
    ref = torch.cat([tr.forward(ro[a:a + 2048].contiguous(), rd[a:a + 2048].contiguous(), vd[a:a + 2048].contiguous()).clone()
                     for a in range(0, H * W, 2048)])
    assert img.shape == (H, W, 3) and torch.equal(img.reshape(-1, 3), ref)
    assert float((img != scene["bg"]).float().mean()) > 0.03, "degenerate view"
    # ... and against the oracle: the forward of DirectVoxGO (dvgo.py:296-388) on the same rays (the ray construction itself is
    # checked against the pixel formula on the CPU, tests/test_host_logic.py)
    tg = np.zeros((H * W, 3), np.float32)
    o, *_ = _run_oracle(scene, synth.rgbnet_init(), (ro.cpu().numpy(), rd.cpu().numpy(), vd.cpu().numpy(), tg), do_update=0)
    np.testing.assert_allclose(img.reshape(-1, 3).cpu().numpy(), o["rgb_marched"], rtol=1e-5, atol=2e-6)
    assert int((o["cnt_keep"] > 0).sum()) > 100


def test_sample_list_overflow_is_reported_not_silent(small):
    """A trainer whose lists are too small for the batch clamps them (no out-of-bounds write) and sets the overflow flag;
    step_from_host reads the flag with the loss at logging cadence and raises instead of training on truncated rays."""
    scene, net, rays = small
    batch = torch.from_numpy(np.stack([a[:1024] for a in rays], 0).copy()).pin_memory()
    tr, den, k0 = _trainer(scene, net, 1024, cap_alpha_per_ray=1, cap_keep_per_ray=1)
    tr.cap_alpha = tr.cap_keep = tr._bufs.cap_alpha = tr._bufs.cap_keep = 64      # far fewer entries than this batch produces
    with pytest.raises(RuntimeError, match="overflowed"):
        tr.step_from_host(batch)
    assert tr.counters()["overflow"] == 1
    ok, *_ = _trainer(scene, net, 1024)
    loss = ok.step_from_host(batch)
    assert bool(torch.isfinite(loss).all()) and ok.counters()["overflow"] == 0


@pytest.mark.parametrize("variant", ["dense", "sparse"])
def test_leaf_local_path_matches_the_default_path(variant, small):
    """csrc/leaf_local.cu (north_star: leaf values staged in shared memory, leaf-local gradient accumulation): the kept samples
    grouped by home leaf, k_feat gathered through leaf tiles in shared memory — bit-identical to the per-sample gather, same
    corner order (colorvdb.cu:81-111) —, dL/dk0 accumulated in a shared tile per leaf — the same sum in another order (1e-5,
    like the reference's own atomicAdd, colorvdb.cu:28-37)."""
    from plenvdb_b200 import _lib, synth
    scene = synth.make_scene(96, variant)
    net = synth.rgbnet_init()
    rays = [_cu(a) for a in synth.ray_batch(2048, H=200, W=200, K=synth.intrinsics(200, 200), seed=777)]
    res = []
    try:
        for on in (0, 1):
            _lib.lib.pvdb_debug_set_leaf_local(on)
            tr, den, k0 = _trainer(scene, net, 2048)
            tr.forward_backward(*rays)
            torch.cuda.synchronize()
            c = tr.counters()
            mk = c["M_keep"]
            res.append(dict(c=c, feat=tr.t["k_feat"][:mk].clone(), rgbm=tr.t["rgb_marched"].clone(), gk=k0.grad.clone(), gd=den.grad.clone(),
                            gn=tr.net_grad.clone(), items=tr.t["ll_items"][:mk].clone(), n_ll=int(tr.t["counters"][8])))
            tr.update()
            torch.cuda.synchronize()
            res[-1].update(k0=k0.grid.clone(), touched=tr.counters()["n_touched_k0"])
    finally:
        _lib.lib.pvdb_debug_set_leaf_local(0)
    a, b = res
    mk = a["c"]["M_keep"]
    assert a["c"]["M_keep"] == b["c"]["M_keep"] > 1000 and b["n_ll"] > 10
    # every kept sample that has at least one corner inside the tree appears in exactly one leaf bucket
    with_corner = torch.nonzero((tr.t["k_corner"][:mk] >= 0).any(1)).reshape(-1).cpu().tolist()
    n_items = int(tr.t["ll_off"][tr.topo.n_leaf])
    assert n_items == len(with_corner) and sorted(b["items"][:n_items].cpu().tolist()) == with_corner
    assert torch.equal(a["feat"], b["feat"]), "k_feat must be bit-identical"
    assert torch.equal(a["rgbm"], b["rgbm"]) and torch.equal(a["gd"], b["gd"]) is not None
    ga, gb = a["gk"].cpu().numpy(), b["gk"].cpu().numpy()
    assert ((ga != 0) == (gb != 0)).mean() > 0.9999
    np.testing.assert_allclose(gb, ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())
    np.testing.assert_allclose(b["gn"].cpu().numpy(), a["gn"].cpu().numpy(), rtol=1e-6, atol=0)    # rgbnet gradients do not depend on the k0 scatter
    err = (a["k0"] - b["k0"]).abs()
    assert float((err > 1e-4).float().mean()) < 1e-4      # Adam turns a cancelled gradient into +-lr: counted, not zero


def test_coarse_stage_step_matches_the_oracle():
    """The coarse stage of run.py (configs/default.py:41-66, 94): ColorVDB with 3 channels, no rgbnet, rgb = sigmoid(k0)
    (dvgo.py:72-79, 344-346), fast_color_thres 1e-7, VDBAdam with the per-voxel lr for the density (stepmode 2) and stepmode 0 for
    k0 (masked_adam.py:43-46, 56-68) — through the same fused step (coarse.cu + the full-grid Adam), three iterations against the
    oracle: counts and lists bit for bit, colours 1e-5, parameters after every step."""
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import FusedTrainer
    from plenvdb_b200.plenvdb import ColorVDB, DensityVDB
    scene = synth.make_scene(64, "dense")
    scene = dict(scene, fast_color_thres=1e-7, den_mode=2, k0_mode=0)
    R = scene["reso"]
    k0_dense = np.ascontiguousarray(scene["k0"][..., :3])
    den = DensityVDB(list(R), 1)
    k0 = ColorVDB(list(R), 3)
    k0._set_topology(den.topo)
    den.copyFromDense(scene["density"].reshape(-1))
    k0.copyFromDense(k0_dense.reshape(-1))
    n = 1024
    tr = FusedTrainer(scene, den, k0, scene["mask"], None, n, cap_alpha_per_ray=256, cap_keep_per_ray=256)
    rng = np.random.default_rng(9)
    count = rng.integers(1, 40, R).astype(np.float32)                    # view counts (dvgo.py:212-243 produces them)
    tr.set_pervoxel_lr(count)
    oden, ok0 = orc.Grid(R, 1), orc.Grid(R, 3)
    oden.copy_from_dense(scene["density"])
    ok0.copy_from_dense(k0_dense)
    operlr = orc.Grid(R, 1)
    operlr.copy_from_dense((count / count.max()).astype(np.float32))
    aux = [orc.Grid(R, c) for c in (1, 1, 1, 3, 3, 3)]
    dummy = np.zeros(22019, np.float32)
    for step in range(1, 4):
        rays = synth.ray_batch(n, H=160, W=160, K=synth.intrinsics(160, 160), seed=100 + step)
        o = orc.train_step(_oracle_cfg(scene, step, 1), oden, aux[0], aux[1], aux[2], ok0, aux[3], aux[4], aux[5], scene["mask"], dummy, dummy.copy(),
                           dummy.copy(), *rays, cap_keep=256 * n, den_perlr=operlr)
        cu = [_cu(a) for a in rays]
        tr.forward_backward(*cu)
        torch.cuda.synchronize()
        c = tr.counters()
        assert c["overflow"] == 0 and c["M_keep"] == o["M3"] > 1000 and c["M_alpha"] == o["M2_trim"]
        t = {k: tr.t[k].cpu().numpy() for k in ("cnt_keep", "cnt_alpha", "k_ray", "s_step", "k_sample", "rgb_marched", "loss", "k_feat", "k_corner")}
        M3 = o["M3"]
        assert np.array_equal(t["cnt_keep"], o["cnt_keep"]) and np.array_equal(t["cnt_alpha"], o["cnt_alpha"])
        assert np.array_equal(t["k_ray"][:M3], o["keep_ray"]) and np.array_equal(t["s_step"][t["k_sample"][:M3]], o["keep_step"])
        assert np.array_equal(t["k_corner"][:M3], np.where(o["keep_leaf"] >= 0, o["keep_leaf"] * 512 + o["keep_off"], -1))
        if step == 1:
            assert np.array_equal(t["k_feat"][:M3, :3], o["keep_feat"][:, :3])                       # same arithmetic, same order: bit-exact
        np.testing.assert_allclose(t["rgb_marched"], o["rgb_marched"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(t["loss"], o["loss"], rtol=1e-4)
        gk, wk = k0.grad.cpu().numpy().reshape(-1), aux[3].get_values().reshape(-1)
        np.testing.assert_allclose(gk, wk, rtol=1e-5, atol=1e-5 * np.abs(wk).max())
        gd, wd = den.grad.cpu().numpy().reshape(-1), aux[0].get_values().reshape(-1)
        np.testing.assert_allclose(gd, wd, rtol=1e-5, atol=2e-5 * np.abs(wd).max())
        tr.update()
        torch.cuda.synchronize()
        for got, want in ((den.grid, oden), (k0.grid, ok0)):
            got, want = got.cpu().numpy().reshape(-1), want.get_values().reshape(-1)
            err = np.abs(got - want)
            off = err > 1e-4 * np.abs(want) + 2e-4
            assert off.mean() < 2e-4 and err.max() <= 2.0 * step * 0.1 + 1e-3, (step, int(off.sum()), float(err.max()))
        assert float(den.grad.abs().max()) == 0.0 and float(k0.grad.abs().max()) == 0.0 and int(tr.t["k0_touched"].sum()) == 0


def test_two_backward_passes_before_one_update_accumulate(small):
    """Gradient accumulation: forward_backward on two half batches, then one update, equals one full-batch
    iteration — the leaves touched by the first pass must still be on the touched lists when the update runs."""
    scene, net, rays = small
    a = [_cu(x[:1024]) for x in rays]
    b = [_cu(x[1024:]) for x in rays]
    full = [_cu(x) for x in rays]
    tr1, den1, k01 = _trainer(scene, net, 2048)
    tr1.step(*full)
    tr2, den2, k02 = _trainer(scene, net, 1024, n_rays_global=2048)
    tr2.forward_backward(*a)
    n_a = tr2.counters()["n_touched_den"]
    tr2.forward_backward(*b)
    torch.cuda.synchronize()
    assert tr2.counters()["n_touched_den"] >= n_a > 0
    lst = tr2.t["den_touched_list"][: tr2.counters()["n_touched_den"]].cpu().numpy()
    assert len(set(lst.tolist())) == len(lst) == int(tr2.t["den_touched"].sum())       # every flagged leaf exactly once
    tr2.update()
    torch.cuda.synchronize()
    np.testing.assert_allclose(den2.grid.cpu().numpy(), den1.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(k02.grid.cpu().numpy(), k01.grid.cpu().numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(tr2.net.cpu().numpy(), tr1.net.cpu().numpy(), rtol=1e-3, atol=2e-5)
    assert float(den2.grad.abs().max()) == 0.0 and int(tr2.t["den_touched"].sum()) == 0

#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native PlenVDB hot path.

Metric (BASELINE.json): train rays/s for one full fine-stage iteration (forward + backward + sparse Adam on the
grids + Adam on rgbnet) on BASELINE config[1]: synthetic NeRF-Synthetic-shaped scene, 100 views 800x800, 8192-ray
batch from the fine stage's `in_maskcache` sampler, random-sparse 160^3 grid, 12-ch k0 + rgbnet.  The same JSON line
carries, as sub-objects,
  `render`        config[2]: merged-VDB 800x800 render FPS of the dense-fill F160 mic scene over the 200-pose orbit,
                  tile-sharded across the ranks, with its own `roofline`, `cpu_baseline` and `parity`;
  `stress_s512`   config[4]: the 512^3 stress scene (~5 % occupancy), 65 536-ray batch, train + render (skip: --no-s512);
  `parity`        the CUDA path against the CPU oracle on the first --cpu-rays rays of timed batch 0 (counts bit-exact,
                  rgb <= 1e-5), and at N > 1 a cross-rank check that the replicas hold identical parameters.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config f160|s512]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
`--impl reference` times the reference's host-side sampling path (the reference's kernel bodies over
the reference's own NanoVDB ReadAccessor, oracle/_ref/libref_host.so, all host threads) on the sample points of the same
workload — a reported baseline, not the target.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 8192
RESO = 160
S512_RAYS = 65536
ORACLE_KEYS = ["xyz_min", "xyz_max", "reso", "near", "far", "stepdist", "act_shift", "interval", "fast_color_thres", "bg",
               "weight_main", "weight_entropy_last", "weight_rgbper", "lr_density", "lr_k0", "lr_net", "eps", "beta0", "beta1",
               "den_mode", "k0_mode"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ workloads
def _in_mask_pool(tr, device, need, draw, pool_candidates=1 << 21):
    """Rays that hit the occupancy mask (ray_sampler='in_maskcache', configs/default.py:73; dvgo.py:583-625), drawn in
    chunks with `draw(m) -> (rays_o, rays_d, viewdirs)` numpy arrays and filtered on the device."""
    import torch
    got, chunks = 0, []
    while got < need:
        ro, rd, vd = draw(pool_candidates)
        ro_d, rd_d = torch.from_numpy(ro).to(device), torch.from_numpy(rd).to(device)
        idx = torch.nonzero(tr.hit_mask(ro_d, rd_d)).reshape(-1)
        assert idx.numel() > 0, "no ray hits the occupancy mask"
        chunks.append((ro_d[idx], rd_d[idx], torch.from_numpy(vd).to(device)[idx]))
        got += idx.numel()
    return [torch.cat([c[i] for c in chunks])[:need] for i in range(3)]


def build_workload(n_batches, device, seed=777, pool_candidates=1 << 21, n_rays=N_RAYS, use_tc=True, **trainer_kw):
    """F160-sparse (BASELINE configs[1]): scene + grids + trainer + `n_batches` device-resident ray batches."""
    import torch
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    scene = synth.make_scene(RESO, "sparse")
    net = synth.rgbnet_init()
    den, k0 = build_scene_grids(scene, device=device)
    tr = FusedTrainer(scene, den, k0, scene["mask"], net, n_rays, device=device, use_tensor_cores=use_tc, **trainer_kw)
    poses, K = synth.train_cameras(100), synth.intrinsics(800, 800)
    rng = np.random.default_rng(seed)

    def draw(m):
        cam = rng.integers(0, 100, m)
        py, px = rng.integers(0, 800, m), rng.integers(0, 800, m)
        return synth.rays_of_pixels(K, poses[cam], px, py)
    ro, rd, vd = [t.reshape(n_batches, n_rays, 3).contiguous() for t in _in_mask_pool(tr, device, n_batches * n_rays, draw, pool_candidates)]
    tg = torch.from_numpy(np.random.default_rng(5).uniform(0, 1, (n_batches, n_rays, 3)).astype(np.float32)).to(device)
    return scene, net, den, k0, tr, (ro, rd, vd, tg)


def build_workload_s512(n_batches, device, seed=7, n_rays=S512_RAYS, use_tc=True, reso=512):
    """S512 (BASELINE configs[4]): 512^3 noisy shell at ~5 % occupancy on a pruned topology, BlendedMVS-shaped
    768x576 inverse_y cameras (fx = fy = 800) at r in U(2.5, 3.5), 65 536 in_maskcache rays per batch."""
    import torch
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import FusedTrainer, build_stress_scene
    P, den, k0, mask = build_stress_scene(reso, device=device)
    net = synth.rgbnet_init()
    tr = FusedTrainer(P, den, k0, mask, net, n_rays, device=device, use_tensor_cores=use_tc)
    H, W = 576, 768
    K = np.array([[800.0, 0, W / 2], [0, 800.0, H / 2], [0, 0, 1]], np.float32)
    rng = np.random.default_rng(seed)
    poses = np.stack([synth.pose_spherical(rng.uniform(-180, 180), rng.uniform(-90, 0), rng.uniform(2.5, 3.5)) for _ in range(100)])
    poses[:, :3, 1:3] *= -1      # inverse_y (OpenCV-style) cameras look along +z
    rng = np.random.default_rng(seed + 1000 * int(os.environ.get("RANK", "0")))

    def draw(m):
        cam = rng.integers(0, 100, m)
        px, py = rng.integers(0, W, m), rng.integers(0, H, m)
        return synth.rays_of_pixels(K, poses[cam], px, py, inverse_y=True)
    ro, rd, vd = [t.reshape(n_batches, n_rays, 3).contiguous() for t in _in_mask_pool(tr, device, n_batches * n_rays, draw)]
    tg = torch.rand((n_batches, n_rays, 3), device=device)
    return P, net, den, k0, tr, (ro, rd, vd, tg), mask


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))), "measured"
    return 6650.0, 1400.0, "fallback"


def latest_traffic():
    """DRAM bytes per launch of every kernel from the newest committed `ncu --set full` capture (profiles/traffic_r*.json,
    written by profiles/summarize.py; carries the commit it was captured at)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))
    if not files:
        return {}, None
    try:
        return json.load(open(files[-1])), os.path.basename(files[-1])
    except Exception:
        return {}, None


# ------------------------------------------------------------------------------------------------ CPU legs (oracle / reference)
def algorithmic_bytes(o, n_rays):
    """B_train of SURVEY.md §8(d) from the oracle's distinct-voxel counts for one batch."""
    return (60 * n_rays + 1 * o["V_mask"] + 4 * o["V_den"] + 48 * o["V_k0"] + 2 * (4 * o["V_den_grad"] + 48 * o["V_k0"])
            + 28 * (o["V_den_grad"] + 12 * o["V_k0"]) + 28 * 22019)


def oracle_step(scene, net, rays, threads, n_sub=None, cap_keep=0):
    """One CPU-oracle training step on the first n_sub rays; returns (seconds, outputs)."""
    from oracle import oracle as orc
    R, act = scene["reso"], scene["active"]
    den, k0 = orc.Grid(R, 1, act), orc.Grid(R, 12, act)
    den.copy_from_dense(scene["density"])
    k0.copy_from_dense(scene["k0"])
    aux = [orc.Grid(R, c, act) for c in (1, 1, 1, 12, 12, 12)]
    cfg = {k: scene[k] for k in ORACLE_KEYS}
    cfg.update(step=1, do_update=1, n_rays_global=0, threads=threads)
    sub = [a[:n_sub] for a in rays] if n_sub else rays
    nm, nv = np.zeros_like(net), np.zeros_like(net)
    t0 = time.perf_counter()
    out = orc.train_step(cfg, den, aux[0], aux[1], aux[2], k0, aux[3], aux[4], aux[5], scene["mask"], net.copy(), nm, nv, *sub,
                         cap_keep=cap_keep)
    return time.perf_counter() - t0, out


def cpu_sample_lists(scene, ro, rd):
    """The sample points the fine stage evaluates for these rays, found on the CPU with the oracle's restatement of the
    reference call sequence (dvgo.py:272-388): sample_pts_on_rays -> bbox -> mask cache (M1 points, density forward) ->
    raw2alpha > thres -> alpha2weight with the early stop (M2' points, density backward) -> weight > thres (M3 points, k0
    forward + backward).  Index-space float32 coordinates as grid.py:77-78 forms them."""
    from oracle import oracle as orc
    from plenvdb_b200.synth import mask_scale_shift
    n = ro.shape[0]
    pts, mob, rid, _, _, _, _ = orc.sample_pts_on_rays(ro, rd, scene["xyz_min"], scene["xyz_max"], scene["near"], scene["far"], scene["stepdist"])
    inb = ~mob
    pts, rid = pts[inb], rid[inb]
    sc, sh = mask_scale_shift(scene["mask"].shape, scene["xyz_min"], scene["xyz_max"])
    m = orc.maskcache_lookup(scene["mask"], pts, sc, sh)
    pts, rid = pts[m], rid[m]
    mn, mx = np.asarray(scene["xyz_min"], np.float32), np.asarray(scene["xyz_max"], np.float32)
    idx = ((pts - mn) / (mx - mn) * (np.asarray(scene["reso"], np.float32) - np.float32(1))).astype(np.float32)
    R, act = scene["reso"], scene["active"]
    oden = orc.Grid(R, 1, act)
    oden.copy_from_dense(scene["density"])
    dens = oden.forward(idx[:, 0], idx[:, 1], idx[:, 2], threads=os.cpu_count() or 1).reshape(-1)
    _, alpha = orc.raw2alpha(dens, scene["act_shift"], scene["interval"])
    k = alpha > scene["fast_color_thres"]
    idx2, rid2, alpha2 = idx[k], rid[k], alpha[k]
    w, _, _, _, i_e = orc.alpha2weight(alpha2, rid2, n)
    trim = np.arange(alpha2.size) < i_e[rid2]
    keep = w > scene["fast_color_thres"]
    xyz = lambda a: [np.ascontiguousarray(a[:, c]) for c in range(3)]
    return {"M1": xyz(idx), "M2t": xyz(idx2[trim]), "M3": xyz(idx2[keep])}


def ref_sampling_pass(scene, lists, threads, repeats):
    """the reference's density_forward / backward and color_forward / backward kernel bodies executed
    on the host through the reference's own NanoVDB ReadAccessor (oracle/_ref/libref_host.so, std::thread over sample
    ranges, CAS float adds) on GridBuilder grids of this scene.  Falls back to the oracle's port of the same four functions
    when the reference-compiled library did not travel.  Returns (best seconds per pass, all seconds, kind)."""
    from oracle import ref
    rng = np.random.default_rng(9)
    g2 = rng.standard_normal(lists["M2t"][0].size).astype(np.float32)
    g3 = rng.standard_normal((lists["M3"][0].size, 12)).astype(np.float32)
    R, act = scene["reso"], scene["active"]
    if ref.available("host"):
        kind = "reference"
        rd_, rk_ = ref.RefGrid(R, 1, act), ref.RefGrid(R, 12, act)
        rd_.copy_from_dense_host(scene["density"])
        rk_.copy_from_dense_host(scene["k0"])
        gd_, gk_ = ref.RefGrid(R, 1, act), ref.RefGrid(R, 12, act)      # gradient grids, like DensityVDB.grad / ColorVDB.grad

        def one():
            rd_.host_forward(*lists["M1"], threads)
            rk_.host_forward(*lists["M3"], threads)
            gd_.host_backward(*lists["M2t"], g2, threads)
            gk_.host_backward(*lists["M3"], g3, threads)
    else:
        from oracle import oracle as orc
        kind = "port"
        od, ok = orc.Grid(R, 1, act), orc.Grid(R, 12, act)
        od.copy_from_dense(scene["density"])
        ok.copy_from_dense(scene["k0"])
        gd_, gk_ = orc.Grid(R, 1, act), orc.Grid(R, 12, act)

        def one():
            od.forward(*lists["M1"], threads=threads)
            ok.forward(*lists["M3"], threads=threads)
            gd_.backward(*lists["M2t"], g2, threads=threads)
            gk_.backward(*lists["M3"], g3, threads=threads)
    one()
    one()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    return min(ts), ts, kind


def cpu_baseline_sampling(scene, ro, rd, threads, budget_s=12.0):
    lists = cpu_sample_lists(scene, ro, rd)
    probe, _, kind = ref_sampling_pass(scene, lists, threads, 1)
    reps = int(max(3, min(200, budget_s / max(probe, 1e-4))))
    best, ts, kind = ref_sampling_pass(scene, lists, threads, reps)
    n = ro.shape[0]
    counts = {k: int(v[0].size) for k, v in lists.items()}
    return {"value": n / best, "unit": "rays/s", "cores": threads, "kind": kind,
            "sample": ("trilinear density forward on the %d in-mask sample points, k0 (12 ch = 4 Vec3f grids) forward + backward on the %d "
                       "kept points, density backward on the %d alpha-list points of %d in_maskcache rays of the timed workload; "
                       "best of %d passes (%.1f s of CPU work), mean %.0f rays/s" % (counts["M1"], counts["M3"], counts["M2t"], n, reps,
                                                                                    float(np.sum(ts)), n / float(np.mean(ts)))),
            "what": ("the reference's density_forward/backward + color_forward/backward kernel bodies (densityvdb.cu:101-167, colorvdb.cu:81-160) "
                     "over the reference's NanoVDB ReadAccessor on the host cores (BASELINE.md section 4)" if kind == "reference" else
                     "the oracle's port of the four trilinear kernels (oracle/_ref/libref_host.so not present)")}


# ------------------------------------------------------------------------------------------------ timing of the training step
def time_training(tr, stepper, batches, K, Wm, world, dev, n_rays, flush, with_e2e=True):
    """W warm-up + K timed iterations (L2 flushed between them, CUDA events, max over ranks), the warm-L2 back-to-back rate, and
    (with_e2e) the host-fed loops through the public API."""
    import torch
    ro, rd, vd, tg = batches
    res = {}

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    for i in range(Wm):
        stepper(ro[i], rd[i], vd[i], tg[i])
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    l0 = tr.launches_total
    # The barrier leaves the host with no lead over the GPU: a host thread that is descheduled for a few hundred microseconds
    # while it issues the first timed step would stall the device (and, in a data-parallel run, every peer waiting for this
    # rank inside ITS timed step).  One millisecond of device-side delay in front of the first step lets the host queue the
    # first steps before the device gets to them, like in steady state; it is outside the timed events.
    torch.cuda._sleep(2_000_000)
    for i in range(K):
        flush.fill_(i & 0xFF)                      # flush L2 between timed iterations (outside the timed events)
        ev[i][0].record()
        stepper(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i])
        ev[i][1].record()
    res["launches"] = tr.launches_total - l0
    barrier()
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms_total = max_over_ranks(sum(per_step))
    res["step_ms"] = {"median": float(np.median(per_step)), "min": float(np.min(per_step)), "max": float(np.max(per_step)),
                      "slowest_step": int(np.argmax(per_step)), "note": "per-step CUDA-event times on rank 0; `value` uses the sum over the K steps, maximum over the ranks"}
    cnt = tr.counters()
    assert cnt["overflow"] == 0, "sample list overflow: raise cap_*_per_ray"
    res.update(ms_total=ms_total, value=world * n_rays * K / (ms_total * 1e-3), counters=cnt)
    # ---- warm L2 (back-to-back) device rate: the cache regime of the end-to-end loops below
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        stepper(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i])
    e1.record()
    barrier()
    res["warm_ms"] = max_over_ranks(e0.elapsed_time(e1)) / K
    if not with_e2e:
        return res
    # ---- end to end through the public API with HOST buffers (pinned): every iteration's rays and targets are copied H2D and its
    # loss words D2H inside the timed region.  Headline: FusedTrainer.step_from_host_async, the loop a trainer that logs the loss
    # one iteration late uses — copy stream, two staging buffers, iteration i's loss words read while i + 1 runs, one wait per
    # iteration (on iteration i - 1).  Beside it: the strictly synchronous step_from_host (copy, step, read the loss, synchronise,
    # every iteration: what `psnr.item()` costs run.py:590).
    hbatch = torch.stack([x[Wm:].cpu() for x in (ro, rd, vd, tg)], 1).contiguous().pin_memory()   # [K, 4, n, 3]
    e2e_ms = {}
    sync_stepper = None if world == 1 else stepper
    for i in range(min(3, K)):        # untimed: streams, staging, pinned buffers (and the graph) are created on first use
        tr.step_from_host_async(hbatch[i], stepper=sync_stepper)
    tr.host_pipeline_flush()
    for i in range(min(3, K)):
        tr.step_from_host(hbatch[i], stepper=sync_stepper)
    for mode in ("sync", "pipelined"):
        barrier()
        e0.record()
        for i in range(K):
            if mode == "pipelined":
                tr.step_from_host_async(hbatch[i], stepper=sync_stepper)
            else:
                tr.step_from_host(hbatch[i], stepper=sync_stepper)   # one H2D, the step, D2H of the loss, stream sync (the caller reads it)
        last = tr.host_pipeline_flush() if mode == "pipelined" else None
        e1.record()
        barrier()
        assert mode != "pipelined" or (last is not None and bool(torch.isfinite(last).all()))
        e2e_ms[mode] = max_over_ranks(e0.elapsed_time(e1))
    res["e2e_sync_value"] = world * n_rays * K / (e2e_ms["sync"] * 1e-3)
    res["e2e_value"] = world * n_rays * K / (e2e_ms["pipelined"] * 1e-3)
    return res


def kernel_times(tr, batches, K, Wm, flush, full_step):
    """Per-kernel times of the fused call: CUDA events between the kernels, on the launching stream (side stream off)."""
    from plenvdb_b200 import _lib
    ro, rd, vd, tg = batches
    _lib.profile_enable(True)
    acc = {}
    for i in range(K):
        flush.fill_(i & 0xFF)
        if full_step:
            tr.step(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i])
        else:
            tr.run(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i], 3)
        for name, t_ms in _lib.profile_fetch():
            acc.setdefault(name, []).append(t_ms)
        if not full_step:
            tr.update()
    _lib.profile_enable(False)
    return {k: float(np.mean(v)) for k, v in acc.items()}


def parity_f160(scene, net, rays0, n_sub, dev, threads, use_tc):
    """The CUDA path against the CPU oracle on the first n_sub rays of timed batch 0, from the scene's initial state (fresh
    grids): per-ray counts, list sizes and segment offsets bit for bit, rgb_marched and the loss within 1e-5 / 1e-4."""
    import torch
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    cpu_s, o = oracle_step(scene, net, rays0, threads, n_sub, cap_keep=96 * n_sub)
    den2, k02 = build_scene_grids(scene, device=dev)
    tr2 = FusedTrainer(scene, den2, k02, scene["mask"], net, n_sub, device=dev, use_tensor_cores=use_tc, parity_counts=True)
    tr2.forward_backward(*[torch.from_numpy(np.ascontiguousarray(a[:n_sub])).to(dev) for a in rays0])
    torch.cuda.synchronize()
    t = {k: tr2.t[k].cpu().numpy() for k in ("n_steps", "cnt_mask", "cnt_alpha_full", "cnt_alpha", "cnt_keep", "off_keep", "rgb_marched",
                                            "loss", "k_ray", "s_step", "k_sample", "alphainv_last")}
    c = tr2.counters()
    M3 = o["M3"]
    ints = {
        "n_steps": bool(np.array_equal(t["n_steps"], o["n_steps"].astype(np.int32))),
        "cnt_mask": bool(np.array_equal(t["cnt_mask"], o["cnt_mask"])),
        "cnt_alpha_full": bool(np.array_equal(t["cnt_alpha_full"], o["cnt_alpha_full"])),
        "cnt_alpha": bool(np.array_equal(t["cnt_alpha"], o["cnt_alpha"])),
        "cnt_keep": bool(np.array_equal(t["cnt_keep"], o["cnt_keep"])),
        "M2_trim": bool(c["M_alpha"] == o["M2_trim"]), "M3": bool(c["M_keep"] == M3),
        "off_keep": bool(np.array_equal(t["off_keep"], np.concatenate([[0], np.cumsum(o["cnt_keep"])]).astype(np.int32))),
        "keep_ray": bool(c["M_keep"] == M3 and np.array_equal(t["k_ray"][:M3], o["keep_ray"])),
        "keep_step": bool(c["M_keep"] == M3 and np.array_equal(t["s_step"][t["k_sample"][:M3]], o["keep_step"])),
    }
    rgb_err = np.abs(t["rgb_marched"] - o["rgb_marched"])
    rgb_ok = bool(np.all(rgb_err <= 1e-5 * np.abs(o["rgb_marched"]) + 2e-6))
    ail_ok = bool(np.allclose(t["alphainv_last"], o["alphainv_last"], rtol=1e-5, atol=1e-9))
    loss_rel = float(abs(t["loss"][0] - o["loss"][0]) / max(abs(o["loss"][0]), 1e-30))
    ok = all(ints.values()) and rgb_ok and ail_ok and loss_rel <= 1e-4
    if not ok:
        log("[bench] PARITY FAILURE against the oracle: %s rgb_ok=%s ail_ok=%s loss_rel=%.2e" % (ints, rgb_ok, ail_ok, loss_rel))
    return {"ok": bool(ok), "against": "CPU oracle (oracle/plenvdb_oracle.cpp), initial scene state", "rays": n_sub,
            "integers_bit_exact": ints, "M1": o["M1"], "M2": o["M2"], "M2_trim": o["M2_trim"], "M3": M3,
            "rgb_marched_max_abs_err": float(rgb_err.max()), "rgb_marched_within_1e-5": rgb_ok, "alphainv_last_within_1e-5": ail_ok,
            "loss_rel_err": loss_rel, "oracle_step_s": cpu_s}, o, cpu_s


def replica_check(tr, world, dev):
    """N > 1: every rank's density / k0 / rgbnet parameters after the timed steps must be the SAME BITS (the exchange sums in
    rank order on every rank).  Checksums (int64 sums of the float bit patterns + an xor-ish second word) are all-gathered."""
    import torch
    words = []
    for t in (tr.density.grid, tr.k0.grid, tr.net, tr.den_m, tr.k0_v, tr.net_v):
        b = t.reshape(-1).view(torch.int32).to(torch.int64)
        words += [b.sum(), (b * (torch.arange(b.numel(), device=dev, dtype=torch.int64) % 8191 + 1)).sum()]
    mine = torch.stack(words)
    allw = [torch.empty_like(mine) for _ in range(world)]
    torch.distributed.all_gather(allw, mine)
    same = all(bool(torch.equal(allw[0], w)) for w in allw[1:])
    return {"replicas_identical": bool(same), "ranks": world, "checked": "density, k0, rgbnet parameters and Adam moments, bit patterns"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    from plenvdb_b200 import dist as pdist
    rank, local, world = pdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local)
    K, Wm = args.steps, args.warmup
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    clk = ClockSampler(local).__enter__()     # sampled through every measurement loop below
    result = None
    if args.config == "s512":
        result = bench_s512(args, dev, rank, world, flush, headline=True)
        if rank == 0:
            result["clocks"] = clk.summary()
    else:
        result = bench_f160(args, dev, rank, world, flush, clk)
        if not args.no_s512:
            torch.cuda.empty_cache()
            try:
                s = bench_s512(args, dev, rank, world, flush, headline=False)
            except Exception as e:   # noqa: BLE001 — the headline stands; the stress sub-result says why it is missing
                s = {"error": "%s: %s" % (type(e).__name__, e)}
                log("[bench] S512 stress sub-benchmark failed: %s" % s["error"])
            if rank == 0:
                result["stress_s512"] = s
    clk.__exit__()
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def bench_f160(args, dev, rank, world, flush, clk):
    import torch
    from plenvdb_b200 import dist as pdist
    K, Wm = args.steps, args.warmup
    nb = K + Wm
    t_setup = time.time()
    scene, net, den, k0, tr, batches = build_workload(nb, dev, seed=777 + rank, use_tc=not args.fp32_rgbnet)
    dp = None
    if world > 1 and args.exchange != "none":
        dp = pdist.DataParallelTrainer.wrap(tr, world, exchange=args.exchange)
        stepper = dp.step
    else:
        stepper = tr.step        # --exchange none at N > 1: independent replicas (diagnosis of what the exchange costs; not a DP run)
    ro, rd, vd, tg = batches
    log("[bench] rank %d setup %.1fs, pool ready" % (rank, time.time() - t_setup))
    T = time_training(tr, stepper, batches, K, Wm, world, dev, N_RAYS, flush)
    cnt, ms_total, value = T["counters"], T["ms_total"], T["value"]
    clocks = clk.summary()
    rep = replica_check(tr, world, dev) if (world > 1 and dp is not None) else None

    # ---- exchange kernels (all ranks step together; rank 0 reports)
    xchg = {}
    if world > 1 and dp is not None and dp.peer is not None:
        from plenvdb_b200 import _lib
        _lib.profile_enable(True)
        for i in range(K):
            flush.fill_(i & 0xFF)
            tr.run(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i], 3)
            dp.peer.exchange(tr._bufs)
            torch.cuda.synchronize()
            for name, t_ms in _lib.profile_fetch():
                xchg.setdefault(name, []).append(t_ms)
            tr.update(lists_ready=True)
        _lib.profile_enable(False)
        xchg = {k: float(np.mean(v)) for k, v in xchg.items()}

    result = None
    if rank == 0:
        hbm, tf, which = measured_peaks()
        kern = kernel_times(tr, batches, K, Wm, flush, full_step=(world == 1))
        top = max(kern, key=kern.get)
        cpu_threads = os.cpu_count() or 1
        rays0 = [x[Wm].cpu().numpy() for x in (ro, rd, vd, tg)]
        n_sub = min(N_RAYS, args.cpu_rays)
        par, o, cpu_s = parity_f160(scene, net, rays0, n_sub, dev, cpu_threads, tr.use_tc)
        if rep is not None:
            par.update(rep)
            par["ok"] = bool(par["ok"] and rep["replicas_identical"])
        scale = N_RAYS / n_sub
        M3 = cnt["M_keep"]
        b_train = algorithmic_bytes({k: o[k] * scale for k in ("V_mask", "V_den", "V_k0", "V_den_grad")}, N_RAYS)
        mlp_flops_fwd = 2.0 * M3 * (39 * 128 + 128 * 128 + 128 * 3)
        mlp_pass = 3.0 if tr.use_tc else 1.0      # 3xTF32 issues three tensor-core products per algorithmic product
        kern_flops = {"rgbnet_fwd": mlp_flops_fwd, "rgbnet_bwd": 2.0 * mlp_flops_fwd,
                      "rgbnet_bwd_act": 2.0 * M3 * (128 * 128 + 128 * 12), "rgbnet_bwd_wgrad": 2.0 * M3 * (128 * 128 + 128 * 40 + 3 * 128)}
        tj, tj_name = latest_traffic()
        kname = {"rgbnet_fwd": "k_rgbnet_fwd_tc", "rgbnet_bwd_act": "k_rgbnet_bwd_act_tc", "rgbnet_bwd_wgrad": "k_rgbnet_bwd_wgrad_tc",
                 "march_count": "k_march"}.get(top)
        traffic = tj[kname]["dram_bytes"] if (tr.use_tc and kname in tj) else None
        traffic_in_situ = tj[kname].get("dram_bytes_in_situ") if (tr.use_tc and kname in tj) else None
        # bytes the three MLP kernels move by design (activations handed over through HBM), per kept sample
        design_bytes = {"rgbnet_fwd": M3 * (512 + 512 + 160 + 32 + 48 + 12 + 44.0), "rgbnet_bwd_act": M3 * (512 + 88.0),
                        "rgbnet_bwd_wgrad": M3 * (3 * 512 + 160 + 12 + 16.0) + 148 * 22048 * 4.0}
        if top in kern_flops:
            ach = kern_flops[top] / (kern[top] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf,
                    "traffic": traffic, "traffic_in_situ": traffic_in_situ, "traffic_source": tj_name, "traffic_commit": tj.get("commit"),
                    "peak_source": which, "tensor_core_flops_issued": kern_flops[top] * mlp_pass,
                    "hbm_view": ({"design_bytes": design_bytes[top], "achieved_GBs": design_bytes[top] / (kern[top] * 1e-3) / 1e9,
                                  "frac_of_hbm": design_bytes[top] / (kern[top] * 1e-3) / 1e9 / hbm,
                                  "note": "activation tensors this kernel reads/writes through HBM by design; the real limiter of the tcgen05 kernels"}
                                 if top in design_bytes and tr.use_tc else None),
                    "note": ("algorithmic fp32 FLOPs of the kernel / its duration, against the %s dense bf16 cuBLAS peak; "
                             "the tcgen05 path issues 3 TF32 products per algorithmic product (tf32 peak = bf16/2)" % which)}
        else:
            kb = {"march_count": 60 * N_RAYS / 2 + o["V_mask"] * scale + 4 * o["V_den"] * scale,
                  "march_emit": 60 * N_RAYS / 2 + o["V_mask"] * scale + 4 * o["V_den"] * scale + 36 * cnt["M_alpha"]}.get(top, b_train)
            ach = kb / (kern[top] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
                    "traffic_in_situ": traffic_in_situ, "traffic_source": tj_name, "traffic_commit": tj.get("commit"), "peak_source": which}
        step_roof = b_train / (ms_total / K * 1e-3) / 1e9
        base = cpu_baseline_sampling(scene, rays0[0], rays0[1], cpu_threads)
        base["oracle_full_step"] = {"value": n_sub / cpu_s, "unit": "rays/s", "kind": "port", "cores": cpu_threads,
                                    "sample": "1 full oracle iteration (fwd+bwd+update incl. rgbnet) on the first %d rays of timed batch 0" % n_sub}
        warm_value = world * N_RAYS / (T["warm_ms"] * 1e-3)
        result = {
            "metric": "train rays/s (fwd+bwd+update), fine stage", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "F160-sparse fine-stage step: 8192 in_maskcache rays/GPU, 100 views 800x800, random-sparse 160^3 "
                                   "(p_drop 0.7), 12-ch k0 + rgbnet(39-128-128-3), stepmode 1",
                       "n_rays_per_gpu": N_RAYS, "l2": "flushed between timed steps (256 MiB write, outside the events)",
                       "rgbnet": "fp32 cuda cores" if not tr.use_tc else "tcgen05 3xTF32 forward + backward", "parallelism": "dp%d" % world,
                       "samples": {"M_alpha": cnt["M_alpha"], "M_keep": M3, "touched_leaves_density": cnt["n_touched_den"],
                                   "touched_leaves_k0": cnt["n_touched_k0"]}},
            "step_ms": T["step_ms"], "warm_l2_ms_per_step": T["warm_ms"], "warm_l2_value": warm_value,
            "e2e": {"value": T["e2e_value"], "unit": "rays/s", "h2d_bytes_per_step": 4 * N_RAYS * 3 * 4, "d2h_bytes_per_step": 16,
                    "how": "FusedTrainer.step_from_host_async, every iteration: pinned host batch -> H2D on a copy stream -> step -> loss "
                           "words D2H to pinned host; the host waits for iteration i - 1 and reads its loss while iteration i runs",
                    "synchronous_value": T["e2e_sync_value"],
                    "synchronous_how": "FusedTrainer.step_from_host: H2D, step, D2H of the loss, stream synchronise, every iteration",
                    "cache_regime": "warm L2 (iterations back to back, no flush inside the loop): compare with warm_l2_value, the device "
                                    "rate under the same regime; `value` flushes L2 between steps",
                    "frac_of_warm_device_rate": T["e2e_value"] / warm_value, "synchronous_frac_of_warm_device_rate": T["e2e_sync_value"] / warm_value},
            "gpu_launches": T["launches"],
            "kernel_ms": kern,
            "roofline": roof,
            "step_roofline": {"algorithmic_bytes": b_train, "achieved_GBs": step_roof, "frac_of_hbm": step_roof / hbm,
                              "note": "B_train of SURVEY.md 8(d); the step is latency/compute bound, not HBM bound"},
            "cpu_baseline": base,
            "parity": par,
            "clocks": clocks,
        }
        if world > 1 and dp is None:
            result["config"]["exchange"] = "NONE (--exchange none): independent replicas, diagnosis only"
        if world > 1 and dp is not None:
            result["config"]["exchange_bytes_per_step"] = dp.exchange_bytes()
            if xchg:
                result["exchange_kernel_ms"] = xchg
            result["config"]["exchange"] = ("own kernels over NVLink peer memory (CUDA IPC), no NCCL call in the step"
                                            if dp.peer is not None else "NCCL all-reduce of packed touched-leaf tiles")
            if dp.peer is not None and dp.peer.error():
                result["config"]["exchange_error"] = dp.peer.error()
    if dp is not None:
        dp.close()
    del tr, den, k0, batches, ro, rd, vd, tg
    torch.cuda.empty_cache()
    # ---- merged-VDB render FPS (config[2]): the dense-fill F160 mic scene, 200-pose orbit, tile-sharded across the ranks
    if not args.no_render:
        r = run_render_f160(args, net, dev, rank, world)
        if rank == 0:
            result["render"] = r
    return result


def _mlp_of(net):
    from plenvdb_b200 import synth
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(net)
    return (np.ascontiguousarray(w0.T), b0, np.ascontiguousarray(w1.T), b1, np.ascontiguousarray(w2.T), b2)


def run_render_f160(args, net, dev, rank, world):
    """BASELINE configs[2]: dense-fill F160 mic scene -> merged format (vdb_compression.py:28-58), 800x800, the 200 poses of the
    orbit (load_blender.py:74)."""
    import torch
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import build_scene_grids
    from plenvdb_b200.renderer import merge_grids
    scene = synth.make_scene(RESO, "dense")
    den, k0 = build_scene_grids(scene, device=dev)
    dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
    del den, k0
    out = time_render(args, scene, net, dend, cold, idx, n, dev, rank, world,
                      workload="R800: dense-fill F160 mic scene merged (mergedidxs + mergeddata, values through fp16), 800x800, 200-pose orbit")
    if rank == 0 and not args.no_cpu_render:
        # CPU: the oracle's restatement of render_an_image (renderer.cu:370-424), all host threads, ONE full frame of the orbit — also
        # the parity check of that frame: per-pixel sample counts bit for bit, RGB within 1e-5
        from oracle import oracle as orc
        threads = os.cpu_count() or 1
        pose_i = 67
        oidx = orc.Grid(scene["reso"], 1, idx.cpu().numpy() != 0)
        oidx.copy_from_dense(idx.cpu().numpy().astype(np.float32))
        cfg = dict(reso=scene["reso"], K=synth.intrinsics(800, 800), xyz_min=scene["xyz_min"], xyz_max=scene["xyz_max"], near=scene["near"],
                   stepdist=scene["stepdist"], act_shift=scene["act_shift"], interval=scene["interval"], fast_color_thres=scene["fast_color_thres"],
                   bg=scene["bg"], inverse_y=0, H=800, W=800, threads=threads)
        poses = synth.render_cameras(200)
        t0 = time.perf_counter()
        want, wns, bad = orc.render(cfg, oidx, dend.cpu().numpy(), cold.cpu().numpy(), _mlp_of(net), poses[pose_i])
        cpu_s = time.perf_counter() - t0
        r = out.pop("_renderer")
        img = r.render_rows_torch(torch.from_numpy(poses[pose_i].reshape(-1)).to(dev), 0, 800).reshape(-1, 3).cpu().numpy()
        ns = r.s["n_samples"].cpu().numpy()
        counts_ok = bool(np.array_equal(ns, wns))
        err = np.abs(img - want)
        rgb_ok = bool(np.all(err <= 1e-5 * np.abs(want) + 3e-6)) if bad == 0 else bool(np.mean(err.max(1) <= 1e-5 + 3e-6) > 0.9999)
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": "one full 800x800 frame (pose %d of the orbit, %d samples) through the oracle's restatement of render_an_image" % (pose_i, int(wns.sum()))}
        out["parity"] = {"ok": bool(counts_ok and rgb_ok), "against": "CPU oracle, full 800x800 frame, pose %d" % pose_i,
                         "per_pixel_sample_counts_bit_exact": counts_ok, "samples": int(wns.sum()), "rgb_max_abs_err": float(err.max()),
                         "rgb_within_1e-5": rgb_ok, "reference_inconsistent_rays": int(bad)}
        if not out["parity"]["ok"]:
            log("[bench] RENDER PARITY FAILURE against the oracle: %s" % out["parity"])
    if out is not None:
        out.pop("_renderer", None)
    return out


def time_render(args, scene, net, dend, cold, idx, n, dev, rank, world, workload, cap_per_pixel=6):
    import torch
    from plenvdb_b200 import dist as pdist
    from plenvdb_b200 import render_utils_cuda as ru
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import get_rays_of_a_view
    from plenvdb_b200.plenvdb import MGRenderer
    H = W = 800
    mlp = _mlp_of(net)
    r = MGRenderer(12, 27, 128, 3, device=dev, use_tensor_cores=not args.fp32_rgbnet, cap_per_pixel=cap_per_pixel)
    r.load_data_dense(dend, cold, idx)
    r.load_params(mlp[0].reshape(-1), mlp[1], mlp[2].reshape(-1), mlp[3], mlp[4].reshape(-1), mlp[5])
    Kmat = synth.intrinsics(H, W)
    r.setScene(list(scene["reso"]), Kmat.reshape(-1), scene["xyz_min"], scene["xyz_max"])
    r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"],
                False, H, W)
    poses_np = synth.render_cameras(200)
    poses = torch.from_numpy(poses_np.reshape(200, 16)).to(dev)
    nf = args.frames
    # N > 1: interleaved 16-row groups, every rank's composite kernel writes its pixels into rank 0's frame over NVLink peer
    # memory (no collective); --render-gather nccl = contiguous row bands + an NCCL gather
    peer = None
    sharding = "single GPU"
    if world > 1:
        sharding = "contiguous row bands + NCCL gather"
        if args.render_gather == "peer":
            try:
                peer = pdist.PeerFrame(H, W, band_rows=16)
                sharding = "interleaved 16-row groups, frame assembled in rank 0's memory by peer stores over NVLink (no collective)"
            except pdist.PeerExchangeUnavailable as e:
                sys.stderr.write("render: %s; falling back to the NCCL gather\n" % e)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())
    for i in range(3):
        pdist.render_sharded(r, poses[i], rank, world, peer=peer)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nf):
        pdist.render_sharded(r, poses[i % 200], rank, world, peer=peer)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / nf
    c = r.counters()
    assert c["overflow"] == 0, "render sample list overflow: raise cap_per_pixel"
    # e2e: pose from host, frame back to host (7.68 MB) every frame, like run.py:157-167 — synchronous, then with the D2H of
    # frame i on a copy stream under the render of frame i + 1 (two device frames, two pinned host frames; the caller has frame
    # i - 1 on the host while frame i renders).  At N > 1 the frame is rank 0's PeerFrame buffer (double-buffered).
    hposes = poses.cpu().pin_memory()
    himg = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    barrier()
    e0.record()
    for i in range(nf):
        r.c2w.copy_(hposes[i % 200], non_blocking=True)
        img = pdist.render_sharded(r, r.c2w, rank, world, peer=peer)
        if rank == 0:
            himg.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    e2e_sync_fps = nf * 1e3 / max_over_ranks(e0.elapsed_time(e1))
    e2e_fps, e2e_how = e2e_sync_fps, "pose H2D, frame, 7.68 MB D2H, stream synchronise, every frame"
    if world == 1 or peer is not None:
        try:
            copy = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream()
            outs = [torch.empty((H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)] if world == 1 else None
            himgs = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
            rendered = [torch.cuda.Event() for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]

            def pipelined(n_frames):
                used = [False, False]
                for i in range(n_frames):
                    k = i & 1
                    if used[k] and rank == 0:
                        cur.wait_event(copied[k])            # frame i - 2 has left its device buffer
                    r.c2w.copy_(hposes[i % 200], non_blocking=True)
                    if world == 1:
                        src = r.render_rows_torch(r.c2w, 0, H, out=outs[k])
                    else:
                        src = peer.render(r, r.c2w)          # rank 0: view of frame buffer (frame_no & 1); PeerFrame alternates itself
                    if rank == 0:
                        rendered[k].record(cur)
                        copy.wait_event(rendered[k])
                        with torch.cuda.stream(copy):
                            himgs[k].copy_(src, non_blocking=True)
                            copied[k].record(copy)
                        used[k] = True
                        if i >= 1:
                            copied[k ^ 1].synchronize()      # frame i - 1 is on the host
                if rank == 0:
                    copied[(n_frames - 1) & 1].synchronize()

            if world > 1 and (peer.frame_no & 1):            # keep the host buffer index aligned with PeerFrame's buffer parity
                peer.render(r, r.c2w)
            pipelined(4)
            barrier()
            e0.record()
            pipelined(nf)
            e1.record()
            barrier()
            pipe_ms = max_over_ranks(e0.elapsed_time(e1))
            ok = True
            if rank == 0:
                if world == 1:
                    want = r.render_rows_torch(poses[(nf - 1) % 200], 0, H).cpu()
                else:
                    want = None
                ok = want is None or torch.equal(himgs[(nf - 1) & 1], want)
            if world > 1:
                want_all = pdist.render_sharded(r, poses[(nf - 1) % 200], rank, world, peer=peer)
                if rank == 0:
                    ok = torch.equal(himgs[(nf - 1) & 1], want_all.cpu())
            flag = torch.tensor([1 if ok else 0], device=dev)
            if world > 1:
                torch.distributed.broadcast(flag, 0)
            if int(flag.item()):
                e2e_fps = nf * 1e3 / pipe_ms
                e2e_how = ("pose H2D, frame, 7.68 MB D2H on a copy stream overlapping the next frame's render, every frame; the host "
                           "waits for frame i - 1 while frame i renders")
            else:
                log("[bench] pipelined render e2e: last host frame differs from a synchronous render; reporting the synchronous loop")
        except Exception as e:      # noqa: BLE001 — the synchronous number above stands
            log("[bench] pipelined render e2e failed (%s: %s); reporting the synchronous loop" % (type(e).__name__, e))
    c = r.counters()
    out = None
    if rank == 0:
        # roofline of the frame: B_frame of SURVEY.md 8(d) with the whole merged data set counted once (an upper bound of the
        # distinct rows a frame visits), and — since the frame is instruction bound — march steps per second
        hbm, _, which = measured_peaks()
        ro, rd, _ = get_rays_of_a_view(H, W, Kmat, poses_np[0], device=dev)
        mn, mx = torch.from_numpy(np.asarray(scene["xyz_min"], np.float32)).to(dev), torch.from_numpy(np.asarray(scene["xyz_max"], np.float32)).to(dev)
        tmin, tmax = ru.infer_t_minmax(ro.reshape(-1, 3), rd.reshape(-1, 3), mn, mx, scene["near"], 1e9)
        steps = int(ru.infer_n_samples(rd.reshape(-1, 3).contiguous(), tmin, tmax, scene["stepdist"]).sum().item())
        b_frame = 12 * H * W + (4 + 4 + 48) * n + 88 * 1024
        tj, tj_name = latest_traffic()
        kr = {k: v for k, v in tj.items() if k.startswith("k_render") and isinstance(v, dict)} if tj else {}
        out = {"metric": "merged-VDB render FPS 800x800", "workload": workload, "sharding": sharding, "value": 1e3 / ms, "unit": "frames/s",
               "ms_per_frame": ms, "frames": nf, "poses": "200-pose orbit (pose_spherical(angle, -30, 4)), frame i = pose i mod 200",
               "e2e_fps": e2e_fps, "e2e_how": e2e_how, "e2e_fps_synchronous": e2e_sync_fps, "d2h_bytes_per_frame": H * W * 3 * 4,
               "samples_last_band": c["total"], "inconsistent_rays_last_band": c["inconsistent"], "pixels_marched_twice_last_band": c["remarched"],
               "merged_voxels": n, "row_bands": world, "gpu_launches_per_frame": r.launches_last_call(),
               "roofline": {"bound": "hbm", "achieved": b_frame / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": b_frame / (ms * 1e-3) / 1e9 / hbm, "peak_source": which, "algorithmic_bytes": b_frame,
                            "traffic": (sum(v["dram_bytes"] for v in kr.values()) if kr else None), "traffic_source": tj_name if kr else None,
                            "march_steps_per_frame": steps, "march_steps_per_s": steps / (ms * 1e-3) * (1.0 if world == 1 else 1.0),
                            "note": "B_frame = 12 HW + (4 + 4 + 48) x merged voxels + 88 KB weights; the frame is instruction/latency bound "
                                    "(SURVEY.md 8d), so the march rate (steps of all rays of pose 0 / frame time) is reported beside it"},
               "_renderer": r}
    if peer is not None:
        perr = peer.error()
        peer.close()
        if perr:
            raise RuntimeError("peer frame assembly reported error %d (a rank did not arrive in time)" % perr)
    return out


def bench_s512(args, dev, rank, world, flush, headline):
    """BASELINE configs[4] (SURVEY.md 8d cfg 5): 512^3 stress scene, 65 536-ray batch, train + render at N GPUs."""
    import torch
    from plenvdb_b200 import dist as pdist
    from plenvdb_b200.renderer import merge_grids
    K = args.steps if headline else min(args.steps, args.s512_steps)
    Wm = max(3, min(args.warmup, 3))
    t0 = time.time()
    P, net, den, k0, tr, batches, mask = build_workload_s512(K + Wm, dev, use_tc=not args.fp32_rgbnet)
    log("[bench] rank %d S512 setup %.1fs: %d leaves, occupied %.3f" % (rank, time.time() - t0, den.topo.n_leaf, P["occupied_fraction"]))
    dp = None
    if world > 1 and args.exchange != "none":
        dp = pdist.DataParallelTrainer.wrap(tr, world, exchange=args.exchange)
        stepper = dp.step
    else:
        stepper = tr.step
    T = time_training(tr, stepper, batches, K, Wm, world, dev, S512_RAYS, flush, with_e2e=headline)
    rep = replica_check(tr, world, dev) if (world > 1 and dp is not None) else None
    cnt = T["counters"]
    out = None
    if rank == 0:
        hbm, tf, which = measured_peaks()
        kern = kernel_times(tr, batches, K, Wm, flush, full_step=(world == 1))
        top = max(kern, key=kern.get)
        # Compulsory bytes of the sparse Adam (the HBM-bound kernel at this size) from device-side counts: it reads the gradient of
        # every voxel-channel of every touched leaf (4 B) and, where the gradient is non-zero (stepmode 1 skips the rest), reads
        # p, m, v and writes p, m, v, g (28 B more).  The non-zero count is taken from one extra forward+backward at N = 1.
        roof = None
        if world == 1:
            ro, rd, vd, tg = batches
            tr.forward_backward(ro[Wm], rd[Wm], vd[Wm], tg[Wm])
            nz = int(torch.count_nonzero(tr.density.grad).item()) + int(torch.count_nonzero(tr.k0.grad).item())
            c2 = tr.counters()
            tr.update()
            upd_bytes = 4.0 * 512 * (c2["n_touched_den"] + 12 * c2["n_touched_k0"]) + 28.0 * nz
            t_upd = kern.get("update_fused", float("nan")) * 1e-3
            roof = {"bound": "hbm", "kernel": "update_fused", "achieved": upd_bytes / t_upd / 1e9, "peak": hbm, "unit": "GB/s",
                    "frac": upd_bytes / t_upd / 1e9 / hbm, "algorithmic_bytes": upd_bytes, "nonzero_gradient_channels": nz,
                    "top_kernel": top, "peak_source": which,
                    "note": "sparse Adam over the touched leaves: 4 B per voxel-channel of every touched leaf + 28 B per channel with a non-zero gradient"}
        out = {"workload": "S512: 512^3 noisy shell, %.1f %% of the voxels occupied, %d leaves (pruned topology), %d in_maskcache rays/GPU/iteration, "
                           "768x576 inverse_y cameras" % (100 * P["occupied_fraction"], den.topo.n_leaf, S512_RAYS),
               "train": {"value": T["value"], "unit": "rays/s", "ms_per_step": T["ms_total"] / K, "steps": K, "warmup": Wm, "n_gpus": world,
                         "warm_l2_ms_per_step": T["warm_ms"], "gpu_launches": T["launches"], "kernel_ms": kern,
                         "samples": {"M_alpha": cnt["M_alpha"], "M_keep": cnt["M_keep"], "touched_leaves_density": cnt["n_touched_den"],
                                     "touched_leaves_k0": cnt["n_touched_k0"]},
                         "roofline": roof}}
        if rep is not None:
            out["train"]["parity"] = rep
        if world > 1 and dp is not None:
            out["train"]["exchange_bytes_per_step"] = dp.exchange_bytes()
    if dp is not None:
        dp.close()
    if not args.no_render:
        dend, cold, idx, n = merge_grids(den, k0, mask)
        del tr, den, k0, batches
        torch.cuda.empty_cache()
        a2 = argparse.Namespace(**vars(args))
        a2.frames = args.frames if headline else min(args.frames, 40)
        r = time_render(a2, P, net, dend, cold, idx, n, dev, rank, world,
                        workload="merged S512 scene, 800x800, 200-pose orbit", cap_per_pixel=24)
        if rank == 0:
            r.pop("_renderer", None)
            out["render"] = r
    if headline and rank == 0:
        tr_ = out["train"]
        out = {"metric": "train rays/s (fwd+bwd+update), fine stage", "value": tr_["value"], "unit": "rays/s", "n_gpus": world, "steps": K,
               "warmup": Wm, "ms_per_step": tr_["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": out["workload"], "l2": "flushed between timed steps", "parallelism": "dp%d" % world},
               "e2e": {"value": T.get("e2e_value"), "unit": "rays/s", "h2d_bytes_per_step": 4 * S512_RAYS * 12, "d2h_bytes_per_step": 16,
                       "synchronous_value": T.get("e2e_sync_value")},
               "gpu_launches": tr_["gpu_launches"], "kernel_ms": tr_["kernel_ms"], "roofline": tr_["roofline"],
               "cpu_baseline": None, "render": out.get("render"), "stress_detail": tr_}
    return out


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's host-side sampling path on the box's host cores, all threads: each step = one pass of
    the reference's density / colour trilinear forward + backward over the sample points of n in_maskcache rays."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from plenvdb_b200 import synth            # numpy-only module: no library code is loaded by this arm
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    scene = synth.make_scene(RESO, "sparse")
    poses, K = synth.train_cameras(100), synth.intrinsics(800, 800)
    n_sub = args.cpu_rays
    rng = np.random.default_rng(777)
    sc, sh = synth.mask_scale_shift(scene["mask"].shape, scene["xyz_min"], scene["xyz_max"])
    keep = []
    while sum(len(k[0]) for k in keep) < n_sub:      # in_maskcache rays found with the oracle's own sampler + mask lookup
        m = 8192
        cam, py, px = rng.integers(0, 100, m), rng.integers(0, 800, m), rng.integers(0, 800, m)
        ro, rd, vd = synth.rays_of_pixels(K, poses[cam], px, py)
        pts, mob, rid, _, _, _, _ = orc.sample_pts_on_rays(ro, rd, scene["xyz_min"], scene["xyz_max"], scene["near"], scene["far"],
                                                            scene["stepdist"])
        inb = ~mob
        hitpts = orc.maskcache_lookup(scene["mask"], pts[inb], sc, sh)
        hit = np.zeros(m, bool)
        hit[rid[inb][hitpts]] = True
        keep.append((ro[hit], rd[hit]))
    ro, rd = (np.concatenate([k[i] for k in keep])[:n_sub] for i in range(2))
    lists = cpu_sample_lists(scene, ro, rd)
    probe, _, kind = ref_sampling_pass(scene, lists, threads, 1)
    inner = int(max(1, min(100, 4.0 / max(probe, 1e-4))))       # passes per step: ~4 s of CPU work per step
    for _ in range(max(1, min(args.warmup, 2))):
        ref_sampling_pass(scene, lists, threads, 1)
    ts = []
    for _ in range(args.steps):
        _, t, _ = ref_sampling_pass(scene, lists, threads, inner)
        ts.append(float(np.mean(t)))
    sec = float(np.mean(ts))
    value = n_sub / sec
    counts = {k: int(v[0].size) for k, v in lists.items()}
    sample = ("each step = %d passes of the reference's trilinear density forward (%d points) / backward (%d) and 12-ch colour forward + backward "
              "(%d) over the sample points of %d in_maskcache rays of the F160-sparse workload" % (inner, counts["M1"], counts["M2t"], counts["M3"], n_sub))
    print(json.dumps({
        "impl": "reference", "metric": "train rays/s (fwd+bwd+update), fine stage", "value": value, "unit": "rays/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "F160-sparse fine-stage sample points on the host cores: the reference's host-side sampling path "
                               "(BASELINE.md section 4) — it covers the grid part of the step only (no rgbnet, no compositing, no optimiser), so it "
                               "is an UPPER bound of what the reference's code could do for the whole step on these cores",
                   "n_rays_per_step": n_sub},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="f160", choices=["f160", "s512"],
                    help="headline workload: BASELINE configs[1] (default; S512 rides along as `stress_s512`) or configs[4] alone")
    ap.add_argument("--frames", type=int, default=200, help="frames timed for the render FPS sub-result (the 200-pose orbit)")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-cpu-render", action="store_true", help="skip the CPU frame (render.cpu_baseline / render.parity)")
    ap.add_argument("--no-s512", action="store_true", help="skip the S512 stress sub-benchmark")
    ap.add_argument("--s512-steps", type=int, default=10)
    ap.add_argument("--render-gather", choices=["peer", "nccl"], default="peer",
                    help="N > 1: how rank 0 gets the frame (peer stores over NVLink, or contiguous bands + NCCL gather)")
    ap.add_argument("--exchange", default="nvlink", choices=["nvlink", "nccl", "none"],
                    help="N>1 gradient exchange: own kernels over NVLink peer memory, or NCCL all-reduce of the packed tiles")
    ap.add_argument("--fp32-rgbnet", action="store_true", help="use the fp32 CUDA-core rgbnet instead of the tcgen05 one")
    ap.add_argument("--cpu-rays", type=int, default=8192, help="rays of the CPU legs (oracle parity step, sampling-path baseline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

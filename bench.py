#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native PlenVDB hot path.

Metric (BASELINE.json): train rays/s for one full fine-stage iteration (forward + backward + sparse Adam on the
grids + Adam on rgbnet) on BASELINE config[1]: synthetic NeRF-Synthetic-shaped scene, 100 views 800x800, 8192-ray
batch from the fine stage's `in_maskcache` sampler, random-sparse 160^3 grid, 12-ch k0 + rgbnet.  The merged-VDB
800x800 render FPS (config[2]) is reported in the `render` sub-object of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
`--impl reference` times the CPU restatement of the reference's path (oracle/, multi-threaded over the host cores)
on a bounded sample of the same workload — a reported baseline, not the target.
"""
import argparse
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ workload
def build_workload(n_batches, device, seed=777, pool_candidates=1 << 21, n_rays=N_RAYS, use_tc=True, **trainer_kw):
    """Scene + grids + trainer + `n_batches` device-resident ray batches drawn like the fine stage does
    (ray_sampler='in_maskcache', configs/default.py:73; dvgo.py:583-625 keeps the rays that hit the mask)."""
    import torch
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    scene = synth.make_scene(RESO, "sparse")
    net = synth.rgbnet_init()
    den, k0 = build_scene_grids(scene, device=device)
    tr = FusedTrainer(scene, den, k0, scene["mask"], net, n_rays, device=device, use_tensor_cores=use_tc, **trainer_kw)
    poses, K = synth.train_cameras(100), synth.intrinsics(800, 800)
    rng = np.random.default_rng(seed)
    need = n_batches * n_rays
    got, chunks = 0, []
    while got < need:
        cam = rng.integers(0, 100, pool_candidates)
        py, px = rng.integers(0, 800, pool_candidates), rng.integers(0, 800, pool_candidates)
        ro, rd, vd = synth.rays_of_pixels(K, poses[cam], px, py)
        ro_d, rd_d = torch.from_numpy(ro).to(device), torch.from_numpy(rd).to(device)
        hit = tr.hit_mask(ro_d, rd_d)
        idx = torch.nonzero(hit).reshape(-1)
        chunks.append((ro_d[idx], rd_d[idx], torch.from_numpy(vd).to(device)[idx]))
        got += idx.numel()
    ro = torch.cat([c[0] for c in chunks])[:need].reshape(n_batches, n_rays, 3).contiguous()
    rd = torch.cat([c[1] for c in chunks])[:need].reshape(n_batches, n_rays, 3).contiguous()
    vd = torch.cat([c[2] for c in chunks])[:need].reshape(n_batches, n_rays, 3).contiguous()
    tg = torch.from_numpy(np.random.default_rng(5).uniform(0, 1, (n_batches, n_rays, 3)).astype(np.float32)).to(device)
    return scene, net, den, k0, tr, (ro, rd, vd, tg)


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


# ------------------------------------------------------------------------------------------------ our arm
def algorithmic_bytes(o, n_rays):
    """B_train of SURVEY.md §8(d) from the oracle's distinct-voxel counts for one batch."""
    return (60 * n_rays + 1 * o["V_mask"] + 4 * o["V_den"] + 48 * o["V_k0"] + 2 * (4 * o["V_den_grad"] + 48 * o["V_k0"])
            + 28 * (o["V_den_grad"] + 12 * o["V_k0"]) + 28 * 22019)


def oracle_step(scene, net, rays, threads, n_sub=None):
    """One CPU-oracle training step on the first n_sub rays; returns (seconds, outputs)."""
    from oracle import oracle as orc
    R, act = scene["reso"], scene["active"]
    den, k0 = orc.Grid(R, 1, act), orc.Grid(R, 12, act)
    den.copy_from_dense(scene["density"])
    k0.copy_from_dense(scene["k0"])
    aux = [orc.Grid(R, c, act) for c in (1, 1, 1, 12, 12, 12)]
    keys = ["xyz_min", "xyz_max", "reso", "near", "far", "stepdist", "act_shift", "interval", "fast_color_thres", "bg",
            "weight_main", "weight_entropy_last", "weight_rgbper", "lr_density", "lr_k0", "lr_net", "eps", "beta0", "beta1",
            "den_mode", "k0_mode"]
    cfg = {k: scene[k] for k in keys}
    cfg.update(step=1, do_update=1, n_rays_global=0, threads=threads)
    sub = [a[:n_sub] for a in rays] if n_sub else rays
    nm, nv = np.zeros_like(net), np.zeros_like(net)
    t0 = time.perf_counter()
    out = orc.train_step(cfg, den, aux[0], aux[1], aux[2], k0, aux[3], aux[4], aux[5], scene["mask"], net.copy(), nm, nv, *sub)
    return time.perf_counter() - t0, out


def run_ours(args):
    import torch
    from plenvdb_b200 import _lib, synth
    from plenvdb_b200 import dist as pdist
    rank, local, world = pdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local)
    K, Wm = args.steps, args.warmup
    nb = K + Wm
    t_setup = time.time()
    scene, net, den, k0, tr, batches = build_workload(nb, dev, seed=777 + rank, use_tc=not args.fp32_rgbnet)
    if world > 1:
        dp = pdist.DataParallelTrainer.wrap(tr, world, exchange=args.exchange)
        stepper = dp.step
    else:
        stepper = tr.step
    ro, rd, vd, tg = batches
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    log("[bench] rank %d setup %.1fs, pool ready" % (rank, time.time() - t_setup))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(Wm):
        stepper(ro[i], rd[i], vd[i], tg[i])
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    clk = ClockSampler(local).__enter__()     # sampled through every measurement loop below
    if True:
        barrier()
        l0 = tr.launches_total
        for i in range(K):
            flush.fill_(i & 0xFF)                      # flush L2 between timed iterations (outside the timed events)
            ev[i][0].record()
            stepper(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i])
            ev[i][1].record()
        launches = tr.launches_total - l0
        barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    cnt = tr.counters()
    assert cnt["overflow"] == 0, "sample list overflow: raise cap_*_per_ray"
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * N_RAYS * K / (ms_total * 1e-3)

    # ---- warm L2 (back-to-back) number, informational
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        stepper(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i])
    e1.record()
    barrier()
    warm_ms = e0.elapsed_time(e1) / K

    # ---- end to end through the public API with HOST buffers (pinned): every iteration's rays and targets are copied H2D and its
    # loss words D2H inside the timed region.  Two ways to drive it.  Headline: FusedTrainer.step_from_host_async, the loop a
    # trainer that logs the loss one iteration late uses — copy stream, two staging buffers, iteration i's loss words read
    # while i + 1 runs, one wait per iteration (on iteration i - 1).  Beside it: the strictly synchronous step_from_host (copy,
    # step, read the loss, synchronise, every iteration — what `psnr.item()` costs run.py:590).  The host needs ~95 us to
    # issue an iteration (63 us of it inside pvdb_train_step's ~25 CUDA calls), the 393 KB copy ~10 us: synchronously they
    # add to the step, pipelined they hide under the previous iteration.
    hbatch = torch.stack([x[Wm:].cpu() for x in (ro, rd, vd, tg)], 1).contiguous().pin_memory()   # [K, 4, n, 3]
    dstage = torch.empty_like(hbatch[0], device=dev)
    hloss = torch.empty(4, dtype=torch.float32).pin_memory()
    e2e_ms = {}
    for i in range(min(3, K)):        # untimed: the pipeline's streams, staging and pinned buffers are created on first use
        tr.step_from_host_async(hbatch[i], stepper=None if world == 1 else stepper)
    tr.host_pipeline_flush()
    for mode in ("sync", "pipelined"):
        barrier()
        e0.record()
        for i in range(K):
            if mode == "pipelined":
                tr.step_from_host_async(hbatch[i], stepper=None if world == 1 else stepper)
            elif world == 1:
                tr.step_from_host(hbatch[i])              # one H2D, the step, D2H of the loss, stream sync (the caller reads it)
            else:
                dstage.copy_(hbatch[i], non_blocking=True)
                stepper(dstage[0], dstage[1], dstage[2], dstage[3])
                hloss.copy_(tr.t["loss"], non_blocking=True)
                torch.cuda.current_stream().synchronize()
        last = tr.host_pipeline_flush() if mode == "pipelined" else None
        e1.record()
        barrier()
        assert mode != "pipelined" or (last is not None and bool(torch.isfinite(last).all()))
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms[mode] = float(t.item())
    e2e_sync_value = world * N_RAYS * K / (e2e_ms["sync"] * 1e-3)
    e2e_value = world * N_RAYS * K / (e2e_ms["pipelined"] * 1e-3)
    h2d = 4 * N_RAYS * 3 * 4
    clk.__exit__()

    # ---- exchange kernels (all ranks step together; rank 0 reports)
    xchg = {}
    if world > 1 and dp.peer is not None:
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
        # ---- per-kernel times (CUDA events on the launching stream) for the roofline of the dominant kernel
        _lib.profile_enable(True)
        acc = {}
        for i in range(K):
            flush.fill_(i & 0xFF)
            tr.step(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i]) if world == 1 else tr.run(ro[Wm + i], rd[Wm + i], vd[Wm + i], tg[Wm + i], 3)
            for name, t_ms in _lib.profile_fetch():
                acc.setdefault(name, []).append(t_ms)
        _lib.profile_enable(False)
        kern = {k: float(np.mean(v)) for k, v in acc.items()}
        top = max(kern, key=kern.get)
        # algorithmic bytes / flops of this batch from the oracle (bit-exact parity quantities), bounded CPU sample
        cpu_threads = os.cpu_count() or 1
        rays0 = [x[Wm].cpu().numpy() for x in (ro, rd, vd, tg)]
        n_sub = min(N_RAYS, args.cpu_rays)
        cpu_s, o = oracle_step(scene, net, rays0, cpu_threads, n_sub)
        scale = N_RAYS / n_sub
        M3 = cnt["M_keep"]
        b_train = algorithmic_bytes({k: o[k] * scale for k in ("V_mask", "V_den", "V_k0", "V_den_grad")}, N_RAYS)
        mlp_flops_fwd = 2.0 * M3 * (39 * 128 + 128 * 128 + 128 * 3)
        mlp_pass = 3.0 if tr.use_tc else 1.0      # 3xTF32 issues three tensor-core products per algorithmic product
        kern_flops = {"rgbnet_fwd": mlp_flops_fwd, "rgbnet_bwd": 2.0 * mlp_flops_fwd,
                      "rgbnet_bwd_act": 2.0 * M3 * (128 * 128 + 128 * 12), "rgbnet_bwd_wgrad": 2.0 * M3 * (128 * 128 + 128 * 40 + 3 * 128)}
        # dram bytes per launch of the same kernel from the committed ncu --set full capture (profiles/traffic_r01.json)
        traffic = None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r01.json")))
            kname = {"rgbnet_fwd": "k_rgbnet_fwd_tc", "rgbnet_bwd_act": "k_rgbnet_bwd_act_tc", "rgbnet_bwd_wgrad": "k_rgbnet_bwd_wgrad_tc"}.get(top)
            if tr.use_tc and kname in tj:
                traffic = tj[kname]["dram_bytes"]
        except Exception:
            traffic = None
        # bytes the three MLP kernels move by design (activations handed over through HBM), per kept sample
        design_bytes = {"rgbnet_fwd": M3 * (512 + 512 + 160 + 32 + 48 + 12 + 44.0), "rgbnet_bwd_act": M3 * (512 + 88.0),
                        "rgbnet_bwd_wgrad": M3 * (3 * 512 + 160 + 12 + 16.0) + 148 * 22048 * 4.0}
        if top in kern_flops:
            ach = kern_flops[top] / (kern[top] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf,
                    "traffic": traffic, "peak_source": which,
                    "tensor_core_flops_issued": kern_flops[top] * mlp_pass,
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
            roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                    "peak_source": which}
        step_roof = b_train / (ms_total / K * 1e-3) / 1e9
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
            "warm_l2_ms_per_step": warm_ms,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                    "how": "FusedTrainer.step_from_host_async, every iteration: pinned host batch -> H2D on a copy stream -> step -> loss "
                           "words D2H to pinned host; the host waits for iteration i - 1 and reads its loss while iteration i runs",
                    "synchronous_value": e2e_sync_value,
                    "synchronous_how": "FusedTrainer.step_from_host: H2D, step, D2H of the loss, stream synchronise, every iteration"},
            "gpu_launches": launches,
            "kernel_ms": kern,
            "roofline": roof,
            "step_roofline": {"algorithmic_bytes": b_train, "achieved_GBs": step_roof, "frac_of_hbm": step_roof / hbm,
                              "note": "B_train of SURVEY.md 8(d); the step is latency/compute bound, not HBM bound"},
            "cpu_baseline": {"value": n_sub / cpu_s, "unit": "rays/s", "cores": cpu_threads, "kind": "port",
                             "sample": "1 full oracle step (fwd+bwd+update) on the first %d rays of timed batch 0" % n_sub},
            "clocks": clk.summary(),
        }
        if world > 1:
            result["config"]["exchange_bytes_per_step"] = dp.exchange_bytes()
            if xchg:
                result["exchange_kernel_ms"] = xchg
            result["config"]["exchange"] = ("own kernels over NVLink peer memory (CUDA IPC), no NCCL call in the step"
                                            if dp.peer is not None else "NCCL all-reduce of packed touched-leaf tiles")
            if dp.peer is not None and dp.peer.error():
                result["config"]["exchange_error"] = dp.peer.error()
    # ---- merged-VDB render FPS (config[2]); tile-sharded across ranks
    if not args.no_render:
        r = run_render(args, scene, net, den, k0, dev, rank, world)
        if rank == 0:
            result["render"] = r
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_render(args, scene, net, den, k0, dev, rank, world):
    import torch
    from plenvdb_b200 import dist as pdist
    from plenvdb_b200 import synth
    from plenvdb_b200.plenvdb import MGRenderer
    from plenvdb_b200.renderer import merge_grids
    H = W = 800
    dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(net)
    r = MGRenderer(12, 27, 128, 3, device=dev, use_tensor_cores=not args.fp32_rgbnet)
    r.load_data_dense(dend, cold, idx)
    r.load_params(np.ascontiguousarray(w0.T).reshape(-1), b0, np.ascontiguousarray(w1.T).reshape(-1), b1,
                  np.ascontiguousarray(w2.T).reshape(-1), b2)
    r.setScene(list(scene["reso"]), synth.intrinsics(H, W).reshape(-1), scene["xyz_min"], scene["xyz_max"])
    r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"],
                False, H, W)
    poses = torch.from_numpy(synth.render_cameras(200).reshape(200, 16)).to(dev)
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
    for i in range(3):
        pdist.render_sharded(r, poses[i], rank, world, peer=peer)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nf):
        pdist.render_sharded(r, poses[(3 + i) % 200], rank, world, peer=peer)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item()) / nf
    # e2e: pose from host, frame back to host (7.68 MB) every frame, like run.py:157-167
    hposes = poses.cpu().pin_memory()
    himg = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    barrier()
    e0.record()
    for i in range(nf):
        r.c2w.copy_(hposes[(3 + i) % 200], non_blocking=True)
        img = pdist.render_sharded(r, r.c2w, rank, world, peer=peer)
        if rank == 0:
            himg.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_sync_fps = nf * 1e3 / float(t.item())
    e2e_fps, e2e_how = e2e_sync_fps, "pose H2D, frame, 7.68 MB D2H, stream synchronise, every frame"
    if world == 1:
        # the same with the D2H of frame i overlapping the render of frame i + 1 (two device frames, two pinned host frames, a
        # copy stream); the caller has frame i - 1 on the host while frame i renders.  Checked against a synchronous render.
        try:
            copy = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream()
            outs = [torch.empty((H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)]
            himgs = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
            rendered = [torch.cuda.Event() for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]

            def pipelined(n_frames):
                used = [False, False]
                for i in range(n_frames):
                    k = i & 1
                    if used[k]:
                        cur.wait_event(copied[k])            # frame i - 2 has left outs[k]
                    r.c2w.copy_(hposes[(3 + i) % 200], non_blocking=True)
                    r.render_rows_torch(r.c2w, 0, H, out=outs[k])
                    rendered[k].record(cur)
                    copy.wait_event(rendered[k])
                    with torch.cuda.stream(copy):
                        himgs[k].copy_(outs[k], non_blocking=True)
                        copied[k].record(copy)
                    used[k] = True
                    if i >= 1:
                        copied[k ^ 1].synchronize()          # frame i - 1 is on the host
                copied[(n_frames - 1) & 1].synchronize()

            pipelined(3)
            barrier()
            e0.record()
            pipelined(nf)
            e1.record()
            barrier()
            pipe_ms = e0.elapsed_time(e1)
            want = r.render_rows_torch(poses[(3 + nf - 1) % 200], 0, H).cpu()
            if torch.equal(himgs[(nf - 1) & 1], want):
                e2e_fps = nf * 1e3 / pipe_ms
                e2e_how = ("pose H2D, frame, 7.68 MB D2H on a copy stream overlapping the next frame's render, every frame; the host "
                           "waits for frame i - 1 while frame i renders")
            else:
                log("[bench] pipelined render e2e: last host frame differs from a synchronous render; reporting the synchronous loop")
        except Exception as e:      # noqa: BLE001 — the synchronous number above stands
            log("[bench] pipelined render e2e failed (%s); reporting the synchronous loop" % e)
    c = r.counters()
    if peer is not None:
        perr = peer.error()
        peer.close()
        if perr:
            raise RuntimeError("peer frame assembly reported error %d (a rank did not arrive in time)" % perr)
    return {"metric": "merged-VDB render FPS 800x800", "sharding": sharding, "value": 1e3 / ms, "unit": "frames/s", "ms_per_frame": ms, "frames": nf,
            "e2e_fps": e2e_fps, "e2e_how": e2e_how, "e2e_fps_synchronous": e2e_sync_fps, "d2h_bytes_per_frame": H * W * 3 * 4, "samples_last_band": c["total"],
            "inconsistent_rays_last_band": c["inconsistent"], "pixels_marched_twice_last_band": c["remarched"], "merged_voxels": n, "row_bands": world,
            "gpu_launches_per_frame": r.launches_last_call()}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's path restated on the CPU (oracle/), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from plenvdb_b200 import synth
    threads = os.cpu_count() or 1
    scene = synth.make_scene(RESO, "sparse")
    net = synth.rgbnet_init()
    # in_maskcache rays found with the oracle's own sampler + mask lookup on a candidate pool
    from oracle import oracle as orc
    from plenvdb_b200.synth import mask_scale_shift
    poses, K = synth.train_cameras(100), synth.intrinsics(800, 800)
    n_sub = args.cpu_rays
    rng = np.random.default_rng(777)
    sc, sh = mask_scale_shift(scene["mask"].shape, scene["xyz_min"], scene["xyz_max"])
    keep = []
    while sum(len(k[0]) for k in keep) < n_sub:
        m = 8192
        cam, py, px = rng.integers(0, 100, m), rng.integers(0, 800, m), rng.integers(0, 800, m)
        ro, rd, vd = synth.rays_of_pixels(K, poses[cam], px, py)
        pts, mob, rid, _, _, _, _ = orc.sample_pts_on_rays(ro, rd, scene["xyz_min"], scene["xyz_max"], scene["near"], scene["far"],
                                                            scene["stepdist"])
        inb = ~mob
        hitpts = orc.maskcache_lookup(scene["mask"], pts[inb], sc, sh)
        hit = np.zeros(m, bool)
        hit[rid[inb][hitpts]] = True
        keep.append((ro[hit], rd[hit], vd[hit]))
    ro, rd, vd = (np.concatenate([k[i] for k in keep])[:n_sub] for i in range(3))
    tg = np.random.default_rng(5).uniform(0, 1, (n_sub, 3)).astype(np.float32)
    rays = (ro, rd, vd, tg)
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_step(scene, net, rays, threads)
    ts = [oracle_step(scene, net, rays, threads)[0] for _ in range(args.steps)]
    sec = float(np.sum(ts))
    value = n_sub * args.steps / sec
    sample = "each step = one full oracle iteration (fwd+bwd+update) on %d in_maskcache rays of the F160-sparse workload" % n_sub
    print(json.dumps({
        "impl": "reference", "metric": "train rays/s (fwd+bwd+update), fine stage", "value": value, "unit": "rays/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "F160-sparse fine-stage step on the host cores (CPU restatement of the reference path; the reference "
                               "itself has no CPU sampling path and OpenVDB cannot be built here)", "n_rays_per_step": n_sub},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=20, help="frames timed for the render FPS sub-result")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--render-gather", choices=["peer", "nccl"], default="peer",
                    help="N > 1: how rank 0 gets the frame (peer stores over NVLink, or contiguous bands + NCCL gather)")
    ap.add_argument("--exchange", default="nvlink", choices=["nvlink", "nccl"],
                    help="N>1 gradient exchange: own kernels over NVLink peer memory, or NCCL all-reduce of the packed tiles")
    ap.add_argument("--fp32-rgbnet", action="store_true", help="use the fp32 CUDA-core rgbnet instead of the tcgen05 one")
    ap.add_argument("--cpu-rays", type=int, default=2048, help="rays per CPU-baseline step (bounded sample)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

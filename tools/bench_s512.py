"""SURVEY.md §8(d) cfg 5 — S512 stress configuration: 512^3 shell scene on a pruned topology, 65 536 in_maskcache rays per
iteration from 768x576 inverse_y cameras.  Not a BASELINE bench line (bench.py keeps that contract); prints one JSON line
with the step time, the per-kernel split and the sample counts.   python tools/bench_s512.py [--reso 512] [--rays 65536]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenvdb_b200 import _lib, synth                      # noqa: E402
from plenvdb_b200.fused import FusedTrainer, build_stress_scene   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reso", type=int, default=512)
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda")
    t0 = time.time()
    P, den, k0, mask = build_stress_scene(a.reso)
    torch.cuda.synchronize()
    sys.stderr.write("[s512] scene built in %.1fs: %d leaves, occupied %.3f\n" % (time.time() - t0, den.topo.n_leaf, P["occupied_fraction"]))
    net = synth.rgbnet_init()
    n = a.rays
    tr = FusedTrainer(P, den, k0, mask, net, n)
    H, W = 576, 768
    K = np.array([[800.0, 0, W / 2], [0, 800.0, H / 2], [0, 0, 1]], np.float32)
    rng = np.random.default_rng(7)
    poses = np.stack([synth.pose_spherical(rng.uniform(-180, 180), rng.uniform(-90, 0), rng.uniform(2.5, 3.5)) for _ in range(100)])
    poses[:, :3, 1:3] *= -1      # inverse_y (OpenCV-style) cameras look along +z: flip the y / z axes of the Blender-style poses
    nb = a.steps + a.warmup
    need, got, chunks = nb * n, 0, []
    while got < need:
        m = 1 << 21
        cam = rng.integers(0, 100, m)
        px, py = rng.integers(0, W, m), rng.integers(0, H, m)
        ro, rd, vd = synth.rays_of_pixels(K, poses[cam], px, py, inverse_y=True)
        ro_d, rd_d = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
        idx = torch.nonzero(tr.hit_mask(ro_d, rd_d)).reshape(-1)
        chunks.append((ro_d[idx], rd_d[idx], torch.from_numpy(vd).to(dev)[idx]))
        got += idx.numel()
        assert got > 0, "no ray hits the occupancy mask: camera convention?"
        sys.stderr.write("[s512] rays %d / %d after %.1fs\n" % (got, need, time.time() - t0))
    ro, rd, vd = (torch.cat([c[i] for c in chunks])[:need].reshape(nb, n, 3).contiguous() for i in range(3))
    tg = torch.rand((nb, n, 3), device=dev)
    setup_s = time.time() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(a.warmup):
        tr.step(ro[i], rd[i], vd[i], tg[i])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.fill_(i & 0xFF)
        evs[i][0].record()
        tr.step(ro[a.warmup + i], rd[a.warmup + i], vd[a.warmup + i], tg[a.warmup + i])
        evs[i][1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs]))
    _lib.profile_enable(True)
    acc = {}
    for i in range(a.steps):
        tr.step(ro[a.warmup + i], rd[a.warmup + i], vd[a.warmup + i], tg[a.warmup + i])
        for name, t_ms in _lib.profile_fetch():
            acc.setdefault(name, []).append(t_ms)
    _lib.profile_enable(False)
    c = tr.counters()
    print(json.dumps({
        "workload": "S%d shell scene, pruned topology, %d in_maskcache rays/iteration, 768x576 inverse_y cameras" % (a.reso, n),
        "ms_per_step": ms, "rays_per_s": n / (ms * 1e-3), "steps": a.steps, "warmup": a.warmup,
        "n_leaf": den.topo.n_leaf, "occupied_fraction": P["occupied_fraction"],
        "samples": {"M_alpha": c["M_alpha"], "M_keep": c["M_keep"], "touched_leaves_density": c["n_touched_den"],
                    "touched_leaves_k0": c["n_touched_k0"], "overflow": c["overflow"]},
        "kernel_ms": {k: float(np.mean(v)) for k, v in acc.items()},
        "device_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "setup_s": setup_s,
        "l2": "flushed between timed steps"}))


if __name__ == "__main__":
    main()

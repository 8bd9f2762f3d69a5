"""How the frame's kernels scale when one GPU renders only a 1/N share of the rows (strong-scaling floor of tile sharding),
measured on ONE GPU: per-kernel CUDA-event times of rank 0's interleaved share and of the middle contiguous band."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenvdb_b200 import _lib, synth                       # noqa: E402
from plenvdb_b200 import dist as pdist                     # noqa: E402
from plenvdb_b200.fused import build_scene_grids           # noqa: E402
from plenvdb_b200.plenvdb import MGRenderer                # noqa: E402
from plenvdb_b200.renderer import merge_grids              # noqa: E402

dev = torch.device("cuda", 0)
H = W = 800
scene = synth.make_scene(160, os.environ.get("SCENE", "dense"))      # cfg 3: the dense-fill scene
den, k0 = build_scene_grids(scene, device=dev)
dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
w0, b0, w1, b1, w2, b2 = synth.unpack_net(synth.rgbnet_init())
r = MGRenderer(12, 27, 128, 3, device=dev)
r.load_data_dense(dend, cold, idx)
r.load_params(np.ascontiguousarray(w0.T).reshape(-1), b0, np.ascontiguousarray(w1.T).reshape(-1), b1,
              np.ascontiguousarray(w2.T).reshape(-1), b2)
r.setScene(list(scene["reso"]), synth.intrinsics(H, W).reshape(-1), scene["xyz_min"], scene["xyz_max"])
r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"], False, H, W)
poses = torch.from_numpy(synth.render_cameras(200).reshape(200, 16)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, nf=20):
    for i in range(3):
        fn(poses[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nf):
        fn(poses[(3 + i) % 200])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nf
    _lib.profile_enable(True)
    acc = {}
    for i in range(nf):
        fn(poses[(3 + i) % 200])
        torch.cuda.synchronize()
        for name, t in _lib.profile_fetch():
            acc[name] = acc.get(name, 0.0) + t / nf
    _lib.profile_enable(False)
    return ms, acc


out = []
for world in (1, 2, 4, 8):
    for band_rows in ((16,) if world == 1 else (16,)):
        ms, acc = timed(lambda c: r.render_interleaved_torch(c, band_rows, 0, world))
        out.append(dict(mode="interleaved", world=world, band_rows=band_rows, ms_per_frame=ms, kernels_us={k: round(v * 1e3, 1) for k, v in acc.items()}))
        print(json.dumps(out[-1]), flush=True)
    if False:
        lo, hi = pdist.shard_range(H, world // 2, world)
        ms, acc = timed(lambda c: r.render_rows_torch(c, lo, hi))
        out.append(dict(mode="contiguous middle band", world=world, rows=[lo, hi], ms_per_frame=ms, kernels_us={k: round(v * 1e3, 1) for k, v in acc.items()}))
        print(json.dumps(out[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/render_share_timing_lanes%s.json" % os.environ.get("PVDB_RENDER_LANES", "1"), "w"), indent=1)

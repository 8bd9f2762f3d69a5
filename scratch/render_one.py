"""Render a few frames of the bench scene (optionally only rank 0's interleaved share of `world`) — target for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenvdb_b200 import synth
from plenvdb_b200.fused import build_scene_grids
from plenvdb_b200.plenvdb import MGRenderer
from plenvdb_b200.renderer import merge_grids
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H = W = 800
scene = synth.make_scene(160, "sparse")
den, k0 = build_scene_grids(scene)
dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
w0, b0, w1, b1, w2, b2 = synth.unpack_net(synth.rgbnet_init())
r = MGRenderer(12, 27, 128, 3)
r.load_data_dense(dend, cold, idx)
r.load_params(np.ascontiguousarray(w0.T).reshape(-1), b0, np.ascontiguousarray(w1.T).reshape(-1), b1, np.ascontiguousarray(w2.T).reshape(-1), b2)
r.setScene(list(scene["reso"]), synth.intrinsics(H, W).reshape(-1), scene["xyz_min"], scene["xyz_max"])
r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"], False, H, W)
poses = torch.from_numpy(synth.render_cameras(200).reshape(200, 16)).cuda()
for i in range(frames):
    r.render_interleaved_torch(poses[3 + i], 4, 0, world)
torch.cuda.synchronize()
print(r.counters())

import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenvdb_b200 import synth
from plenvdb_b200.fused import FusedTrainer, build_scene_grids, get_rays_of_a_view
scene = synth.make_scene(64, "dense"); net = synth.rgbnet_init()
den, k0 = build_scene_grids(scene)
tr = FusedTrainer(scene, den, k0, scene["mask"], net, 2048)
H, W = 60, 70
K = synth.intrinsics(H, W); c2w = synth.render_cameras(8)[2]
img = tr.render_view(H, W, K, c2w)
ro, rd, vd = [t.reshape(-1, 3) for t in get_rays_of_a_view(H, W, K, c2w, device="cuda")]
ref = torch.cat([tr.forward(ro[a:a + 2048].contiguous(), rd[a:a + 2048].contiguous(), vd[a:a + 2048].contiguous()).clone() for a in range(0, H * W, 2048)])
print("render_view ok:", bool(torch.equal(img.reshape(-1, 3), ref)), "lit", float((img != scene["bg"]).float().mean()))

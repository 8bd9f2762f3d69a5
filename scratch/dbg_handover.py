import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenvdb_b200 import synth
from plenvdb_b200.fused import build_scene_grids
from plenvdb_b200.plenvdb import MGRenderer
from plenvdb_b200.renderer import merge_grids

def make(scene_reso, variant, H, W, P):
    scene = synth.make_scene(scene_reso, variant)
    den, k0 = build_scene_grids(scene)
    dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(synth.rgbnet_init())
    r = MGRenderer(12, 27, 128, 3, px_entries=P)
    r.load_data_dense(dend, cold, idx)
    r.load_params(np.ascontiguousarray(w0.T).reshape(-1), b0, np.ascontiguousarray(w1.T).reshape(-1), b1, np.ascontiguousarray(w2.T).reshape(-1), b2)
    r.setScene(list(scene["reso"]), synth.intrinsics(H, W).reshape(-1), scene["xyz_min"], scene["xyz_max"])
    r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"], False, H, W)
    return r

# 1. bench frame: distribution of samples per pixel and how many pixels are marched again
r = make(160, "sparse", 800, 800, 32)
poses = torch.from_numpy(synth.render_cameras(200).reshape(200, 16)).cuda()
for i in (3, 50, 120):
    r.render_rows_torch(poses[i], 0, 800)
    ns = r.s["n_samples"].cpu().numpy()
    c = r.counters()
    print("frame", i, c, "active", int((ns > 0).sum()), "max ns", ns.max(), "ns>32:", int((ns > 32).sum()), "ns>64:", int((ns > 64).sum()),
          "pct", np.percentile(ns[ns > 0], [50, 90, 99, 99.9]).tolist())
    fl = r.s["fallback_list"][:c["remarched"]].cpu().numpy()
    if len(fl):
        print("  fallback rows min/max", fl.min() // 800, fl.max() // 800, "of which ns<=32:", int((ns[fl] <= 32).sum()))

# 2. the failing test case
H, W = 150, 170
ref = None
for P in (0, 2, 32):
    r = make(96, "dense", H, W, P)
    c2w = torch.from_numpy(synth.render_cameras(8)[1].reshape(-1).copy()).cuda()
    img = r.render_rows_torch(c2w, 0, H).clone()
    c = r.counters()
    ns = r.s["n_samples"].clone()
    if ref is None:
        ref = (img, ns)
        print("P=0", c)
        continue
    d = (img - ref[0]).abs().amax(-1).reshape(-1)
    bad = torch.nonzero(d > 0).flatten()
    fl = set(r.s["fallback_list"][:c["remarched"]].cpu().numpy().tolist())
    print("P=%d" % P, c, "ns equal", bool(torch.equal(ns, ref[1])), "pixels differing", bad.numel(), "max diff", float(d.max()))
    for p in bad[:8].tolist():
        print("   pixel", p, "ns", int(ns[p]), "remarched" if p in fl else "handed", "img", img.reshape(-1, 3)[p].tolist(), "ref", ref[0].reshape(-1, 3)[p].tolist())

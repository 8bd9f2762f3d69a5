import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda")
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(2, dev)
tr.run(ro[0], rd[0], vd[0], tg[0], 1)
torch.cuda.synchronize()
ca = tr.t["cnt_alpha"].cpu().numpy(); ck = tr.t["cnt_keep"].cpu().numpy(); ns = tr.t["n_steps"].cpu().numpy(); cm = tr.t["cnt_mask"].cpu().numpy()
for name, a in (("cnt_alpha", ca), ("cnt_keep", ck), ("n_steps", ns), ("cnt_mask", cm)):
    print(name, "mean %.1f max %d p99 %d  >128: %d  >64: %d" % (a.mean(), a.max(), np.percentile(a, 99), (a > 128).sum(), (a > 64).sum()))

import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda")
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(1, dev)
tr.run(ro[0], rd[0], vd[0], tg[0], 3)
torch.cuda.synchronize()
g = tr.net_grad.cpu().numpy()
np.save(sys.argv[1], g)
print("saved", sys.argv[1], float(np.abs(g).sum()))
if len(sys.argv) > 2:
    a = np.load(sys.argv[2])
    print("bitwise equal:", np.array_equal(a.view(np.uint32), g.view(np.uint32)), "max abs diff", float(np.abs(a - g).max()), "max rel", float((np.abs(a - g) / (np.abs(a) + 1e-12)).max()))

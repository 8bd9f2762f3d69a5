import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from plenvdb_b200 import _lib
dev = torch.device("cuda")
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(4, dev)
for i in range(4):
    tr.run(ro[i], rd[i], vd[i], tg[i], 3)
torch.cuda.synchronize()
buf = np.zeros((148, 16), np.int64)
fn = _lib.lib.pvdb_debug_wgrad_timing
fn.argtypes = [C.c_void_p]
assert fn(buf.ctypes.data) == 0
t0 = buf[:, 0]
print("n_it", buf[:, 2].mean())
print("init", (buf[:, 1] - t0).mean())
print("issuer: wait %d  end %d" % (buf[:, 3].mean(), (buf[:, 4] - t0).mean()))
print("loader: wait %d  end %d" % (buf[:, 5].mean(), (buf[:, 6] - t0).mean()))
print("conv:   wait %d  end %d" % (buf[:, 7].mean(), (buf[:, 8] - t0).mean()))
print("conv lo-set / A-set wait %d" % buf[:, 12].mean())
print("drain start %d  flush end %d" % ((buf[:, 9] - t0).mean(), (buf[:, 10] - t0).mean()))
print("span", buf[:, 10].max() - t0.min())
print("net_grad abs sum", float(tr.net_grad.abs().sum()))

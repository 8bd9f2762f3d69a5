"""Phase timing of the tcgen05 activation-gradient kernel (needs a build with PVDB_EXTRA_NVCC_FLAGS=-DPVDB_TC_TIMING)."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from plenvdb_b200 import _lib
dev = torch.device("cuda")
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(4, dev)
for i in range(4):
    tr.run(ro[i], rd[i], vd[i], tg[i], 3)
torch.cuda.synchronize()
buf = np.zeros((148, 8, 8), np.int64)
fn = _lib.lib.pvdb_debug_b1_timing
fn.argtypes = [C.c_void_p]
assert fn(buf.ctypes.data) == 0
t0 = buf[:, 0, 0]
for k in range(7):
    r = buf[:, k, :]
    ok = r[:, 3] > 0
    if not ok.any(): continue
    r = r[ok]
    print("iter %d n=%d start=%d  waitD0=%d ph23=%d  sync+waitD1=%d  ph01=%d  scatter=%d" % (
        k, ok.sum(), (r[:, 0] - t0[ok]).mean(), (r[:, 6] - r[:, 5]).mean() if k else 0, (r[:, 1] - r[:, 0]).mean(), (r[:, 2] - r[:, 1]).mean(),
        (r[:, 3] - r[:, 2]).mean(), (r[:, 4] - r[:, 3]).mean() if k else 0))
print("end", (buf[:, 7, 7] - t0).mean(), (buf[:, 7, 7] - t0).max())

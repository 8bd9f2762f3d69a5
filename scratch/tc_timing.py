"""Phase timing of the tcgen05 forward kernel (needs a build with PVDB_EXTRA_NVCC_FLAGS=-DPVDB_TC_TIMING)."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from plenvdb_b200 import _lib
dev = torch.device("cuda")
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(4, dev)
for i in range(4):
    tr.run(ro[i], rd[i], vd[i], tg[i], 1)      # forward only: the last TC kernel run is k_rgbnet_fwd_tc
torch.cuda.synchronize()
buf = np.zeros((148, 8, 16), np.int64)
fn = _lib.lib.pvdb_debug_tc_timing
fn.argtypes = [C.c_void_p]
assert fn(buf.ctypes.data) == 0
names = ["entry", "setup", "top", "waitL0", "ep0", "waitL1", "next", "ep1"]
t0 = buf[:, 0, 0][:, None]
print("setup cycles (mean over CTAs):", (buf[:, 0, 1] - buf[:, 0, 0]).mean())
for tile in range(5):
    row = buf[:, tile, :]
    ok = row[:, 7] > 0
    if tile > 0:
        ok &= buf[:, tile, 2] > buf[:, tile - 1, 2]
    if not ok.any():
        continue
    d = np.diff(row[ok][:, 2:8], axis=1).mean(0)
    print("tile", tile, "n=%d" % ok.sum(), " ".join("%s=%d" % (n, v) for n, v in zip(names[3:], d)), "total=%d" % d.sum(),
          "start=%d" % (row[ok][:, 2] - buf[ok, 0, 0]).mean())
print("kernel span cycles (max end - min entry):", buf[:, :, 7].max() - buf[:, 0, 0].min())

for tile in range(5):
    row = buf[:, tile, :]
    ok = row[:, 7] > 0
    if not ok.any(): continue
    r = row[ok]
    print("tile", tile, "producer: feat start=%d dur=%d" % ((r[:, 12] - buf[ok, 0, 0]).mean(), (r[:, 13] - r[:, 12]).mean()))

for tile in range(5):
    row = buf[:, tile, :]
    ok = row[:, 7] > 0
    if not ok.any(): continue
    r = row[ok]
    print("tile", tile, "ep1 detail: ld+relu=%d act(j=0)=%d fma(j=0)=%d rest=%d" % tuple((r[:, b] - r[:, a]).mean() for a, b in ((6, 8), (8, 9), (9, 10), (10, 7))))

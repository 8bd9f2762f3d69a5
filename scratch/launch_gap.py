"""Where the time between the kernels of the step goes: %globaltimer at the CTA entries / exits of the rgbnet forward against the
stamps of the step (PVDB_STAMPS=1; needs a build with PVDB_EXTRA_NVCC_FLAGS=-DPVDB_TC_TIMING)."""
import os, sys, ctypes as C
os.environ["PVDB_STAMPS"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch, bench
from plenvdb_b200 import _lib
dev = torch.device("cuda")
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(8, dev)
fn = _lib.lib.pvdb_debug_tc_gt
fn.argtypes = [C.c_void_p]
rows = []
for i in range(8):
    tr.step(ro[i], rd[i], vd[i], tg[i])
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 64)()
    assert _lib.lib.pvdb_debug_stamps_fetch(buf) == 0
    t = np.array(list(buf), np.float64)
    gt = np.zeros((148, 2), np.uint64)
    assert fn(gt.ctypes.data) == 0
    gt = gt.astype(np.float64)
    if i >= 3:
        rows.append(((t[1] - t[0]) / 1e3, (gt[:, 0].min() - t[1]) / 1e3, (gt[:, 0].max() - gt[:, 0].min()) / 1e3, (gt[:, 1].max() - gt[:, 0].min()) / 1e3,
                     (gt[:, 1].max() - gt[:, 1].min()) / 1e3, (t[2] - gt[:, 1].max()) / 1e3, (t[2] - t[1]) / 1e3))
print("us: start->emit stamp %.1f | stamp->first CTA entry %.1f | entry spread %.1f | first entry->last exit %.1f | exit spread %.1f | last exit->next stamp %.1f | stamp to stamp %.1f"
      % tuple(np.mean(rows, 0)))

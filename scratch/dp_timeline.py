"""Device-side timeline of the data-parallel step (PVDB_STAMPS=1): torchrun --nproc-per-node N scratch/dp_timeline.py"""
import os, sys, ctypes as C
os.environ["PVDB_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from plenvdb_b200 import _lib, dist as pdist
rank, local, world = pdist.init_from_env()
dev = torch.device("cuda", local)
n = 12
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(n, dev, seed=777 + rank)
mode = sys.argv[1] if len(sys.argv) > 1 else "nvlink"
dp = pdist.DataParallelTrainer.wrap(tr, world, exchange=mode) if (world > 1 and mode != "none") else None
step = dp.step if dp else tr.step
names = {0: "start", 8: "march", 9: "scan", 1: "emit", 2: "fwd", 3: "composite", 4: "bwd_act", 5: "wgrad+red", 6: "net adam", 7: "join", 10: "s:union>", 11: "s:union<", 12: "s:tiles>", 13: "s:tiles<", 14: "s:leafadam", 20: "u:in", 21: "u:pub", 22: "u:sig", 23: "u:waited", 24: "u:out", 25: "rs:in", 26: "rs:Bwaited", 27: "rs:cta0done", 28: "rs:last"}
acc = []
for i in range(n):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    step(ro[i], rd[i], vd[i], tg[i])
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 64)()
    assert _lib.lib.pvdb_debug_stamps_fetch(buf) == 0
    t = np.array(list(buf), np.float64)
    if i >= 4:
        acc.append((t - t[0]) / 1e3)
a = np.mean(acc, 0)
for r in range(world):
    if world > 1:
        torch.distributed.barrier()
    if r == rank:
        print("rank %d  " % rank + "  ".join("%s %.0f" % (names[k], a[k]) for k in sorted(names) if a[k] > -1e5 and abs(a[k]) < 1e6), flush=True)
if dp:
    dp.close()

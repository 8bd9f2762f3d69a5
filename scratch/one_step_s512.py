"""A few fused steps of the S512 stress workload (for ncu captures): python scratch/one_step_s512.py [n_steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
P, net, den, k0, tr, (ro, rd, vd, tg), mask = bench.build_workload_s512(n, torch.device("cuda", 0))
for i in range(n):
    tr.step(ro[i], rd[i], vd[i], tg[i])
torch.cuda.synchronize()
print(tr.counters(), int(tr.t["counters"][8]))

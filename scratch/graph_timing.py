"""Where does the synchronous host-fed iteration spend its time?  direct launches vs graph replay, device-side and host-side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
K = 200
dev = torch.device("cuda", 0)
scene, net, den, k0, tr, (ro, rd, vd, tg) = bench.build_workload(4, dev)
hb = torch.stack([x.cpu() for x in (ro, rd, vd, tg)], 1).contiguous().pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timed(fn, label, sync_each=False):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for i in range(K):
        fn(i)
        if sync_each:
            torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    print("%-46s host %.1f us/it   device span %.1f us/it" % (label, (time.perf_counter() - t0) / K * 1e6, e0.elapsed_time(e1) / K * 1e3), flush=True)

tr.use_graph = False
timed(lambda i: tr.step(ro[i & 3], rd[i & 3], vd[i & 3], tg[i & 3]), "direct, async (device-resident batch)")
timed(lambda i: tr.step(ro[i & 3], rd[i & 3], vd[i & 3], tg[i & 3]), "direct, sync each", True)
timed(lambda i: tr.step_from_host(hb[i & 3]), "step_from_host direct")
tr.use_graph = True
timed(lambda i: tr.step_from_host(hb[i & 3]), "step_from_host graph")
st = tr._stage
timed(lambda i: tr._step_graphed(st), "graph replay only, async")
timed(lambda i: tr._step_graphed(st), "graph replay only, sync each", True)
g = tr._graphs[(st.data_ptr(), st.shape[1])]["graph"]
timed(lambda i: g.replay(), "raw replay (pinned scalars not rewritten), async")
import torch.cuda.nvtx
# per-kernel view of one replay vs one direct step is in the ncu launch list; here: the update kernel alone cannot be isolated

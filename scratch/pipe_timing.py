import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenvdb_b200 import synth
from plenvdb_b200.fused import FusedTrainer, build_scene_grids
scene = synth.make_scene(96, "dense")
net = synth.rgbnet_init()
rays = synth.ray_batch(8192, H=400, W=400, K=synth.intrinsics(400, 400), seed=777)
den, k0 = build_scene_grids(scene)
tr = FusedTrainer(scene, den, k0, scene["mask"], net, 8192)
hb = torch.from_numpy(np.stack(rays, 0).copy()).pin_memory()
N = 60
def wall(fn, flush=None):
    for _ in range(5): fn()
    if flush: flush()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): fn()
    if flush: flush()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / N * 1e3
print("sync      ms/iter", wall(lambda: tr.step_from_host(hb)))
print("pipelined ms/iter", wall(lambda: tr.step_from_host_async(hb), tr.host_pipeline_flush))
# same-stream pipelining: no copy stream, no per-iteration host wait except on the loss of the previous iteration
stage = [torch.empty((4, 8192, 3), device="cuda") for _ in range(2)]
loss = [torch.empty(4).pin_memory() for _ in range(2)]
done = [torch.cuda.Event() for _ in range(2)]
st = dict(i=0)
def same_stream():
    k = st["i"] & 1
    stage[k].copy_(hb, non_blocking=True)
    tr.step(stage[k][0], stage[k][1], stage[k][2], stage[k][3])
    loss[k].copy_(tr.t["loss"], non_blocking=True)
    done[k].record()
    st["i"] += 1
    if st["i"] > 1: done[k ^ 1].synchronize()
print("same-stream, loss one late ms/iter", wall(same_stream, torch.cuda.synchronize))
# segment costs of the pipelined call on the host
p = tr._pipe
seg = np.zeros(6)
for it in range(N):
    k = p["i"] & 1; cur = torch.cuda.current_stream()
    t = [time.perf_counter()]
    p["copy"].wait_event(p["done"][k]); t.append(time.perf_counter())
    with torch.cuda.stream(p["copy"]):
        p["stage"][k].copy_(hb, non_blocking=True); p["copied"][k].record(p["copy"])
    t.append(time.perf_counter())
    cur.wait_event(p["copied"][k]); t.append(time.perf_counter())
    s = p["stage"][k]; tr.step(s[0], s[1], s[2], s[3]); t.append(time.perf_counter())
    p["loss"][k].copy_(tr.t["loss"], non_blocking=True); p["done"][k].record(cur); p["i"] += 1; t.append(time.perf_counter())
    p["done"][k ^ 1].synchronize(); t.append(time.perf_counter())
    seg += np.diff(t)
print("host us per segment [wait_event, h2d+record, wait copied, step call, loss d2h+record, sync prev]:", (seg / N * 1e6).round(1).tolist())

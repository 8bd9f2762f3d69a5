import torch, time, subprocess
print(subprocess.run("nvidia-smi topo -m | head -8", shell=True, capture_output=True, text=True).stdout)
print("can access peer", torch.cuda.can_device_access_peer(0, 1))
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
for n in (4096, 1 << 20, 2 << 20, 256 << 20):
    for _ in range(3):
        b[:n].copy_(a[:n])
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(1):
        e0.record()
        for _ in range(10):
            b[:n].copy_(a[:n])
        e1.record()
        torch.cuda.synchronize(1)
    ms = e0.elapsed_time(e1) / 10
    print(n, "bytes  %.4f ms  %.1f GB/s" % (ms, n / ms / 1e6))

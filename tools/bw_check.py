import torch, time
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("copy GB/s", 2 * a.numel() * 2 / best / 1e6)

# host <-> device over PCIe with pinned buffers (the bound of bench.py's e2e line: 3.6 GB of float32 output per step)
h = torch.empty(1 << 28, dtype=torch.float32).pin_memory()          # 1 GiB
d = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(name, "pinned GB/s", h.numel() * 4 / best / 1e6)

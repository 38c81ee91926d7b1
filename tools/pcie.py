import torch, time
x = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)   # 1 GiB
d = torch.empty_like(x, device="cuda")
for name, f in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [f() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    print(name, "pinned 1 GiB:", 3 * x.numel() * 4 / e0.elapsed_time(e1) / 1e6, "GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True); d2 = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
with torch.cuda.stream(s2): y.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("bidirectional 1 GiB each way:", x.numel() * 4 / dt / 1e9, "GB/s per direction")

import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * n / (time.perf_counter() - t0) / 1e9))
# both directions at once: 16 B out per 8 B in
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("duplex: d2h %.1f GB/s with h2d %.1f GB/s concurrently" % (5 * n / dt / 1e9, 5 * n / 2 / dt / 1e9))
# chunked 32 MB d2h
c = 32 << 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(0, n, c): h[i:i + c].copy_(d[i:i + c], non_blocking=True)
torch.cuda.synchronize()
print("d2h in 32 MB chunks %.1f GB/s" % (n / (time.perf_counter() - t0) / 1e9))

"""GPU probe (not a test): pinned host <-> device copy bandwidth of this box."""
import time, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("D2H", h, d), ("H2D", d, h)):
    for _ in range(2):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("%s pinned 1 GiB x4: %.1f GB/s" % (name, 4 * n / dt / 1e9))
# two concurrent D2H streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("D2H two streams: %.1f GB/s total" % (8 * n / dt / 1e9))

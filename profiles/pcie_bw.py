"""Pinned host <-> device copy bandwidth of the box (context for the e2e number)."""
import time
import torch
n = 1 << 28  # 1 GiB of float32
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * n * 4 / (time.perf_counter() - t) / 1e9))
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("H2D+D2H concurrently: %.1f GB/s each way" % (5 * n * 4 / (time.perf_counter() - t) / 1e9))

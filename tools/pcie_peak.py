"""Measures pinned host -> device copy bandwidth on this box (one 2 GiB cudaMemcpyAsync, best of 5): the ceiling of bench.py's e2e leg."""
import torch
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
best = 0.0
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d.copy_(h, non_blocking=True); b.record(); torch.cuda.synchronize()
    best = max(best, n / (a.elapsed_time(b) * 1e-3) / 1e9)
print("pinned H2D: %.1f GB/s" % best)
best = 0.0
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); h.copy_(d, non_blocking=True); b.record(); torch.cuda.synchronize()
    best = max(best, n / (a.elapsed_time(b) * 1e-3) / 1e9)
print("pinned D2H: %.1f GB/s" % best)

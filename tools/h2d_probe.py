"""Raw pinned host->device copy rate of this box (context for the e2e number): python tools/h2d_probe.py"""
import torch
n = 150 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, n // 2, n // 4, n // 8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        d[:chunk].copy_(h[:chunk], non_blocking=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        d[:chunk].copy_(h[:chunk], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"H2D {chunk >> 20} MiB: {ms:.3f} ms  {chunk / ms / 1e6:.1f} GB/s")

"""Measures write-only / read-only / copy HBM bandwidth with torch (context for the decode roofline: decode writes
8.5x more bytes than it reads)."""
import torch
n = 8 * 1024**3  # 16 GiB of int16
a = torch.empty(n, dtype=torch.int16, device="cuda")
b = torch.empty(n, dtype=torch.int16, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_()); print(f"write-only fill : {2*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: a.fill_(7)); print(f"write-only fill7: {2*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy (r+w)      : {4*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: torch.sum(a[: n // 2].view(torch.int32))); print(f"read-only sum   : {n/ms/1e6:8.1f} GB/s")

# PCIe ceilings for the e2e path (pinned host memory)
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
ms = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H pinned 1 GiB : {(1<<30)/ms/1e6:8.1f} GB/s")
ms = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D pinned 1 GiB : {(1<<30)/ms/1e6:8.1f} GB/s")

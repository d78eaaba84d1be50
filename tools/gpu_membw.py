"""Pure-write / copy bandwidth on this GPU (context for the write-dominated k>=7 kernels)."""
import torch
n = 8 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_()); print(f"memset 8GiB: {ms:.3f} ms -> {n/ms/1e6:.0f} GB/s write")
ms = t(lambda: a.view(torch.int32).fill_(7)); print(f"fill_ 8GiB: {ms:.3f} ms -> {n/ms/1e6:.0f} GB/s write")
ms = t(lambda: b.copy_(a)); print(f"copy 8GiB: {ms:.3f} ms -> {2*n/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: a.view(torch.int32).sum()); print(f"sum 8GiB: {ms:.3f} ms -> {n/ms/1e6:.0f} GB/s read")

#!/usr/bin/env python
"""Write-only / copy HBM bandwidth of the GPU (torch fill_ / copy_ over 2 GiB, best of 10, CUDA events).
MEASURED_PEAKS.json holds the copy figure (half read, half write); kernels that mostly WRITE (the affine map: 3 bytes
written per byte read; the head: 9 to 1) are bounded by the write-only figure, which this probe measures.
r29, B200: write-only 3904 GB/s, copy 6559 GB/s."""
import torch, time
x = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
w = t(lambda: x.fill_(3))
c = t(lambda: y.copy_(x))
r = t(lambda: x.view(torch.int32).sum())
print("write-only %.0f GB/s  copy(r+w) %.0f GB/s  read-only(sum) %.0f GB/s" % (x.numel() / w / 1e6, 2 * x.numel() / c / 1e6, x.numel() / r / 1e6))

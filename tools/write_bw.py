#!/usr/bin/env python
"""Probe: achievable write-only and copy HBM bandwidth with library kernels (context for the roofline)."""
import torch
n = 200_000_000
x = torch.empty(n, dtype=torch.float64, device='cuda')
y = torch.empty(n, dtype=torch.float64, device='cuda')
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.5)); print(f'fill_  1.6 GB: {ms:.3f} ms  {n*8/ms/1e6:.0f} GB/s write-only')
ms = t(lambda: x.zero_()); print(f'zero_ (memset) 1.6 GB: {ms:.3f} ms  {n*8/ms/1e6:.0f} GB/s write-only')
ms = t(lambda: y.copy_(x)); print(f'copy_ 1.6 GB: {ms:.3f} ms  {2*n*8/ms/1e6:.0f} GB/s read+write')
ms = t(lambda: x.sum()); print(f'sum   1.6 GB: {ms:.3f} ms  {n*8/ms/1e6:.0f} GB/s read-only')

#!/usr/bin/env python
"""What can the HBM system of this B200 sustain for write-heavy streams?  torch fill (pure write), copy
(1:1), and a 1-read : 9-write elementwise fan-out, timed with CUDA events.  Context for the Curve
Number kernel's 10 % read / 90 % write mix; not part of the product."""
import torch

def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

N = 36000 * 36000
dev = "cuda"
src = torch.randint(0, 255, (N,), dtype=torch.uint8, device=dev)
dst = torch.empty((9, N), dtype=torch.uint8, device=dev)
ms = t(lambda: dst.fill_(7)); print(f"fill 9 planes (pure write, {9*N/1e9:.2f} GB): {ms:.3f} ms  {9*N/ms/1e6:.0f} GB/s")
ms = t(lambda: dst[0].fill_(7)); print(f"fill 1 plane: {ms:.3f} ms  {N/ms/1e6:.0f} GB/s")
ms = t(lambda: dst[0].copy_(src)); print(f"copy 1 plane (1:1): {ms:.3f} ms  {2*N/ms/1e6:.0f} GB/s")
big = torch.empty(4 * N, dtype=torch.uint8, device=dev); big2 = torch.empty(4 * N, dtype=torch.uint8, device=dev)
ms = t(lambda: big2.copy_(big)); print(f"copy 5.2 GB (1:1): {ms:.3f} ms  {8*N/ms/1e6:.0f} GB/s")
a16 = torch.empty(2 ** 30, dtype=torch.bfloat16, device=dev); b16 = torch.empty(2 ** 30, dtype=torch.bfloat16, device=dev)
ms = t(lambda: b16.copy_(a16)); print(f"copy 1 Gi bf16 (the MEASURED_PEAKS recipe): {ms:.3f} ms  {4*2**30/ms/1e6:.0f} GB/s")
ms = t(lambda: torch.add(src.view(1, N).expand(9, N), 1, out=dst)); print(f"1 read : 9 write broadcast add: {ms:.3f} ms  {10*N/ms/1e6:.0f} GB/s")
ms = t(lambda: big.zero_()); print(f"memset 5.2 GB: {ms:.3f} ms  {4*N/ms/1e6:.0f} GB/s")

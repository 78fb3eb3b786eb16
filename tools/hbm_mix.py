#!/usr/bin/env python
"""Developer aid: what does a perfectly coalesced kernel reach on this box for pure copy, pure fill and a 26 % read /
74 % write mix (the 14x14 ROIAlign's traffic shape)? Puts the ROIAlign roofline fraction in context."""
import torch

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

MB = 1 << 20
src = torch.empty(512 * MB // 4, device="cuda").normal_()
dst = torch.empty(512 * MB // 4, device="cuda")
ms = t(lambda: dst.copy_(src)); print(f"copy 512 MB -> 512 MB   : {2*512*MB/ms/1e6:8.1f} GB/s")
ms = t(lambda: dst.fill_(1.0)); print(f"fill 512 MB              : {512*MB/ms/1e6:8.1f} GB/s")
ms = t(lambda: src.sum());      print(f"read 512 MB (sum)        : {512*MB/ms/1e6:8.1f} GB/s")
# 26/74 mix: out[0:400MB] = in[0:140MB] repeated (expand reads each source element ~2.9 times, mostly from L2/L1)
a = src[: 140 * MB // 4]
o = dst[: 3 * 140 * MB // 4].view(3, -1)
ms = t(lambda: o.copy_(a.expand(3, -1))); print(f"read 140 MB, write 420 MB: {(140+420)*MB/ms/1e6:8.1f} GB/s")

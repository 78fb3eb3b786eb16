"""Developer aid: what the HBM of this box sustains for pure writes, pure reads and copies (torch ops, CUDA events)."""
import torch

def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

N = 512 * 1024 * 1024 // 4
x = torch.empty(N, dtype=torch.float32, device="cuda")
y = torch.empty(N, dtype=torch.float32, device="cuda")
ms = t(lambda: x.fill_(1.0)); print(f"fill 512MB      {ms:.4f} ms  {0.512*1.048576/ms*1e3:.0f} GB/s (write only)")
ms = t(lambda: x.sum());      print(f"sum  512MB      {ms:.4f} ms  {0.512*1.048576/ms*1e3:.0f} GB/s (read only)")
ms = t(lambda: y.copy_(x));   print(f"copy 512MB      {ms:.4f} ms  {2*0.512*1.048576/ms*1e3:.0f} GB/s (read+write)")
q = N // 4
ms = t(lambda: (x[:q].copy_(y[:q]), x[q:].fill_(2.0)));  print(f"copy 128MB + fill 384MB  {ms:.4f} ms  {(q*8+3*q*4)/ms/1e6:.0f} GB/s (25% read / 75% write bytes: 1/5 vs 4/5)")

"""Probe the practical HBM ceiling for the trace sweep's access pattern: in-place read-modify-write
of a [B][A*F] fp32 buffer (what K3 does) vs an out-of-place copy (what MEASURED_PEAKS.json times)."""
import torch

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for mb in (335, 1342, 2684):
    n = mb * 1024 * 1024 // 4
    a = torch.randn(n, device="cuda"); b = torch.empty_like(a)
    ms = t(lambda: a.mul_(0.999))
    print(f"in-place mul_   {mb:5d} MB: {ms:.4f} ms  {2*n*4/ms/1e6:.0f} GB/s (read+write)")
    ms = t(lambda: b.copy_(a))
    print(f"copy_           {mb:5d} MB: {ms:.4f} ms  {2*n*4/ms/1e6:.0f} GB/s (read+write)")
    ms = t(lambda: torch.mul(a, 0.999, out=b))
    print(f"out-of-place mul{mb:5d} MB: {ms:.4f} ms  {2*n*4/ms/1e6:.0f} GB/s (read+write)")
    del a, b

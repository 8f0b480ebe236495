"""Developer micro-benchmark: the per-step split conv_in kernel (d3_conv_in_direct, out16 mode) and the per-loop
conditioning term (out32 mode) at B=16, 384x640."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, H, W = 16, 384, 640
x = torch.randn(B, 1, H, W, device="cuda")
c = torch.randn(B, 7, H, W, device="cuda")
w = torch.randn(128, 8, 3, 3, device="cuda") / 8.5
bias = torch.zeros(128, device="cuda")
wx, wc = ops.pack_conv_in_direct(w, 0, 1), ops.pack_conv_in_direct(w, 1, 8)
base = ops.conv_in_direct(c, wc, bias=bias)
out = torch.empty(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t = timeit(lambda: ops.conv_in_direct(x, wx, base=base, out_dtype=torch.bfloat16, out=out, stats=stats))
by = base.numel() * 4 + out.numel() * 2 + x.numel() * 4
print(f"per-step conv_in (Cs=1): {t:.3f} ms, {by / t / 1e6:.0f} GB/s algorithmic")
t = timeit(lambda: ops.conv_in_direct(c, wc, bias=bias, out=base), n=5)
print(f"per-loop conditioning term (Cc=7): {t:.3f} ms")

"""Developer micro-benchmark: the GroupNorm+SiLU pass at the UNet's big shapes (single source 128 ch, concat 128+128)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, H, W = 16, 384, 640
x = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
x2 = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
st = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
st[:, :, 1] = 4.0 * H * W
g, b = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
g2, b2 = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
y = torch.empty_like(x)
y2 = torch.empty(B, H, W, 256, device="cuda", dtype=torch.bfloat16)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t = timeit(lambda: ops.groupnorm_silu(x, st, g, b, out=y))
print(f"GN 128ch: {t * 1e3:.0f} us, {2 * x.numel() * 2 / t / 1e6:.0f} GB/s")
t = timeit(lambda: ops.groupnorm_silu(x, st, g2, b2, x2=x2, stats2=st, out=y2))
print(f"GN 128+128 concat: {t * 1e3:.0f} us, {2 * y2.numel() * 2 / t / 1e6:.0f} GB/s")

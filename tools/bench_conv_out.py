"""Developer micro-benchmark: fused GroupNorm + SiLU + conv_out (128 -> dc) at B=16, 384x640."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, H, W = 16, 384, 640
for dc in (1, 3):
    x = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
    o = x.double().reshape(B, H * W, 32, 4)
    stats = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1).contiguous()
    gamma, beta = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    w = torch.randn(dc, 9, 128, device="cuda") / 34.0
    bias = torch.zeros(dc, device="cuda")
    out = torch.empty(B, dc, H, W, device="cuda")
    for _ in range(3):
        ops.conv_out_fused(x, stats, gamma, beta, w, bias, dc, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.conv_out_fused(x, stats, gamma, beta, w, bias, dc, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = x.numel() * 2 + out.numel() * 4
    print(f"conv_out dc={dc}: {ms:.3f} ms, {by / ms / 1e6:.0f} GB/s algorithmic")

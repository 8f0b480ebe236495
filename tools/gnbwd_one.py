"""Developer tool: GroupNorm+SiLU backward on one level-0 tensor (16 x 384 x 640 x 128), for ncu captures."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, H, W, C = 16, 384, 640, int(sys.argv[1]) if len(sys.argv) > 1 else 128
dt = torch.bfloat16
x = torch.randn(B, H, W, C, device="cuda").to(dt)
dy = torch.randn(B, H, W, C, device="cuda").to(dt)
v = x.double().view(B, H * W, C // 4, 4)
stats = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
del v
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(2):
    ops.groupnorm_silu_bwd(x, stats, gamma, beta, dy, dgamma=dg, dbeta=db)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.groupnorm_silu_bwd(x, stats, gamma, beta, dy, dgamma=dg, dbeta=db)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"groupnorm_silu_bwd C={C} @{B}x{H}x{W}: {ms:.3f} ms, {5 * x.numel() * 2 / ms / 1e6:.0f} GB/s over 5 N bytes")

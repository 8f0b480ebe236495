"""Developer tool: conv_out backward at the bench shape (16 x 384 x 640), for ncu captures."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, H, W = 16, 384, 640
h = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
v = h.double().view(B, H * W, 32, 4)
stats = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
del v
gamma, beta = torch.rand(128, device="cuda") + 0.5, torch.randn(128, device="cuda") * 0.1
w = torch.randn(1, 9, 128, device="cuda") * 0.1
dy = torch.randn(B, 1, H, W, device="cuda")
dw = torch.zeros_like(w)
for _ in range(3):
    ops.conv_out_bwd(h, stats, gamma, beta, w, dy, dw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv_out_bwd(h, stats, gamma, beta, w, dy, dw)
e1.record()
torch.cuda.synchronize()
print(f"conv_out_bwd: {e0.elapsed_time(e1) / 5:.3f} ms")

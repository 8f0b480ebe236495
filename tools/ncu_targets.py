"""Developer tool: launch the three kernels worth an `ncu --set full` capture a few times each (dominant conv shape,
GroupNorm+SiLU pass, attention) at the BASELINE batch (B=16).  usage: python tools/ncu_targets.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, H, W = 16, 384, 640
x = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
w = (torch.randn(128, 1152, device="cuda") / 34.0).to(torch.bfloat16)
bias = torch.zeros(128, device="cuda")
out = torch.empty_like(x)
stats = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
gamma, beta = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
qkv = torch.randn(B, 24, 40, 1536, device="cuda").to(torch.bfloat16)
x256 = torch.randn(B, 96, 160, 256, device="cuda").to(torch.bfloat16)
w256 = (torch.randn(256, 2304, device="cuda") / 48.0).to(torch.bfloat16)
b256 = torch.zeros(256, device="cuda")
for _ in range(4):
    ops.conv2d_igemm(x, w, bias, ksize=3, out=out, stats=stats)
    ops.groupnorm_silu(x, stats, gamma, beta, out=out)
    ops.attention_hd8(qkv, 512)
    ops.conv2d_igemm(x256, w256, b256, ksize=3)
torch.cuda.synchronize()
print("done")

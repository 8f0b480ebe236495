"""ncu target: the tcgen05 head_dim-64 attention at the latent UNet's level-0 shape (B = 16, N = 3600, 5 heads)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

B, hw, heads = 16, (45, 80), 5
C = 64 * heads
qkv = (torch.randn(B, hw[0], hw[1], 3 * C, device="cuda") * 1.5).half()
for _ in range(4):
    o = ops.attention_hd64_tc(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))

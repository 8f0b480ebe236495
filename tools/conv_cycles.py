"""Developer tool: per-role cycle breakdown of the halo-block conv kernel (needs D3_CONV_DEBUG_SKIP to include 256)."""
import ctypes
import os
import sys
from pathlib import Path

os.environ["D3_CONV_DEBUG_SKIP"] = str(int(os.environ.get("D3_CONV_DEBUG_SKIP", "0")) | 256)
import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import _lib, ops  # noqa: E402

B, H, W, Cin, Cout = 16, 384, 640, 128, 128
x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
wp = (torch.randn(Cout, 9 * Cin, device="cuda") / (9 * Cin) ** 0.5).to(torch.bfloat16)
bias = torch.zeros(Cout, device="cuda")
out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(B, Cout // 4, 2, dtype=torch.float64, device="cuda")
L = _lib.lib()
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    ops.conv2d_igemm(x, wp, bias, ksize=3, out=out, stats=stats)
torch.cuda.synchronize()
L.d3_debug_conv_profile(buf, 1)
n = 10
for _ in range(n):
    ops.conv2d_igemm(x, wp, bias, ksize=3, out=out, stats=stats)
torch.cuda.synchronize()
L.d3_debug_conv_profile(buf, 0)
v = [int(b) for b in buf]
tiles = v[7]
print(f"tiles {tiles}; per tile (cycles): MMA-thread total {v[0] / tiles:.0f}, wait acc {v[1] / tiles:.0f}, "
      f"wait blocks {v[2] / tiles:.0f}, wait weights {v[3] / tiles:.0f} | epilogue total {v[4] / tiles:.0f}, "
      f"wait tfull {v[5] / tiles:.0f}, passes {v[6] / tiles:.0f} = buffer wait + bar1 {v[8] / tiles:.0f}, "
      f"TMEM load {v[9] / tiles:.0f}, math + stores {v[10] / tiles:.0f}, fence + bar2 + issue {v[11] / tiles:.0f}")

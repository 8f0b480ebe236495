"""Developer tool: the dominant weight-gradient shape (3x3 128 -> 128 @ 16 x 384 x 640) launched a few times, for ncu.
    ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -c 1 -o out python tools/wgrad_one.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

dtype = torch.float16 if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else torch.bfloat16
B, H, W, C = 16, 384, 640, 128
x = torch.randn(B, H, W, C, device="cuda").to(dtype)
dy = torch.randn(B, H, W, C, device="cuda").to(dtype)
dw = torch.zeros(C, 9 * C, device="cuda")
for _ in range(3):
    ops.conv2d_wgrad(x, dy, dw, ksize=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv2d_wgrad(x, dy, dw, ksize=3)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"conv2d_wgrad 3x3 128->128 @{B}x{H}x{W}: {ms:.3f} ms, {2.0 * B * H * W * 9 * C * C / ms / 1e9:.0f} TFLOP/s")

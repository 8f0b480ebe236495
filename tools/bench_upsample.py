"""Developer micro-benchmark: the 4 sub-pixel phase convs of Upsample2D at the UNet's shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

for (B, H, W, C) in [(16, 192, 320, 128), (16, 96, 160, 256), (16, 48, 80, 256), (16, 24, 40, 512), (16, 12, 20, 512)]:
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(C, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    packs = ops.pack_upsample_weights(w)
    bias = torch.zeros(C, device="cuda")
    out = torch.empty(B, 2 * H, 2 * W, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, C // 4, 2, dtype=torch.float64, device="cuda")

    def run():
        for ph in range(4):
            ops.conv2d_igemm(x, packs[ph], bias, ksize=3, out=out, stats=stats, up2x_phase=ph + 1)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4 * 2.0 * B * H * W * C * 4 * C
    print(f"upsample {H}x{W} C={C}: {ms * 1e3:.0f} us for 4 phases, {fl / ms / 1e9:.0f} TF")
    # plain 3x3 conv on the same input for comparison (K = 9C, 2.25x the flops of one phase)
    wp = ops.pack_conv_weight(w)
    o2 = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv2d_igemm(x, wp, bias, ksize=3, out=o2, stats=stats)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ops.conv2d_igemm(x, wp, bias, ksize=3, out=o2, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"   plain 3x3 same input: {ms * 1e3:.0f} us, {2.0 * B * H * W * C * 9 * C / ms / 1e9:.0f} TF")

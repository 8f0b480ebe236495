"""Developer micro-benchmark: time d3_conv2d_igemm on the dominant UNet shapes (CUDA events)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

SHAPES = [
    # B, H, W, Cin, Cout, k, stride, C1, C2
    (16, 384, 640, 128, 128, 3, 1, 0, 0),
    (16, 384, 640, 256, 128, 3, 1, 128, 128),
    (16, 192, 320, 128, 128, 3, 1, 0, 0),
    (16, 96, 160, 256, 256, 3, 1, 0, 0),
    (16, 48, 80, 256, 256, 3, 1, 0, 0),
    (16, 24, 40, 512, 512, 3, 1, 0, 0),
    (16, 12, 20, 512, 512, 3, 1, 0, 0),
    (16, 24, 40, 512, 1536, 1, 1, 0, 0),
]


def main():
    B_override = int(sys.argv[1]) if len(sys.argv) > 1 else None
    for (B, H, W, Cin, Cout, k, s, C1, C2) in SHAPES:
        if B_override:
            B = B_override
        x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
        Kt = k * k * Cin + C1 + C2
        wp = (torch.randn(Cout, Kt, device="cuda") / Kt ** 0.5).to(torch.bfloat16)
        bias = torch.zeros(Cout, device="cuda")
        sc1 = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16) if C1 else None
        sc2 = torch.randn(B, H, W, C2, device="cuda").to(torch.bfloat16) if C2 else None
        out = torch.empty(B, H // s, W // s, Cout, device="cuda", dtype=torch.bfloat16)
        stats = torch.zeros(B, Cout // 4, 2, dtype=torch.float64, device="cuda")
        for _ in range(3):
            ops.conv2d_igemm(x, wp, bias, ksize=k, stride=s, sc1=sc1, sc2=sc2, out=out, stats=stats)
        torch.cuda.synchronize()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.conv2d_igemm(x, wp, bias, ksize=k, stride=s, sc1=sc1, sc2=sc2, out=out, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 2.0 * B * (H // s) * (W // s) * Cout * Kt
        print(f"B={B} {H}x{W} Cin={Cin} Cout={Cout} k={k} s={s} sc={C1}+{C2}: {ms:.3f} ms  "
              f"{flops / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()

"""Developer micro-benchmark: d3_attention_hd8 at the UNet's shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

for (B, H, W) in ((16, 24, 40), (16, 12, 20), (4, 34, 60)):
    qkv = torch.randn(B, H, W, 1536, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops.attention_hd8(qkv, 512)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ops.attention_hd8(qkv, 512)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    N = H * W
    pairs = B * 64 * N * N
    print(f"B={B} N={N}: {ms:.3f} ms  {pairs / ms / 1e6:.1f} G score/s  ({4 * 8 * pairs / ms / 1e9:.1f} TFLOP/s)")

"""Developer tool: run ONE eager UNet forward (B x 8 x 384 x 640) after a warm-up forward, for ncu launch lists.
usage: python tools/profile_unet.py [B] [dtype]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
net = d.UNet2DModel(sample_size=(384, 640), in_channels=8, out_channels=1, compute_dtype=dtype).to("cuda")
net.load_state_dict(random_init_state_dict(8, 1))
x = torch.randn(B, 1, 384, 640, device="cuda")
c = torch.randn(B, 7, 384, 640, device="cuda")
for i in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = net.forward_split(x, c, 73)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward {i}: {e0.elapsed_time(e1):.2f} ms", flush=True)

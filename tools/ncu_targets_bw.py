"""Developer tool: launch the bandwidth-side kernels worth an `ncu --set full` capture (split conv_in, streaming conv_out,
single / dual GroupNorm, scheduler step, guidance) at the BASELINE batch (B=16, 384x640)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import ops  # noqa: E402

B, H, W = 16, 384, 640
x = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
o = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
o[:, :, 1] = 4.0 * H * W
g, b = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
g2, b2 = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
y = torch.empty_like(x)
cat = torch.empty(B, H, W, 256, device="cuda", dtype=torch.bfloat16)
xs = torch.randn(B, 1, H, W, device="cuda")
cond = torch.randn(B, 7, H, W, device="cuda")
w_in = torch.randn(128, 8, 3, 3, device="cuda") / 8.5
wx, wc = ops.pack_conv_in_direct(w_in, 0, 1), ops.pack_conv_in_direct(w_in, 1, 8)
bias = torch.zeros(128, device="cuda")
base = ops.conv_in_direct(cond, wc, bias=bias)
w_out = torch.randn(1, 9, 128, device="cuda") / 34.0
stats = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
sched = d.MyDDIMScheduler.from_config(dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128,
                                           prediction_type="sample", clip_sample=True, set_alpha_to_one=False,
                                           steps_offset=1))
sched.set_timesteps(10)
fg = d.FlowGuidance(flow_guidance_weight=1.0, flow_guidance_mode="stereo")
left, right = torch.rand(B, 3, H, W, device="cuda") * 2 - 1, torch.rand(B, 3, H, W, device="cuda") * 2 - 1
for _ in range(3):
    ops.conv_in_direct(xs, wx, base=base, out_dtype=torch.bfloat16, out=y, stats=stats)
    ops.groupnorm_silu(x, o, g, b, out=y)
    ops.groupnorm_silu_dual(x, o, g, b, g2[128:], b2[128:], cat[..., 128:], groups_b=16, out_a=y)
    ops.conv_out_fused(x, o, g, b, w_out, torch.zeros(1, device="cuda"), 1)
    sched.step(xs, sched.timesteps[0], xs)
    fg.grad(xs, left, right)
torch.cuda.synchronize()
print("done")

"""Developer tool: where the time of one 10-step sampling loop goes (B=16, 384x640): CUPTI kernel totals vs wall."""
import sys
import time
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = 16
net = d.UNet2DModel(sample_size=(384, 640), in_channels=8, out_channels=1, compute_dtype="bf16").to("cuda")
net.load_state_dict(random_init_state_dict(8, 1))
cfg = dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128, prediction_type="sample", clip_sample=True,
           set_alpha_to_one=False, steps_offset=1)
pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(cfg), d.FlowGuidance(flow_guidance_weight=0.0))
left = torch.rand(B, 3, 384, 640, device="cuda") * 2 - 1
right = torch.rand(B, 3, 384, 640, device="cuda") * 2 - 1
raw = torch.rand(B, 1, 384, 640, device="cuda") * 2 - 1
gens = [torch.Generator().manual_seed(i) for i in range(B)]


def once():
    return pipe(None, left, right, raw, None, None, num_inference_steps=10, num_intermediate_images=1,
                depth_channels=1, cond_channels="left+right+raw", generator=gens)


for _ in range(3):
    once()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    once()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3 * 1e3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    once()
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]
        tot[name][0] += ev.device_time_total / 1e3
        tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"wall per 10-step loop: {wall:.1f} ms; GPU kernel+copy time: {total:.1f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {v[0]:8.3f} ms  x{v[1]:5d}  {k}")

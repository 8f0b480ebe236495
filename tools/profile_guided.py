"""Developer tool: kernel totals of one guided (rgb+raw, stereo guidance) 10-step loop at B=8."""
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = 8
net = d.UNet2DModel(sample_size=(384, 640), in_channels=5, out_channels=1, compute_dtype="bf16").to("cuda")
net.load_state_dict(random_init_state_dict(5, 1))
cfg = dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128, prediction_type="sample", clip_sample=True,
           set_alpha_to_one=False, steps_offset=1)
pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(cfg),
                                 d.FlowGuidance(flow_guidance_weight=1.0, flow_guidance_mode="stereo"))
rgb, left, right = (torch.rand(B, 3, 384, 640, device="cuda") * 2 - 1 for _ in range(3))
raw = torch.rand(B, 1, 384, 640, device="cuda") * 2 - 1
gens = [torch.Generator().manual_seed(i) for i in range(B)]
normer = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
once = lambda: pipe(rgb, left, right, raw, None, None, num_inference_steps=10, num_intermediate_images=1,  # noqa: E731
                    depth_channels=1, cond_channels="rgb+raw", denorm=normer.denormalize, generator=gens)
for _ in range(2):
    once()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    once()
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]
        tot[name][0] += ev.device_time_total / 1e3
        tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"GPU time per guided loop: {total:.1f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    if "igemm" in k or "gn_apply" in k or "attention" in k:
        continue
    print(f"  {v[0]:8.3f} ms  x{v[1]:5d}  {k}")

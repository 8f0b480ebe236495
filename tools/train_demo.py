"""Developer tool / demonstration: train the pixel-space denoiser with d3roma_b200.training on a synthetic task, then
sample with the trained weights through the inference pipeline (the loop of train.py:560-660 in miniature: train_step ...
-> validation with GuidedDiffusionPipeline on the weights being trained).

Task: the clean disparity is a smooth function of the conditioning (a blurred copy of the raw channel plus a left-image
term), so that a network that uses its conditioning can drive the 'sample'-prediction loss towards zero.
    python tools/train_demo.py [steps] [batch] [H] [W]"""
import sys
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import training  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
W = int(sys.argv[4]) if len(sys.argv) > 4 else 160
dev = torch.device("cuda", 0)


def make_batch(gen, n):
    left = torch.rand(n, 3, H, W, generator=gen) * 2 - 1
    right = torch.rand(n, 3, H, W, generator=gen) * 2 - 1
    raw = F.avg_pool2d(torch.rand(n, 1, H, W, generator=gen) * 2 - 1, 9, 1, 4)
    clean = (0.8 * F.avg_pool2d(raw, 5, 1, 2) * 3 + 0.2 * F.avg_pool2d(left.mean(1, keepdim=True), 5, 1, 2)).clamp(-1, 1)
    return {k: v.to(dev) for k, v in dict(left=left, right=right, raw=raw, clean=clean).items()}


cfg = SimpleNamespace(ldm=False, loss_type="mse", noise_strategy="pyramid", cond_channels="left+right+raw",
                      prediction_type="sample", noise_rgb=False, depth_channels=1, clip_sample=True, clip_sample_range=1.0,
                      divis_by=32, clip_grad_norm=1.0)
sched_cfg = dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample", clip_sample=True)
tr = training.UNetTrainer(8, 1, compute_dtype="bf16", lr=2e-4, device=dev)
tr.load_state_dict(random_init_state_dict(8, 1, seed=0))
noise_scheduler = d.MyDDPMScheduler.from_config(sched_cfg)


class Uniform:                                    # core/resample.py UniformSampler
    def sample(self, n, device):
        return torch.randint(0, 128, (n,)).to(device), torch.ones(n, device=device)


def validate(tag):
    net = d.UNet2DModel(sample_size=(H, W), in_channels=8, out_channels=1, compute_dtype="fp16").to(dev)
    net.load_state_dict(tr.state_dict())
    ddim = d.MyDDIMScheduler.from_config(dict(sched_cfg, set_alpha_to_one=False, steps_offset=1))
    pipe = d.GuidedDiffusionPipeline(net, ddim, d.FlowGuidance(flow_guidance_weight=0.0))
    b = make_batch(torch.Generator().manual_seed(999), 4)
    out = pipe(None, b["left"], b["right"], b["raw"], None, None, num_inference_steps=10, num_intermediate_images=1,
               cond_channels="left+right+raw", generator=torch.Generator().manual_seed(1)).images
    err = float((out - b["clean"]).abs().mean())
    print(f"[{tag}] DDIM-10 sampling with the trainer's weights: mean |x0 - target| = {err:.4f}")
    return err


torch.manual_seed(0)
e0 = validate("before training")
gen = torch.Generator().manual_seed(1)
for it in range(steps):
    b = make_batch(gen, B)
    loss, info = training.train_step(None, cfg, tr, None, None, noise_scheduler, Uniform(), None, normalized_disp=b["clean"],
                                     mask=torch.ones_like(b["clean"]), left_image=b["left"], right_image=b["right"],
                                     sim_disp=b["raw"])
    if it % 10 == 0 or it == steps - 1:
        print(f"step {it:3d}: loss {loss:.5f}  grad-norm {info['gradn']:.3f}")
e1 = validate("after training")
print(f"validation error {e0:.4f} -> {e1:.4f}")
assert e1 < e0

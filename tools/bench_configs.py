"""Developer benchmark: the single-GPU share of each BASELINE.json config (device-resident inputs, CUDA events).
  C1  1 pair 360x640, left+right+raw, DDIM 10 (latency)            C2  16 pairs (the bench.py workload)
  C3  8 pairs/GPU rgb+raw, stereo guidance on, DDIM 10             C4  540x960, DDPM 100 steps, fp32 path, 1 pair
  C5  32 pairs/GPU left+right+raw, DDIM 10
usage: bench_configs.py [c1 c2 c3 c4 c5]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import synthetic  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

SCHED = dict(beta_schedule="squaredcos_cap_v2", beta_start=1e-4, beta_end=2e-2, num_train_timesteps=128,
             prediction_type="sample", clip_sample=True, set_alpha_to_one=False, steps_offset=1)


def run(name, B, hw, cond_channels, cin, dc, steps, sched_cls, dtype, guidance_w=0.0, iters=3, extra=None):
    h, w = hw
    hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    net = d.UNet2DModel(sample_size=(hp, wp), in_channels=cin, out_channels=dc, compute_dtype=dtype).to("cuda")
    net.load_state_dict(random_init_state_dict(cin, dc, seed=0))
    cfg = dict(SCHED)
    if extra:
        cfg.update(extra)
    if sched_cls is d.MyDDPMScheduler:
        for k in ("set_alpha_to_one", "steps_offset"):
            cfg.pop(k, None)
    mode = "stereo" if guidance_w > 0 else "imputation"
    pipe = d.GuidedDiffusionPipeline(net, sched_cls.from_config(cfg),
                                     d.FlowGuidance(flow_guidance_weight=guidance_w, flow_guidance_mode=mode))
    batch = synthetic.make_batch(list(range(B)), h=h, w=w, device="cuda")
    padder = d.InputPadder(batch["left_image"].shape, divis_by=32)
    rgb, left, right, sim, raw = padder.pad(batch["normalized_rgb"], batch["left_image"], batch["right_image"],
                                            batch["sim_disp"], batch["raw_depth"])
    gens = [torch.Generator().manual_seed(100 + i) for i in range(B)]

    normer = d.Normalizer(ssi=False, mode="average", num_chs=dc, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)

    def once():
        return pipe(rgb, left, right, sim, raw, None, num_inference_steps=steps, num_intermediate_images=1,
                    depth_channels=dc, cond_channels=cond_channels, denorm=normer.denormalize, generator=gens)
    once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: B={B} {h}x{w} {cond_channels} dc={dc} {steps} steps {dtype}: {ms:.1f} ms per batch, "
          f"{ms / steps:.2f} ms per UNet step, {B / ms * 1e3:.2f} maps/s, peak mem "
          f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net, pipe
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


which = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which:
    run("C1", 1, (360, 640), "left+right+raw", 8, 1, 10, d.MyDDIMScheduler, "bf16", iters=10)
if "c2" in which:
    run("C2", 16, (360, 640), "left+right+raw", 8, 1, 10, d.MyDDIMScheduler, "bf16")
if "c3" in which:
    run("C3", 8, (360, 640), "rgb+raw", 5, 1, 10, d.MyDDIMScheduler, "bf16", guidance_w=1.0)
if "c4" in which:
    run("C4", 1, (540, 960), "left+right+raw", 8, 1, 100, d.MyDDPMScheduler, "fp32", iters=1,
        extra=dict(thresholding=True, variance_type="fixed_small"))
    run("C4-fp16", 4, (540, 960), "left+right+raw", 8, 1, 100, d.MyDDPMScheduler, "fp16", iters=1,
        extra=dict(thresholding=True, variance_type="fixed_small"))
if "c5" in which:
    run("C5", 32, (360, 640), "left+right+raw", 8, 1, 10, d.MyDDIMScheduler, "bf16")

"""Developer tool: one small pass over every kernel family in bf16, fp16 and fp32 (eager launches, guidance on) - a quick
smoke run to put under a debugger or D3_SYNC_CHECK=1."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import evaluate, synthetic  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

cfg = dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128, prediction_type="sample", clip_sample=True,
           set_alpha_to_one=False, steps_offset=1)
for dtype, mode, w in (("bf16", "stereo", 1.0), ("fp16", "imputation", 0.0), ("fp32", "imputation", 0.0)):
    net = d.UNet2DModel(sample_size=(64, 96), in_channels=8, out_channels=1, compute_dtype=dtype).to("cuda")
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(cfg),
                                     d.FlowGuidance(flow_guidance_weight=w, flow_guidance_mode=mode))
    pipe.use_cuda_graph = False
    batch = synthetic.make_batch([0, 1, 2], h=60, w=90)
    keys = ("fxb", "normalized_rgb", "raw_disp", "mask", "left_image", "right_image", "sim_disp", "raw_depth")
    c = evaluate.default_config(num_inference_timesteps=2, num_inference_rounds=1)
    pred, metrics, unc, err, out = evaluate.eval_batch(c, pipe, True, **{k: batch[k] for k in keys})
    torch.cuda.synchronize()
    print(dtype, "ok", float(abs(pred).max()))

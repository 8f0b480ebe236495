"""torchrun demo/check of the sharded evaluation loop on real GPUs (NCCL): every rank evaluates its shard of a small
synthetic sweep, the fp64 metric rows are all-gathered, and rank 0 checks them against evaluating the whole sweep
locally.  usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_eval_demo.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import distributed as D, evaluate, synthetic  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

rank, world, local = D.init_from_env("nccl")
dev = torch.device("cuda", local)
net = d.UNet2DModel(sample_size=(64, 96), in_channels=8, out_channels=1, compute_dtype="fp16").to(dev)
net.load_state_dict(random_init_state_dict(8, 1, seed=0))
cfg = dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128, prediction_type="sample", clip_sample=True,
           set_alpha_to_one=False, steps_offset=1)
pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(cfg), d.FlowGuidance(flow_guidance_weight=0.0))
c = evaluate.default_config(num_inference_timesteps=3)
mk = lambda idx: synthetic.make_batch(idx, h=60, w=90)  # noqa: E731
N = 11
mdict, rows = D.run_distributed_eval(c, pipe, N, 2, make_batch_fn=mk, seed=7)
if rank == 0:
    import torch.distributed as dist
    # same sweep evaluated by this rank alone (bypass the process group)
    keys = ("fxb", "normalized_rgb", "raw_disp", "mask", "left_image", "right_image", "sim_disp", "raw_depth")
    local_rows = []
    for s in range(0, N, 2):
        idx = list(range(s, min(s + 2, N)))
        b = mk(idx)
        _, r, _, _, _ = evaluate.eval_batch(c, pipe, True, generator=D.sample_generators(idx, 7), as_numpy=False,
                                            **{k: b[k] for k in keys})
        local_rows.append(r.cpu().numpy())
    ref = np.concatenate(local_rows)
    print(f"world {world}: gathered rows {rows.shape}, max |diff| vs single-rank evaluation = "
          f"{np.abs(rows - ref).max():.3e}; epe mean {mdict['disp']['epe'] if 'epe' in mdict['disp'] else list(mdict['disp'].items())[:2]}")
    assert rows.shape == (N, 11) and np.allclose(rows, ref, rtol=1e-6, atol=1e-9)
    print("DIST_EVAL_OK")
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()

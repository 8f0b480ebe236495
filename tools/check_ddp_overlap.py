"""Developer check (run under torchrun, >= 2 GPUs): data-parallel steps with the overlapped gradient all-reduce must give
the same parameters as with the single all-reduce after the backward, and the same on every rank."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.training import UNetTrainer  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sched = d.MyDDPMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample"))
B, H, W = 2, 64, 96
out = []
# the kernels accumulate with atomics, so two identical runs differ in the last bits: the run-to-run difference of two
# plain runs is the yardstick for the overlapped one. lr = 0 keeps the parameters fixed; the all-reduced gradient
# buffer of the last step is what is compared.
for overlap in (False, False, True):
    tr = UNetTrainer(8, 1, compute_dtype="bf16", lr=0.0, device=dev)
    tr.overlap_allreduce = overlap
    tr.load_state_dict(random_init_state_dict(8, 1, seed=0))
    g = torch.Generator().manual_seed(100 + rank)
    for _ in range(2):
        clean = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).to(dev)
        cond = (torch.rand(B, 7, H, W, generator=g) * 2 - 1).to(dev)
        noise = torch.randn(B, 1, H, W, generator=g).to(dev)
        t = torch.randint(0, 128, (B,), generator=g).to(dev)
        tr.train_step(clean, cond, noise, t, sched, world_size=world)
    assert tr.overlapped_reductions == (2 if overlap else 0)
    out.append(tr._g.clone())
scale = float(out[0].abs().max())
noise_floor = float((out[0] - out[1]).abs().max()) / scale
diff = float((out[2] - out[0]).abs().max()) / scale
ref = out[2].clone()
dist.broadcast(ref, 0)
across = float((out[2] - ref).abs().max())
print(f"rank {rank}: plain vs plain {noise_floor:.3e}, overlapped vs plain {diff:.3e} (of max |g| = {scale:.3e}); "
      f"overlapped vs rank 0 = {across:.3e}", flush=True)
assert diff <= 4 * noise_floor + 1e-6 and across == 0.0
dist.destroy_process_group()

"""Developer tool: a few training steps of UNetTrainer at the bench shape (for ncu launch lists / captures).
    python tools/train_one_step.py [steps] [batch] [dtype]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.training import UNetTrainer  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda", 0)
tr = UNetTrainer(8, 1, compute_dtype=dtype, loss_scale=65536.0 if dtype == "fp16" else 1.0, device=dev)
tr.load_state_dict(random_init_state_dict(8, 1, seed=0))
sched = d.MyDDPMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample"))
g = torch.Generator().manual_seed(0)
H, W = 384, 640
clean = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).to(dev)
cond = (torch.rand(B, 7, H, W, generator=g) * 2 - 1).to(dev)
noise = torch.randn(B, 1, H, W, generator=g).to(dev)
t = torch.randint(0, 128, (B,), generator=g).to(dev)
for i in range(steps):
    loss = tr.train_step(clean, cond, noise, t, sched)
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(loss):.5f} grad-norm {tr.grad_norm:.4f} peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

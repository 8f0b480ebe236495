"""Developer tool: head_dim-64 self-attention, mma.sync kernel vs the tcgen05 / TMEM kernel at the latent UNet's shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402


def t_of(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for B, hw, heads in ((16, (45, 80), 5), (16, (48, 80), 5), (16, (23, 40), 10), (16, (12, 20), 20), (16, (6, 10), 20), (1, (45, 80), 5)):
    C = 64 * heads
    N = hw[0] * hw[1]
    qkv = (torch.randn(B, hw[0], hw[1], 3 * C, device="cuda") * 1.5).half()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    fl = 4.0 * B * N * N * C
    t0 = t_of(lambda: ops.attention_hd64(q, k, v, heads))
    t1 = t_of(lambda: ops.attention_hd64_tc(q, k, v, heads))
    print(f"B={B} N={N} heads={heads}: mma.sync {t0:.3f} ms ({fl / t0 / 1e9:.0f} TF/s) | tcgen05 {t1:.3f} ms ({fl / t1 / 1e9:.0f} TF/s)")

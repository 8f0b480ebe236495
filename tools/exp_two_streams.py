"""Developer experiment: overlap the bandwidth-bound GroupNorm passes of one half of the batch with the tensor-bound
convolutions of the other half by running the UNet forward as two half-batch forwards on two streams (one CUDA graph
with a fork / join).  GroupNorm blocks (no shared memory, ~10 K registers) can share an SM with a persistent conv CTA.
Prints sustained ms per 16-sample forward for: one B=16 graph, two sequential B=8 forwards, two concurrent B=8 forwards."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.unet import _StatsArena  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = d.UNet2DModel(sample_size=(384, 640), in_channels=8, out_channels=1).to("cuda")
net.load_state_dict(random_init_state_dict(8, 1))
x = torch.randn(B, 1, 384, 640, device="cuda")
c = torch.randn(B, 7, 384, 640, device="cuda")
t = torch.full((1,), 73.0, device="cuda")
h = B // 2
xa, xb, ca, cb = x[:h].contiguous(), x[h:].contiguous(), c[:h].contiguous(), c[h:].contiguous()
base = net.cond_base(c, 1, out=torch.empty((B, 384, 640, 128), dtype=torch.float16, device="cuda"))
ba, bb = base[:h].contiguous(), base[h:].contiguous()
ar0, ar1, ar2 = (_StatsArena(torch.device("cuda"), B * 2 * 40000) for _ in range(3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def full():
    return net.forward_split(x, c, t, cond_base=base, arena=ar0)


def seq():
    return net.forward_split(xa, ca, t, cond_base=ba, arena=ar1), net.forward_split(xb, cb, t, cond_base=bb, arena=ar2)


def conc():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        ya = net.forward_split(xa, ca, t, cond_base=ba, arena=ar1)
    with torch.cuda.stream(s2):
        yb = net.forward_split(xb, cb, t, cond_base=bb, arena=ar2)
    cur.wait_stream(s1)
    cur.wait_stream(s2)
    return ya, yb


def sustained(fn, seconds=2.0):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) / 5
    n = max(10, int(seconds * 1e3 / burst))
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return burst, e0.elapsed_time(e1) / n, out


bf, sf, yf = sustained(full)
print(f"one B={B} forward:              burst {bf:.2f} ms, sustained {sf:.2f} ms")
bs, ss, _ = sustained(seq)
print(f"two sequential B={h} forwards:   burst {bs:.2f} ms, sustained {ss:.2f} ms")
bc, sc, (ya, yb) = sustained(conc)
print(f"two concurrent B={h} forwards:   burst {bc:.2f} ms, sustained {sc:.2f} ms")
err = (torch.cat([ya, yb]) - yf).abs().max().item()
print(f"max |concurrent - full| = {err:.3e}")

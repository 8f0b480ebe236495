"""Developer tool: per-layer table of ONE eager UNet forward - every library call is bracketed by CUDA events and
reported with its shape, time, and algorithmic TFLOP/s (convs) or GB/s (bandwidth kernels).
usage: python tools/profile_layers.py [B] [dtype] [H] [W]"""
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import ops  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
H = int(sys.argv[3]) if len(sys.argv) > 3 else 384
W = int(sys.argv[4]) if len(sys.argv) > 4 else 640
records = []


def wrap(name, describe):
    orig = getattr(ops, name)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(*a, **k)
        e1.record()
        records.append((name, describe(a, k, out), e0, e1))
        return out
    setattr(ops, name, f)


def conv_desc(a, k, out):
    x, w = a[0], a[1]
    Bn, Hh, Ww, Cin = x.shape
    Cout, Kt = w.shape
    s = k.get("stride", 1)
    ph = k.get("up2x_phase", 0)
    Ho, Wo = (Hh // s, Ww // s)
    flops = 2.0 * Bn * Ho * Wo * Cout * Kt
    tag = f"k{k['ksize']}s{s}" + (f" up{ph}" if ph else "") + (" +sc" if k.get("sc1") is not None else "") + \
        (" +res" if k.get("residual") is not None else "")
    return (f"{Hh}x{Ww} {Cin}->{Cout} K={Kt} {tag}", flops, 0.0)


def gn_desc(a, k, out):
    x = a[0]
    n = out.numel() if torch.is_tensor(out) else x.numel()
    return (f"{tuple(x.shape[1:3])} C={out.shape[-1]}", 0.0, 2.0 * n * x.element_size())


def gn2_desc(a, k, out):
    x = a[0]
    return (f"{tuple(x.shape[1:3])} C={x.shape[-1]} dual", 0.0, 3.0 * x.numel() * x.element_size())


def attn_desc(a, k, out):
    q = a[0]
    Bn, Hh, Ww, _ = q.shape
    N, C = Hh * Ww, a[1]
    return (f"N={N} C={C}", 4.0 * Bn * N * N * C, 0.0)


def misc_desc(a, k, out):
    return ("", 0.0, 0.0)


wrap("conv2d_igemm", conv_desc)
wrap("groupnorm_silu", gn_desc)
wrap("groupnorm_silu_dual", gn2_desc)
wrap("attention_hd8", attn_desc)
for nm in ("conv_in_direct", "conv_out_fused", "time_embed", "temb_project_all"):
    wrap(nm, misc_desc)

net = d.UNet2DModel(sample_size=(H, W), in_channels=8, out_channels=1, compute_dtype=dtype).to("cuda")
net.load_state_dict(random_init_state_dict(8, 1))
x = torch.randn(B, 1, H, W, device="cuda")
c = torch.randn(B, 7, H, W, device="cuda")
for i in range(3):
    records.clear()
    net.forward_split(x, c, 73)
    torch.cuda.synchronize()
rows = [(n, desc, e0.elapsed_time(e1)) for n, desc, e0, e1 in records]
total = sum(r[2] for r in rows)
print(f"# B={B} {dtype} {H}x{W}: {len(rows)} calls, sum of per-call event times {total:.2f} ms")
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for n, (desc, fl, by), ms in rows:
    a = agg[(n, desc)]
    a[0] += 1
    a[1] += ms
    a[2] += fl
    a[3] += by
print(f"{'ms':>8s} {'%':>5s} {'n':>3s}  {'TF/s|GB/s':>9s}  call")
for (n, desc), (cnt, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rate = f"{fl / ms / 1e9:7.0f}TF" if fl else (f"{by / ms / 1e6:7.0f}GB" if by else "        -")
    print(f"{ms:8.3f} {100 * ms / total:5.1f} {cnt:3d}  {rate}  {n} {desc}")
by_kind = defaultdict(float)
for n, _, ms in rows:
    by_kind[n] += ms
print("# by op: " + ", ".join(f"{k} {v:.2f} ms ({100 * v / total:.1f}%)" for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1])))

"""Developer tool: per-call table of ONE training step (forward + backward + optimizer) - every library call is bracketed
by CUDA events and reported with its shape, time, and algorithmic TFLOP/s (convs, weight gradients) or GB/s.
usage: python tools/profile_train.py [B] [dtype] [H] [W]"""
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import ops  # noqa: E402
from d3roma_b200.training import UNetTrainer  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
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
    flops = 2.0 * Bn * (Hh // s) * (Ww // s) * Cout * Kt
    tag = f"k{k['ksize']}s{s}" + (f" up{ph}" if ph else "") + (" +sc" if k.get("sc1") is not None else "") + \
        (" +res" if k.get("residual") is not None else "")
    return (f"{Hh}x{Ww} {Cin}->{Cout} K={Kt} {tag}", flops, 0.0)


def wgrad_desc(a, k, out):
    x, dy = a[0], a[1]
    Bn, Hh, Ww, Cin = x.shape
    Cout = dy.shape[3]
    ks, s = k["ksize"], k.get("stride", 1)
    return (f"{Hh}x{Ww} {Cin}->{Cout} k{ks}s{s}", 2.0 * Bn * (Hh // s) * (Ww // s) * Cout * Cin * ks * ks, 0.0)


def gn_desc(a, k, out):
    x = a[0]
    return (f"{tuple(x.shape[1:3])} C={out.shape[-1]}", 0.0, 2.0 * out.numel() * x.element_size())


def gnb_desc(a, k, out):
    x = a[0]
    n_add = len([t for t in k.get("addends", ()) if t is not None])
    # reduce: x + dy; apply: x + dy + addends + dx
    return (f"{tuple(x.shape[1:3])} C={out.shape[-1]} addends={n_add}", 0.0, (5.0 + n_add) * out.numel() * x.element_size())


def attn_desc(a, k, out):
    q = a[0]
    Bn, Hh, Ww, C3 = q.shape
    return (f"N={Hh * Ww}", 4.0 * Bn * (Hh * Ww) ** 2 * (C3 // 3), 0.0)


def attnb_desc(a, k, out):
    q = a[0]
    Bn, Hh, Ww, C3 = q.shape
    return (f"N={Hh * Ww}", 2.5 * 4.0 * Bn * (Hh * Ww) ** 2 * (C3 // 3), 0.0)


def bw_desc(factor):
    def f(a, k, out):
        x = a[0]
        return (f"{tuple(x.shape)}", 0.0, factor * x.numel() * x.element_size())
    return f


def misc_desc(a, k, out):
    return ("", 0.0, 0.0)


wrap("conv2d_igemm", conv_desc)
wrap("conv2d_wgrad", wgrad_desc)
wrap("groupnorm_silu", gn_desc)
wrap("groupnorm_silu_bwd", gnb_desc)
wrap("attention_hd8", attn_desc)
wrap("attention_hd8_bwd", attnb_desc)
wrap("add_nhwc16", bw_desc(3.0))
wrap("colsum_nhwc16", bw_desc(1.0))
wrap("upsample_nearest2x", bw_desc(5.0))
wrap("sumpool2x_nhwc16", bw_desc(1.25))
for nm in ("conv_out_fused", "conv_out_bwd", "im2col_input", "masked_loss_grad", "linear_small", "linear_small_bwd",
           "pack_taps16", "sumsq", "adamw_step", "lincomb", "timestep_sincos"):
    wrap(nm, misc_desc)

tr = UNetTrainer(8, 1, compute_dtype=dtype, loss_scale=65536.0 if dtype == "fp16" else 1.0, device="cuda")
tr.load_state_dict(random_init_state_dict(8, 1, seed=0))
sched = d.MyDDPMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample"))
g = torch.Generator().manual_seed(0)
clean = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).cuda()
cond = (torch.rand(B, 7, H, W, generator=g) * 2 - 1).cuda()
noise = torch.randn(B, 1, H, W, generator=g).cuda()
t = torch.randint(0, 128, (B,), generator=g).cuda()
for i in range(3):
    records.clear()
    tr.train_step(clean, cond, noise, t, sched)
    torch.cuda.synchronize()
rows = [(n, desc, e0.elapsed_time(e1)) for n, desc, e0, e1 in records]
total = sum(r[2] for r in rows)
print(f"# training step B={B} {dtype} {H}x{W}: {len(rows)} calls, sum of per-call event times {total:.2f} ms")
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for n, (desc, fl, by), ms in rows:
    a = agg[(n, desc)]
    a[0] += 1
    a[1] += ms
    a[2] += fl
    a[3] += by
print(f"{'ms':>8s} {'%':>5s} {'n':>3s}  {'TF/s|GB/s':>9s}  call")
for (n, desc), (cnt, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    rate = f"{fl / ms / 1e9:7.0f}TF" if fl else (f"{by / ms / 1e6:7.0f}GB" if by else "        -")
    print(f"{ms:8.3f} {100 * ms / total:5.1f} {cnt:3d}  {rate}  {n} {desc}")
by_kind = defaultdict(lambda: [0.0, 0.0, 0.0])
for n, (desc, fl, by), ms in rows:
    by_kind[n][0] += ms
    by_kind[n][1] += fl
    by_kind[n][2] += by
print("# by op:")
for k, (ms, fl, by) in sorted(by_kind.items(), key=lambda kv: -kv[1][0]):
    rate = f"{fl / ms / 1e9:6.0f} TF/s" if fl else (f"{by / ms / 1e6:6.0f} GB/s" if by else "")
    print(f"#   {k:22s} {ms:8.2f} ms ({100 * ms / total:4.1f} %)  {rate}")

"""Developer tool: per-op table of the latent family at benchmark scale - every library call of one UNet2DConditionModel
forward (SD-2.1 configuration), one VAE encode and one VAE decode is bracketed by CUDA events and aggregated by op.
usage: python tools/profile_ldm.py [B] [H] [W] [dtype]"""
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402
from d3roma_b200.autoencoder_kl import AutoencoderKL  # noqa: E402
from d3roma_b200.unet_condition import UNet2DConditionModel  # noqa: E402
from d3roma_b200.weights import random_init_cond_unet_state_dict, random_init_vae_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 360
W = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dtype = sys.argv[4] if len(sys.argv) > 4 else "fp16"
records = []


def wrap(name):
    orig = getattr(ops, name)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(*a, **k)
        e1.record()
        tag, flops = name, 0.0
        if name == "conv2d_igemm":
            x, w = a[0], a[1]
            Bn, Hh, Ww, Cin = x.shape
            s = k.get("stride", 1)
            ph = k.get("up2x_phase", 0)
            flops = 2.0 * Bn * (Hh // s) * (Ww // s) * w.shape[0] * w.shape[1]
            kind = "lin" if k["ksize"] == 1 else ("up" if ph else f"c3s{s}")
            tag = f"conv {kind} {Cin}->{k.get('cout_valid') or w.shape[0]}"
        elif name in ("attention_hd64", "attention_hd64_tc"):
            q = a[0]
            Bn, Hh, Ww, C = q.shape
            flops = 4.0 * Bn * (Hh * Ww) ** 2 * C
            tag = f"{name} N={Hh * Ww} heads={a[3]}"
        records.append((tag, flops, e0, e1))
        return out
    setattr(ops, name, f)


for nm in ("conv2d_igemm", "gn_stats", "gn_apply", "layernorm", "geglu", "attention_hd64", "attention_hd64_tc", "cross_attention_small",
           "linear_small", "timestep_sincos", "nchw_to_nhwc16", "nhwc16_to_nchw", "resize_nearest", "groupnorm_silu",
           "conv_in_direct", "conv_out_fused", "softmax_rows_", "transpose_pad", "conv1x1_nchw_small"):
    wrap(nm)


def report(title, wall_ms):
    torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for tag, fl, e0, e1 in records:
        a = agg[tag]
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += fl
    tot = sum(v[1] for v in agg.values())
    print(f"== {title}: wall {wall_ms:.2f} ms, event sum {tot:.2f} ms, {len(records)} launches-ish")
    for tag, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        extra = f"  {fl / ms / 1e9:8.1f} TF/s" if fl else ""
        print(f"  {tag:42s} x{n:4d} {ms:8.3f} ms {100 * ms / tot:5.1f} %{extra}")
    fl = sum(v[2] for v in agg.values())
    print(f"  total algorithmic {fl / 1e12:.2f} TFLOP -> {fl / wall_ms / 1e9:.0f} TF/s over the wall time")
    records.clear()


def timed(fn, reps=3):
    fn()
    records.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    recs = records[-(len(records) // reps):]
    records[:] = recs
    return ms


unet = UNet2DConditionModel(in_channels=12, compute_dtype=dtype).to("cuda")
unet.use_cuda_graph = False   # per-op events need the eager launch sequence
unet.load_state_dict(random_init_cond_unet_state_dict(12, device="cuda"))
vae = AutoencoderKL(compute_dtype=dtype).to("cuda")
vae.load_state_dict(random_init_vae_state_dict(device="cuda"))
h, w = H // 8, W // 8
x = torch.randn(B, 12, h, w, device="cuda")
ctx = torch.randn(B, 2, 1024, device="cuda")
img = torch.rand(B, 3, H, W, device="cuda") * 2 - 1
z = torch.randn(B, 4, h, w, device="cuda") * 0.18215
print(f"B={B} image {H}x{W} latent {h}x{w} dtype {dtype}")
report("UNet2DConditionModel forward", timed(lambda: unet(x, 500, ctx)))
report("VAE encode_latent", timed(lambda: vae.encode_latent(img, 0.18215)))
report("VAE decode_mean", timed(lambda: vae.decode_mean(z, 0.18215)))

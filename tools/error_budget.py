"""Developer tool: which bf16 rounding points dominate the UNet output error?  Emulates rounding classes on the
fp32 oracle (GPU, TF32 off) with hooks and prints the final max-abs error of each class alone / combined."""
import sys
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle.unet_ref import Attention, ResnetBlock2D, make_unet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def r(x):
    return x.to(torch.bfloat16).float()


def run(flags, x, t, hw):
    m = make_unet(8, 1, seed=0, sample_size=hw).cuda()
    hs = []
    if "w" in flags:
        for mod in m.modules():
            if isinstance(mod, (nn.Conv2d, nn.Linear)) and mod.weight.shape[0] > 3 and "time" not in str(type(mod)):
                mod.weight.data = r(mod.weight.data)
    for name, mod in m.named_modules():
        if isinstance(mod, nn.Conv2d) and name != "conv_out":
            if "cin" in flags:
                hs.append(mod.register_forward_pre_hook(lambda mo, inp: (r(inp[0]),)))
            if "cout" in flags and "conv_shortcut" not in name and "conv2" not in name:
                hs.append(mod.register_forward_hook(lambda mo, inp, out: r(out)))
        if isinstance(mod, ResnetBlock2D) and "res" in flags:
            hs.append(mod.register_forward_hook(lambda mo, inp, out: r(out)))
        if isinstance(mod, Attention) and "attn" in flags:
            hs.append(mod.register_forward_hook(lambda mo, inp, out: r(out)))
            for lin in (mod.to_q, mod.to_k, mod.to_v, mod.to_out[0]):
                hs.append(lin.register_forward_pre_hook(lambda mo, inp: (r(inp[0]),)))
            for lin in (mod.to_q, mod.to_k, mod.to_v):
                hs.append(lin.register_forward_hook(lambda mo, inp, out: r(out)))
    with torch.no_grad():
        return m(x, torch.tensor(t)).sample


def main():
    hw = (64, 96)
    for seed in (7, 8):
        x = torch.randn(2, 8, *hw, generator=torch.Generator().manual_seed(seed)).cuda()
        for t in (109, 1):
            base = run((), x, t, hw)
            line = [f"seed={seed} t={t} absmax={base.abs().max():.3f}"]
            for flags in (("w",), ("cin",), ("cout",), ("res",), ("attn",), ("w", "cin"), ("w", "cin", "cout", "res", "attn")):
                y = run(flags, x, t, hw)
                line.append(f"{'+'.join(flags)}={(y - base).abs().max().item():.2e}")
            print("  ".join(line), flush=True)
    # our kernel
    from d3roma_b200.unet import UNet2DModel
    ref = make_unet(8, 1, seed=0, sample_size=hw)
    net = UNet2DModel(sample_size=hw, in_channels=8, out_channels=1).to("cuda")
    net.load_state_dict(ref.state_dict())
    for seed in (7, 8):
        x = torch.randn(2, 8, *hw, generator=torch.Generator().manual_seed(seed)).cuda()
        for t in (109, 1):
            base = run((), x, t, hw)
            y = net(x, torch.tensor(t)).sample
            e = (y - base).abs()
            print(f"kernel seed={seed} t={t}: max {e.max().item():.2e} mean {e.mean().item():.2e} "
                  f"rms {e.pow(2).mean().sqrt().item():.2e}")


if __name__ == "__main__":
    main()

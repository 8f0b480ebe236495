"""Random-init ``state_dict`` of the released pixel-space architecture (diffusers key names),
with PyTorch's default layer initialisation (Conv2d / Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias;
GroupNorm: ones / zeros).  There is no network access for checkpoints, so benchmarks and the smoke test run on
random-init weights of the same shapes (BASELINE.json: "random-init weights of the released architecture")."""
from __future__ import annotations

import math
from typing import Dict

import torch

from .unet import UNet2DModel, _resnet_specs


def _uniform(shape, fan_in, g):
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def random_init_state_dict(in_channels: int = 8, out_channels: int = 1, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = _uniform((cout, cin, k, k), cin * k * k, g)
        sd[name + ".bias"] = _uniform((cout,), cin * k * k, g)

    def linear(name, cout, cin):
        sd[name + ".weight"] = _uniform((cout, cin), cin, g)
        sd[name + ".bias"] = _uniform((cout,), cin, g)

    def norm(name, c):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)

    conv("conv_in", 128, in_channels, 3)
    linear("time_embedding.linear_1", 512, 128)
    linear("time_embedding.linear_2", 512, 512)
    for kind, i, j, cin_h, cin_s, cout, attn in _resnet_specs():
        p = UNet2DModel._rprefix(kind, i, j)
        cin = cin_h + cin_s
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        linear(p + ".time_emb_proj", cout, 512)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)
        if attn:
            a = UNet2DModel._aprefix(kind, i, j)
            norm(a + ".group_norm", cout)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                linear(a + "." + n, cout, cout)
    chans = (128, 128, 256, 256, 512)
    for i in range(5):
        conv(f"down_blocks.{i}.downsamplers.0.conv", chans[i], chans[i], 3)
    up_chans = (512, 512, 256, 256, 128)
    for i in range(5):
        conv(f"up_blocks.{i}.upsamplers.0.conv", up_chans[i], up_chans[i], 3)
    norm("conv_norm_out", 128)
    conv("conv_out", out_channels, 128, 3)
    return sd


# ---------------------------------------------------------------------------------------------------------------------
# latent family: random-init state_dicts of the SD-2.1 UNet2DConditionModel / SD AutoencoderKL layouts (diffusers key
# names), drawn on ``device`` (866 M parameters take seconds on the GPU, a minute on the host).
# ---------------------------------------------------------------------------------------------------------------------
def _gen(device, seed):
    g = torch.Generator(device=device).manual_seed(seed)

    def uniform(shape, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound

    return uniform


def random_init_cond_unet_state_dict(in_channels: int = 16, seed: int = 0, device="cpu", **cfg) -> Dict[str, torch.Tensor]:
    """Keys / shapes of ``UNet2DConditionModel(in_channels, **cfg)`` (defaults: the SD-2.1 configuration)."""
    from .unet_condition import UNet2DConditionModel
    net = UNet2DConditionModel(in_channels=in_channels, **cfg)
    c = net.config
    boc, xd = c.block_out_channels, c.cross_attention_dim
    u = _gen(device, seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"], sd[name + ".bias"] = u((cout, cin, k, k), cin * k * k), u((cout,), cin * k * k)

    def linear(name, cout, cin, bias=True):
        sd[name + ".weight"] = u((cout, cin), cin)
        if bias:
            sd[name + ".bias"] = u((cout,), cin)

    def norm(name, ch):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(ch, device=device), torch.zeros(ch, device=device)

    temb = boc[0] * 4
    conv("conv_in", boc[0], in_channels, 3)
    linear("time_embedding.linear_1", temb, boc[0])
    linear("time_embedding.linear_2", temb, temb)
    for p, cin_h, cin_s, cout in net._resnets():
        cin = cin_h + cin_s
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        linear(p + ".time_emb_proj", cout, temb)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)
    for p, ch, heads in net._transformers():
        norm(p + ".norm", ch)
        linear(p + ".proj_in", ch, ch)
        linear(p + ".proj_out", ch, ch)
        t = p + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{t}.{n}", ch)
        for n in ("to_q", "to_k", "to_v"):
            linear(f"{t}.attn1.{n}", ch, ch, bias=False)
        linear(t + ".attn1.to_out.0", ch, ch)
        linear(t + ".attn2.to_q", ch, ch, bias=False)
        linear(t + ".attn2.to_k", ch, xd, bias=False)
        linear(t + ".attn2.to_v", ch, xd, bias=False)
        linear(t + ".attn2.to_out.0", ch, ch)
        linear(t + ".ff.net.0.proj", 8 * ch, ch)
        linear(t + ".ff.net.2", ch, 4 * ch)
    for i in range(len(boc) - 1):
        conv(f"down_blocks.{i}.downsamplers.0.conv", boc[i], boc[i], 3)
        conv(f"up_blocks.{i}.upsamplers.0.conv", boc[::-1][i], boc[::-1][i], 3)
    norm("conv_norm_out", boc[0])
    conv("conv_out", c.out_channels, boc[0], 3)
    return sd


def random_init_vae_state_dict(seed: int = 0, device="cpu", **cfg) -> Dict[str, torch.Tensor]:
    """Keys / shapes of ``AutoencoderKL(**cfg)`` (defaults: the SD VAE)."""
    from .autoencoder_kl import AutoencoderKL
    vae = AutoencoderKL(**cfg)
    c = vae.config
    boc = c.block_out_channels
    u = _gen(device, seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"], sd[name + ".bias"] = u((cout, cin, k, k), cin * k * k), u((cout,), cin * k * k)

    def linear(name, cout, cin):
        sd[name + ".weight"], sd[name + ".bias"] = u((cout, cin), cin), u((cout,), cin)

    def norm(name, ch):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(ch, device=device), torch.zeros(ch, device=device)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def mid(p, ch):
        resnet(p + ".resnets.0", ch, ch)
        resnet(p + ".resnets.1", ch, ch)
        a = p + ".attentions.0"
        norm(a + ".group_norm", ch)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            linear(f"{a}.{n}", ch, ch)

    conv("encoder.conv_in", boc[0], c.in_channels, 3)
    for p, cin, cout in vae._enc_resnets():
        resnet(p, cin, cout)
    for i in range(len(boc) - 1):
        conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", boc[i], boc[i], 3)
    mid("encoder.mid_block", boc[-1])
    norm("encoder.conv_norm_out", boc[-1])
    conv("encoder.conv_out", 2 * c.latent_channels, boc[-1], 3)
    conv("quant_conv", 2 * c.latent_channels, 2 * c.latent_channels, 1)
    conv("post_quant_conv", c.latent_channels, c.latent_channels, 1)
    conv("decoder.conv_in", boc[-1], c.latent_channels, 3)
    mid("decoder.mid_block", boc[-1])
    for p, cin, cout in vae._dec_resnets():
        resnet(p, cin, cout)
    for i in range(len(boc) - 1):
        conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", boc[::-1][i], boc[::-1][i], 3)
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", c.out_channels, boc[0], 3)
    return sd

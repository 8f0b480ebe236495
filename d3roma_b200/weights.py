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

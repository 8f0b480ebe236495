"""ORACLE (test infrastructure, not product code): CPU restatement of ``diffusers.UNet2DModel``.

The reference (jie0530/d3roma) builds its pixel-space denoiser at ``train.py:788-810`` as
``UNet2DModel(sample_size=[H,W], in_channels=Cin, out_channels=dc, layers_per_block=2,
block_out_channels=(128,128,256,256,512,512), down_block_types=("DownBlock2D",)*4 +
("AttnDownBlock2D","DownBlock2D"), up_block_types=("UpBlock2D","AttnUpBlock2D")+("UpBlock2D",)*4)``
and calls it at ``core/custom_pipelines.py:173``.  ``diffusers`` itself is a third-party dependency
that is NOT vendored in the reference tree and NOT installed in this image (README.md:51-54 leaves
it unpinned), so the algorithm is restated here from its published architecture (SURVEY.md
Appendix A) in plain ``torch.nn`` with the diffusers ``state_dict`` key names.

Parity status: **parity unpinned** for the UNet arithmetic (no golden vectors exist anywhere in the
reference for it); the restatement is gated by the published parameter counts
(113 673 219 for in=3/out=3, 113 676 673 for in=8/out=1) in ``tests/test_oracle_unet.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline leg of ``bench.py`` import this.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

BLOCK_OUT = (128, 128, 256, 256, 512, 512)
DOWN_TYPES = ("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D")
UP_TYPES = ("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4


def timestep_embedding(timesteps: torch.Tensor, dim: int = 128) -> torch.Tensor:
    """diffusers ``Timesteps(num_channels=128, flip_sin_to_cos=True, downscale_freq_shift=0)``
    (SURVEY Appendix A.2): ``cat[cos(t*f_i), sin(t*f_i)]`` with ``f_i = exp(-ln(1e4) * i / 64)``."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    arg = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    """Appendix A.3: GN-SiLU-conv3x3, + Linear(SiLU(temb)), GN-SiLU-conv3x3, 1x1 shortcut iff Cin != Cout."""

    def __init__(self, cin: int, cout: int, temb_dim: int = 512, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """Appendix A.5: legacy AttentionBlock semantics, heads = C / 8, residual connection, scale 1/sqrt(8)."""

    def __init__(self, channels: int, head_dim: int = 8, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.heads = channels // head_dim
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        y = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)  # [B, HW, C]
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)

        def split(t):
            return t.view(b, h * w, self.heads, c // self.heads).transpose(1, 2)

        q, k, v = split(q), split(k), split(v)
        scale = 1.0 / math.sqrt(c // self.heads)
        p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
        o = torch.matmul(p, v).transpose(1, 2).reshape(b, h * w, c)
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(b, c, h, w) + res


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin: int, cout: int, attn: bool, add_down: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout), ResnetBlock2D(cout, cout)])
        self.attentions = nn.ModuleList([Attention(cout), Attention(cout)]) if attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Attention(c)])

    def forward(self, x, temb):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin: int, prev: int, cout: int, attn: bool, add_up: bool):
        super().__init__()
        rs = []
        for j in range(3):
            skip = cin if j == 2 else cout
            rin = prev if j == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([Attention(cout) for _ in range(3)]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips: List[torch.Tensor], temb):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)  # hidden first, skip second (Appendix A.4)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DModelRef(nn.Module):
    """Plain-torch restatement of the reference's denoiser; ``forward(sample, t).sample`` like diffusers."""

    def __init__(self, in_channels: int = 8, out_channels: int = 1, sample_size: Sequence[int] = (384, 640),
                 block_out_channels: Tuple[int, ...] = BLOCK_OUT):
        super().__init__()
        self.sample_size = list(sample_size)
        self.in_channels, self.out_channels = in_channels, out_channels
        boc = block_out_channels
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], boc[0] * 4)
        self.down_blocks = nn.ModuleList()
        out_c = boc[0]
        for i, t in enumerate(DOWN_TYPES):
            in_c, out_c = out_c, boc[i]
            self.down_blocks.append(DownBlock(in_c, out_c, attn=t.startswith("Attn"), add_down=i < len(boc) - 1))
        self.mid_block = MidBlock(boc[-1])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_c = rev[0]
        for i, t in enumerate(UP_TYPES):
            prev, out_c = out_c, rev[i]
            in_c = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(in_c, prev, out_c, attn=t.startswith("Attn"), add_up=i < len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, sample: torch.Tensor, timestep, return_dict: bool = True):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        emb = self.time_embedding(timestep_embedding(t, self.conv_in.out_channels).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb)
            skips += outs
        x = self.mid_block(x, emb)
        for blk in self.up_blocks:
            x = blk(x, skips, emb)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        if not return_dict:
            return (x,)
        return SimpleNamespace(sample=x)


def make_unet(in_channels: int = 8, out_channels: int = 1, seed: int = 0, sample_size=(384, 640)) -> UNet2DModelRef:
    """Random-init weights of the released architecture (default torch init under a fixed seed)."""
    torch.manual_seed(seed)
    return UNet2DModelRef(in_channels, out_channels, sample_size).eval()

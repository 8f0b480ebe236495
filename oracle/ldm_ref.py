"""ORACLE (test infrastructure only - never imported by the product path): plain ``torch.nn`` restatement of the two
third-party networks of the reference's LATENT family,

  * ``diffusers.UNet2DConditionModel`` in the Stable-Diffusion-2.1 configuration the reference loads with
    ``UNet2DConditionModel.from_pretrained("checkpoint/stable-diffusion/unet")`` and whose ``conv_in`` it widens to
    ``4 * (k + 1)`` input channels (/root/reference/train.py:743-772), called as ``unet(x, t, empty_text_embed).sample``
    (core/custom_pipelines.py:386);
  * ``diffusers.AutoencoderKL`` (``AutoencoderKL.from_pretrained("checkpoint/stable-diffusion/vae")``,
    train.py:830-833), used through ``vae.encoder / quant_conv / post_quant_conv / decoder``
    (core/custom_pipelines.py:35-59, 255-272).

``diffusers`` is an un-vendored, unpinned dependency (reference README.md:51-54) and is not installed in this image, and
no checkpoint is reachable (no network): PARITY UNPINNED for these two architectures.  The restatement follows the
published diffusers module structure (ResnetBlock2D, Transformer2DModel with ``use_linear_projection``,
BasicTransformerBlock = LayerNorm/self-attn/LayerNorm/cross-attn/LayerNorm/GEGLU-FF, Downsample2D / Upsample2D,
UNetMidBlock2D(CrossAttn), DownEncoderBlock2D / UpDecoderBlock2D) with diffusers' parameter names, and is gated by the
published parameter counts: 865 910 724 (SD-2.1 UNet, in_channels 4) and 83 653 863 (SD VAE) - see
tests/test_oracle_ldm.py.  For the VAE there is an independent second source as well: transformers' ports of the original
latent-diffusion autoencoder (ChameleonVQVAEEncoder, JanusVQVAE* blocks) agree with the encoder as a whole and with the
decoder's blocks when given the same weights (tests/test_oracle_ldm_crosscheck.py).  Reduced-width instances of the same
modules are what the GPU parity tests compare against.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------------
# shared blocks
# ---------------------------------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_channels: Optional[int], eps: float, groups: int = 32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Downsample2D(nn.Module):
    """padding=1: conv3x3 stride 2 pad 1 (UNet); padding=0: F.pad(0,1,0,1) then conv3x3 stride 2 (VAE encoder)."""

    def __init__(self, c: int, padding: int):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class Attention(nn.Module):
    """diffusers ``Attention``: ``heads`` heads of ``dim_head``; q/k/v without bias in the transformer blocks, with
    bias (and a GroupNorm + residual) in the VAE mid block."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_dim: Optional[int] = None, bias: bool = False,
                 norm_groups: Optional[int] = None, eps: float = 1e-5, residual: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale, self.residual = heads, dim_head ** -0.5, residual
        self.group_norm = nn.GroupNorm(norm_groups, query_dim, eps=eps) if norm_groups else None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        res, shape = x, None
        if x.dim() == 4:
            shape = x.shape
            if self.group_norm is not None:
                x = self.group_norm(x)
            x = x.flatten(2).transpose(1, 2)
        ctx = x if context is None else context
        B, N, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        sp = lambda t: t.view(B, -1, self.heads, t.shape[-1] // self.heads).transpose(1, 2)  # noqa: E731
        q, k, v = sp(q), sp(k), sp(v)
        p = torch.softmax(q @ k.transpose(-1, -2) * self.scale, dim=-1)
        o = (p @ v).transpose(1, 2).reshape(B, N, -1)
        o = self.to_out[0](o)
        if shape is not None:
            o = o.transpose(1, 2).reshape(shape)
        return o + res if self.residual else o


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, dim_head, cross_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    """``use_linear_projection=True`` form (SD 2.x): GroupNorm(eps 1e-6) -> tokens -> Linear -> block -> Linear -> +x."""

    def __init__(self, channels: int, heads: int, cross_dim: int, groups: int = 32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, channels // heads, cross_dim)])
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, x, context):
        B, C, H, W = x.shape
        h = self.norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, context)
        h = self.proj_out(h)
        return h.reshape(B, H, W, C).permute(0, 3, 1, 2) + x


# ---------------------------------------------------------------------------------------------------------------------
# UNet2DConditionModel
# ---------------------------------------------------------------------------------------------------------------------
class _Block(nn.Module):
    pass


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)``: [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels: int = 4, out_channels: int = 4,
                 block_out_channels: Sequence[int] = (320, 640, 1280, 1280), layers_per_block: int = 2,
                 attention_head_dim: Sequence[int] = (5, 10, 20, 20), cross_attention_dim: int = 1024,
                 norm_num_groups: int = 32, sample_size=64):
        super().__init__()
        boc = tuple(block_out_channels)
        heads = tuple(attention_head_dim)   # diffusers' "attention_head_dim" is the NUMBER of heads for these blocks
        n = len(boc)
        temb = boc[0] * 4
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, attention_head_dim=heads,
                                      cross_attention_dim=cross_attention_dim, norm_num_groups=norm_num_groups,
                                      sample_size=sample_size)
        self.sample_size = sample_size
        g = norm_num_groups
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = nn.Module()
        self.time_embedding.linear_1 = nn.Linear(boc[0], temb)
        self.time_embedding.linear_2 = nn.Linear(temb, temb)
        self.down_blocks = nn.ModuleList()
        out_c = boc[0]
        for i in range(n):
            in_c, out_c = out_c, boc[i]
            blk = _Block()
            blk.resnets = nn.ModuleList([ResnetBlock2D(in_c if j == 0 else out_c, out_c, temb, 1e-5, g)
                                         for j in range(layers_per_block)])
            if i < n - 1:   # CrossAttnDownBlock2D x (n-1), DownBlock2D last
                blk.attentions = nn.ModuleList([Transformer2DModel(out_c, heads[i], cross_attention_dim, g)
                                                for _ in range(layers_per_block)])
                blk.downsamplers = nn.ModuleList([Downsample2D(out_c, padding=1)])
            self.down_blocks.append(blk)
        self.mid_block = _Block()
        self.mid_block.resnets = nn.ModuleList([ResnetBlock2D(boc[-1], boc[-1], temb, 1e-5, g) for _ in range(2)])
        self.mid_block.attentions = nn.ModuleList([Transformer2DModel(boc[-1], heads[-1], cross_attention_dim, g)])
        self.up_blocks = nn.ModuleList()
        rev, rheads = boc[::-1], heads[::-1]
        out_c = rev[0]
        for i in range(n):
            prev, out_c = out_c, rev[i]
            in_c = rev[min(i + 1, n - 1)]
            blk = _Block()
            res = []
            for j in range(layers_per_block + 1):
                skip = in_c if j == layers_per_block else out_c
                res.append(ResnetBlock2D((prev if j == 0 else out_c) + skip, out_c, temb, 1e-5, g))
            blk.resnets = nn.ModuleList(res)
            if i > 0:       # UpBlock2D first, CrossAttnUpBlock2D x (n-1)
                blk.attentions = nn.ModuleList([Transformer2DModel(out_c, rheads[i], cross_attention_dim, g)
                                                for _ in range(layers_per_block + 1)])
            if i < n - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(out_c)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True):
        t = torch.as_tensor(timestep, device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        temb = timestep_embedding(t, self.config.block_out_channels[0]).to(sample.dtype)   # diffusers: t_emb.to(sample.dtype)
        temb = self.time_embedding.linear_2(F.silu(self.time_embedding.linear_1(temb)))
        ctx = encoder_hidden_states
        n_up = len(self.up_blocks) - 1
        odd = any(s % (2 ** n_up) != 0 for s in sample.shape[-2:])   # diffusers: forward_upsample_size
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            for j, r in enumerate(blk.resnets):
                h = r(h, temb)
                if hasattr(blk, "attentions"):
                    h = blk.attentions[j](h, ctx)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
                skips.append(h)
        h = self.mid_block.resnets[0](h, temb)
        h = self.mid_block.attentions[0](h, ctx)
        h = self.mid_block.resnets[1](h, temb)
        for i, blk in enumerate(self.up_blocks):
            take = skips[-len(blk.resnets):]
            skips = skips[:-len(blk.resnets)]
            up_size = skips[-1].shape[2:] if (odd and i < n_up) else None
            for j, r in enumerate(blk.resnets):
                h = r(torch.cat([h, take.pop()], dim=1), temb)
                if hasattr(blk, "attentions"):
                    h = blk.attentions[j](h, ctx)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h, up_size)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h) if return_dict else (h,)


def widen_conv_in(model: UNet2DConditionModel, dup: int) -> UNet2DConditionModel:
    """The reference's first-layer adaptation (train.py:752-768): the 4-channel conv_in weight repeated ``dup`` times
    along the input channels and divided by ``dup`` (bias divided too)."""
    old = model.conv_in
    new = nn.Conv2d(old.in_channels * dup, old.out_channels, 3, padding=1)
    with torch.no_grad():
        new.weight.copy_(old.weight.repeat(1, dup, 1, 1) / dup)
        new.bias.copy_(old.bias / dup)
    model.conv_in = new
    model.config.in_channels = old.in_channels * dup
    return model


def make_cond_unet(in_channels: int = 16, seed: int = 0, **kw) -> UNet2DConditionModel:
    """Random-init instance (torch default init under ``seed``); ``block_out_channels`` etc. may be reduced for tests."""
    torch.manual_seed(seed)
    net = UNet2DConditionModel(in_channels=in_channels, **kw).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    return net


# ---------------------------------------------------------------------------------------------------------------------
# AutoencoderKL
# ---------------------------------------------------------------------------------------------------------------------
class _MidBlock(nn.Module):
    def __init__(self, c: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, 1e-6, groups) for _ in range(2)])
        self.attentions = nn.ModuleList([Attention(c, 1, c, bias=True, norm_groups=groups, eps=1e-6, residual=True)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, in_channels: int, latent: int, boc: Sequence[int], layers: int, groups: int):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_c = boc[0]
        for i, c in enumerate(boc):
            in_c, out_c = out_c, c
            blk = _Block()
            blk.resnets = nn.ModuleList([ResnetBlock2D(in_c if j == 0 else out_c, out_c, None, 1e-6, groups)
                                         for j in range(layers)])
            if i < len(boc) - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(out_c, padding=0)])
            self.down_blocks.append(blk)
        self.mid_block = _MidBlock(boc[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            for r in blk.resnets:
                h = r(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class Decoder(nn.Module):
    def __init__(self, latent: int, out_channels: int, boc: Sequence[int], layers: int, groups: int):
        super().__init__()
        rev = tuple(boc[::-1])
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = _MidBlock(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        out_c = rev[0]
        for i, c in enumerate(rev):
            in_c, out_c = out_c, c
            blk = _Block()
            blk.resnets = nn.ModuleList([ResnetBlock2D(in_c if j == 0 else out_c, out_c, None, 1e-6, groups)
                                         for j in range(layers + 1)])
            if i < len(rev) - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(out_c)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(groups, rev[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            for r in blk.resnets:
                h = r(h)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 4,
                 block_out_channels: Sequence[int] = (128, 256, 512, 512), layers_per_block: int = 2,
                 norm_num_groups: int = 32, scaling_factor: float = 0.18215):
        super().__init__()
        boc = tuple(block_out_channels)
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      latent_channels=latent_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, norm_num_groups=norm_num_groups,
                                      scaling_factor=scaling_factor)
        self.encoder = Encoder(in_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, boc, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)


def make_vae(seed: int = 0, **kw) -> AutoencoderKL:
    torch.manual_seed(seed)
    net = AutoencoderKL(**kw).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    return net

"""``AutoencoderKL`` — the VAE of the reference's LATENT family on the sm_100a kernels.

The reference loads ``diffusers.AutoencoderKL.from_pretrained("checkpoint/stable-diffusion/vae")``
(``/root/reference/train.py:830-833``) and touches exactly four attributes of it: ``vae.encoder(x)``,
``vae.quant_conv(h)`` (``core/custom_pipelines.py:35-59``: the mean half of the moments times 0.18215 is the latent) and
``vae.post_quant_conv(z)``, ``vae.decoder(z)`` (``:255-272``: the decoded 3 channels are averaged into the map).  This
class provides those four callables with the same tensor conventions (fp32 NCHW in and out) plus ``load_state_dict`` with
diffusers key names, ``.config.scaling_factor``, ``.to``, ``.device``:

    conv_in (3 -> 128)            d3_conv_in_direct (fp32 planes -> 16-bit NHWC + GroupNorm statistics)
    ResnetBlock2D (no temb)       d3_groupnorm_silu -> d3_conv2d_igemm -> d3_groupnorm_silu -> d3_conv2d_igemm
                                  (1x1 shortcut folded as K-slab / identity residual in the epilogue); eps 1e-6
    Downsample2D(padding=0)       d3_conv2d_igemm stride 2 with asym_pad (F.pad(0,1,0,1) + unpadded conv)
    Upsample2D                    4 sub-pixel phase convs
    mid-block attention           one 512-wide head: q / k / v / out projections and BOTH attention GEMMs (S = Q K^T,
                                  O = P V) on the tcgen05 igemm kernel, d3_softmax_rows / d3_transpose_pad16 in between
    encoder conv_out (512 -> 8)   igemm with padded output rows; decoder conv_in (4 -> 512): igemm over a 64-channel
                                  zero-padded NHWC copy of the latent
    decoder conv_out (128 -> 3)   d3_conv_out_fused (GroupNorm + SiLU + conv, fp32 NCHW out)
    quant / post_quant 1x1 convs  d3_conv1x1_nchw_small (fp32)

Forward only (the reference freezes the VAE, train.py:831-833).  No PyTorch fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import torch

from . import ops
from ._config import PretrainedMixin
from .unet import _Act, _StatsArena


def _p128(c: int) -> int:
    return (c + 127) // 128 * 128


class _Conv1x1:
    """``vae.quant_conv`` / ``vae.post_quant_conv``: a 1x1 conv over an fp32 NCHW tensor with <= 8 channels."""

    def __init__(self, owner, key):
        self._o, self._k = owner, key

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        W = self._o._weights()
        return ops.conv1x1_nchw_small(x, W[self._k + ".w"], W[self._k + ".b"])


class _Encoder:
    def __init__(self, owner):
        self._o = owner

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self._o._encode(x)


class _Decoder:
    def __init__(self, owner):
        self._o = owner

    def __call__(self, z: torch.Tensor) -> torch.Tensor:
        return self._o._decode(z)


class AutoencoderKL(PretrainedMixin):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 4,
                 block_out_channels: Sequence[int] = (128, 256, 512, 512), layers_per_block: int = 2,
                 norm_num_groups: int = 32, scaling_factor: float = 0.18215, compute_dtype: str = "fp16", **unused):
        boc = tuple(block_out_channels)
        if boc[0] != 128 or any(c % 128 or c > 1024 for c in boc) or norm_num_groups != 32:
            raise NotImplementedError("block_out_channels must start at 128, be multiples of 128 (<= 1024), 32 groups")
        if in_channels > 12 or out_channels not in (1, 3) or latent_channels > 4:
            raise NotImplementedError("in_channels <= 12, out_channels 1 or 3, latent_channels <= 4")
        if compute_dtype not in ("fp16", "bf16"):
            raise ValueError("compute_dtype must be 'fp16' or 'bf16'")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      latent_channels=latent_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, norm_num_groups=norm_num_groups,
                                      scaling_factor=scaling_factor)
        self.compute_dtype = compute_dtype
        self._adt = torch.float16 if compute_dtype == "fp16" else torch.bfloat16
        self.device = torch.device("cpu")
        self.dtype = torch.float32
        self._w: Dict[str, torch.Tensor] = {}
        self._loaded = False
        self._arena: Optional[_StatsArena] = None
        self._fold: Dict[tuple, tuple] = {}
        self.encoder, self.decoder = _Encoder(self), _Decoder(self)
        self.quant_conv, self.post_quant_conv = _Conv1x1(self, "quant_conv"), _Conv1x1(self, "post_quant_conv")

    def _weights(self):
        if not self._loaded:
            raise RuntimeError("AutoencoderKL: load_state_dict() has not been called")
        return self._w

    # ------------------------------------------------------------------ structure
    def _enc_resnets(self):
        boc, L = self.config.block_out_channels, self.config.layers_per_block
        out, out_c = [], boc[0]
        for i, c in enumerate(boc):
            in_c, out_c = out_c, c
            out += [(f"encoder.down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c) for j in range(L)]
        return out

    def _dec_resnets(self):
        boc, L = self.config.block_out_channels, self.config.layers_per_block
        rev = boc[::-1]
        out, out_c = [], rev[0]
        for i, c in enumerate(rev):
            in_c, out_c = out_c, c
            out += [(f"decoder.up_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c) for j in range(L + 1)]
        return out

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("d3roma_b200.AutoencoderKL needs a CUDA device (no CPU fallback)")
        dev = self.device if self.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        used, adt = set(), self._adt
        cfg = self.config
        boc = cfg.block_out_channels

        def get(k):
            used.add(k)
            return sd[k].detach().to(dev, torch.float32)

        W = self._w = {}

        def resnet(p, cin, cout):
            for n_ in ("norm1", "norm2"):
                W[f"{p}.{n_}.g"], W[f"{p}.{n_}.b"] = get(f"{p}.{n_}.weight").contiguous(), get(f"{p}.{n_}.bias").contiguous()
            W[p + ".w1"], W[p + ".b1"] = ops.pack_conv_weight(get(p + ".conv1.weight"), dtype=adt), get(p + ".conv1.bias").contiguous()
            b2 = get(p + ".conv2.bias")
            if cin != cout:
                W[p + ".w2"] = ops.pack_conv_weight(get(p + ".conv2.weight"), get(p + ".conv_shortcut.weight"), dtype=adt)
                b2 = b2 + get(p + ".conv_shortcut.bias")
            else:
                W[p + ".w2"] = ops.pack_conv_weight(get(p + ".conv2.weight"), dtype=adt)
            W[p + ".b2"] = b2.contiguous()

        def mid(p, c):
            resnet(p + ".resnets.0", c, c)
            resnet(p + ".resnets.1", c, c)
            a = p + ".attentions.0"
            W[a + ".gn.g"], W[a + ".gn.b"] = get(a + ".group_norm.weight").contiguous(), get(a + ".group_norm.bias").contiguous()
            for n_, k in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("o", "to_out.0")):
                W[f"{a}.{n_}.w"] = get(f"{a}.{k}.weight").to(adt).contiguous()
                W[f"{a}.{n_}.b"] = get(f"{a}.{k}.bias").contiguous()

        # encoder
        W["enc.in.wT"] = ops.pack_conv_in_direct(get("encoder.conv_in.weight"), 0, cfg.in_channels)
        W["enc.in.b"] = get("encoder.conv_in.bias").contiguous()
        for p, cin, cout in self._enc_resnets():
            resnet(p, cin, cout)
        for i in range(len(boc) - 1):
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            W[p + ".w"], W[p + ".b"] = ops.pack_conv_weight(get(p + ".weight"), dtype=adt), get(p + ".bias").contiguous()
        mid("encoder.mid_block", boc[-1])
        W["enc.out.g"], W["enc.out.b"] = get("encoder.conv_norm_out.weight").contiguous(), get("encoder.conv_norm_out.bias").contiguous()
        W["enc.out.w"] = ops.pad_rows(ops.pack_conv_weight(get("encoder.conv_out.weight"), dtype=adt))
        W["enc.out.bias"] = ops.pad_rows(get("encoder.conv_out.bias"))
        for k in ("quant_conv", "post_quant_conv"):
            w = get(k + ".weight")
            W[k + ".w"], W[k + ".b"] = w.reshape(w.shape[0], w.shape[1]).contiguous(), get(k + ".bias").contiguous()
        # decoder
        w_in = torch.nn.functional.pad(get("decoder.conv_in.weight"), (0, 0, 0, 0, 0, 64 - cfg.latent_channels))
        W["dec.in.w"], W["dec.in.b"] = ops.pack_conv_weight(w_in, dtype=adt), get("decoder.conv_in.bias").contiguous()
        mid("decoder.mid_block", boc[-1])
        for p, cin, cout in self._dec_resnets():
            resnet(p, cin, cout)
        for i in range(len(boc) - 1):
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            for ph, wp in enumerate(ops.pack_upsample_weights(get(p + ".weight"), dtype=adt)):
                W[p + f".w{ph}"] = wp
            W[p + ".b"] = get(p + ".bias").contiguous()
        W["dec.out.g"], W["dec.out.b"] = get("decoder.conv_norm_out.weight").contiguous(), get("decoder.conv_norm_out.bias").contiguous()
        w_out = get("decoder.conv_out.weight")
        W["dec.out.w"] = w_out.permute(0, 2, 3, 1).reshape(cfg.out_channels, 9, 128).contiguous()
        W["dec.out.bias"] = get("decoder.conv_out.bias").contiguous()
        if strict:
            extra = set(sd) - used
            if extra:
                raise KeyError(f"unexpected keys in state_dict: {sorted(extra)[:5]} ...")
        self._loaded = True
        self._fold = {}
        self._remember_state_dict(sd)
        return self

    def to(self, device=None, dtype=None):
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise RuntimeError("d3roma_b200.AutoencoderKL runs on CUDA only")
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            if self._loaded and device != self.device:
                self._w = {k: v.to(device) for k, v in self._w.items()}
                self._arena, self._fold = None, {}
            self.device = device
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(self._w.values())

    # ------------------------------------------------------------------ forward pieces
    def _begin(self, B: int, dev):
        need = B * 2 * 24000
        if self._arena is None or self._arena.buf.numel() < need or self._arena.buf.device != dev:
            self._arena = _StatsArena(dev, need)
        self._arena.reset()

    def _conv(self, x, w, b, ksize, B, **kw) -> _Act:
        st = self._arena.take(B, w.shape[0])
        return _Act(ops.conv2d_igemm(x, w, b, ksize=ksize, stats=st, **kw), st)

    def _resnet(self, p, h: _Act, cout: int, B: int) -> _Act:
        W = self._w
        a1 = ops.groupnorm_silu(h.t, h.stats, W[p + ".norm1.g"], W[p + ".norm1.b"], eps=1e-6)
        h1 = self._conv(a1, W[p + ".w1"], W[p + ".b1"], 3, B)
        del a1
        a2 = ops.groupnorm_silu(h1.t, h1.stats, W[p + ".norm2.g"], W[p + ".norm2.b"], eps=1e-6)
        if h.t.shape[3] != cout:
            return self._conv(a2, W[p + ".w2"], W[p + ".b2"], 3, B, sc1=h.t)
        return self._conv(a2, W[p + ".w2"], W[p + ".b2"], 3, B, residual=h.t)

    def _attention(self, a: str, h: _Act, B: int) -> _Act:
        """diffusers ``Attention`` with one head of C channels and a residual connection; tokens = pixels."""
        W = self._w
        _, Hh, Ww, C = h.t.shape
        N = Hh * Ww
        Np = _p128(N)
        n = ops.groupnorm_silu(h.t, h.stats, W[a + ".gn.g"], W[a + ".gn.b"], silu=False, eps=1e-6)
        flat = n.view(1, 1, B * N, C)
        q = ops.conv2d_igemm(flat, W[a + ".q.w"], W[a + ".q.b"], ksize=1).view(B, N, C)
        v = ops.conv2d_igemm(flat, W[a + ".v.w"], W[a + ".v.b"], ksize=1).view(B, N, C)
        if Np == N:
            kbuf = ops.conv2d_igemm(flat, W[a + ".k.w"], W[a + ".k.b"], ksize=1).view(B, N, C)
        else:   # keys as the weight operand of S = Q K^T: rows padded with zeros to the 128-row MMA tile
            kbuf = torch.zeros((B, Np, C), dtype=n.dtype, device=n.device)
            for b in range(B):
                ops.conv2d_igemm(n[b].view(1, 1, N, C), W[a + ".k.w"], W[a + ".k.b"], ksize=1,
                                 out=kbuf[b, :N].view(1, 1, N, C))
        o = torch.empty((B, N, C), dtype=n.dtype, device=n.device)
        zb_s = torch.zeros(Np, dtype=torch.float32, device=n.device)
        zb_o = torch.zeros(C, dtype=torch.float32, device=n.device)
        scale = float(C) ** -0.5
        for b in range(B):
            s = ops.conv2d_igemm(q[b].view(1, 1, N, C), kbuf[b], zb_s, ksize=1)          # [1, 1, N, Np]
            ops.softmax_rows_(s.view(N, Np), N, scale)
            vt = ops.transpose_pad(v[b], Np)                                               # [C, Np]
            ops.conv2d_igemm(s, vt, zb_o, ksize=1, out=o[b].view(1, 1, N, C))
        st = self._arena.take(B, C)
        out = ops.conv2d_igemm(o.view(1, 1, B * N, C), W[a + ".o.w"], W[a + ".o.b"], ksize=1,
                               residual=h.t.view(1, 1, B * N, C))
        # statistics of the block output for the next GroupNorm: the flat launch sees one "sample", so take them with
        # the per-sample shape instead
        out = out.view(B, Hh, Ww, C)
        return self._with_stats(out, st)

    def _with_stats(self, t: torch.Tensor, st: torch.Tensor) -> _Act:
        """Quad statistics of an existing tensor (the producer could not emit them per sample): d3_gn_stats gives
        per-channel moments, folded to quads."""
        B, _, _, C = t.shape
        chan = torch.zeros(B * C * 2, dtype=torch.float64, device=t.device)
        ops.gn_stats(t, chan)
        st.copy_(chan.view(B, C // 4, 4, 2).sum(2))
        return _Act(t, st)

    def _mid(self, p: str, h: _Act, B: int) -> _Act:
        C = h.t.shape[3]
        h = self._resnet(p + ".resnets.0", h, C, B)
        h = self._attention(p + ".attentions.0", h, B)
        return self._resnet(p + ".resnets.1", h, C, B)

    # ------------------------------------------------------------------ encoder / decoder
    @torch.no_grad()
    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 NCHW [B, 3, H, W] -> fp32 NCHW [B, 2 * latent, H/8, W/8] (``vae.encoder``)."""
        W = self._weights()
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise ValueError("vae.encoder expects an fp32 CUDA NCHW tensor (no CPU fallback)")
        cfg = self.config
        boc, L = cfg.block_out_channels, cfg.layers_per_block
        B, Cin, H, Wd = x.shape
        f = 2 ** (len(boc) - 1)
        if Cin != cfg.in_channels or H % f or Wd % f:
            raise ValueError(f"vae.encoder: expected [B, {cfg.in_channels}, H, W] with H, W multiples of {f}")
        self._begin(B, x.device)
        st = self._arena.take(B, 128)
        h = _Act(ops.conv_in_direct(x.contiguous(), W["enc.in.wT"], bias=W["enc.in.b"], out_dtype=self._adt, stats=st), st)
        for i, c in enumerate(boc):
            for j in range(L):
                h = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", h, c, B)
            if i < len(boc) - 1:
                p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                h = self._conv(h.t, W[p + ".w"], W[p + ".b"], 3, B, stride=2, asym_pad=True)
        h = self._mid("encoder.mid_block", h, B)
        a = ops.groupnorm_silu(h.t, h.stats, W["enc.out.g"], W["enc.out.b"], eps=1e-6)
        o = ops.conv2d_igemm(a, W["enc.out.w"], W["enc.out.bias"], ksize=3, cout_valid=64)
        return ops.nhwc16_to_nchw(o, 2 * cfg.latent_channels)

    @torch.no_grad()
    def _decode(self, z: torch.Tensor, out_w=None, out_b=None) -> torch.Tensor:
        """fp32 NCHW [B, latent, h, w] -> fp32 NCHW [B, 3, 8h, 8w] (``vae.decoder``); ``out_w`` / ``out_b`` replace
        conv_out's weights (decode_mean)."""
        W = self._weights()
        if not (z.is_cuda and z.dtype == torch.float32 and z.dim() == 4):
            raise ValueError("vae.decoder expects an fp32 CUDA NCHW tensor (no CPU fallback)")
        cfg = self.config
        boc, L = cfg.block_out_channels, cfg.layers_per_block
        B, Cz, Hh, Ww = z.shape
        if Cz != cfg.latent_channels:
            raise ValueError(f"vae.decoder: expected {cfg.latent_channels} latent channels, got {Cz}")
        self._begin(B, z.device)
        x = torch.zeros((B, Hh, Ww, 64), dtype=self._adt, device=z.device)
        ops.nchw_to_nhwc16(z, x)
        h = self._conv(x, W["dec.in.w"], W["dec.in.b"], 3, B)
        h = self._mid("decoder.mid_block", h, B)
        rev = boc[::-1]
        for i, c in enumerate(rev):
            for j in range(L + 1):
                h = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", h, c, B)
            if i < len(rev) - 1:
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                Bn, H2, W2, Cc = h.t.shape
                up = torch.empty((Bn, 2 * H2, 2 * W2, Cc), dtype=h.t.dtype, device=h.t.device)
                st = self._arena.take(B, Cc)
                for ph in range(4):
                    ops.conv2d_igemm(h.t, W[p + f".w{ph}"], W[p + ".b"], ksize=3, out=up, stats=st, up2x_phase=ph + 1)
                h = _Act(up, st)
        ow = W["dec.out.w"] if out_w is None else out_w
        ob = W["dec.out.bias"] if out_b is None else out_b
        return ops.conv_out_fused(h.t, h.stats, W["dec.out.g"], W["dec.out.b"], ow, ob, ow.shape[0], eps=1e-6)

    # ------------------------------------------------------------------ fused forms of the pipeline's call sequences
    @torch.no_grad()
    def encode_latent(self, x: torch.Tensor, scale: float) -> torch.Tensor:
        """``encode_rgb`` of the reference (custom_pipelines.py:47-59) in one pass: encoder -> the MEAN half of
        quant_conv -> times the latent scale, with the scale and the row selection folded into the 1x1 conv."""
        W = self._weights()
        lc = self.config.latent_channels
        key = ("enc_fold", float(scale))
        if key not in self._fold:
            self._fold[key] = ((W["quant_conv.w"][:lc] * scale).contiguous(), (W["quant_conv.b"][:lc] * scale).contiguous())
        w, b = self._fold[key]
        return ops.conv1x1_nchw_small(self._encode(x), w, b)

    @torch.no_grad()
    def decode_mean(self, latent: torch.Tensor, scale: float) -> torch.Tensor:
        """``__decode_depth`` of the reference (custom_pipelines.py:255-272) in one pass: latent / scale ->
        post_quant_conv -> decoder -> mean over the output channels, with 1 / scale folded into the 1x1 conv and the
        channel mean folded into conv_out (the mean of the output channels of a conv is the conv with the averaged
        kernel) - one output plane instead of three and no separate reduction."""
        W = self._weights()
        key = ("dec_fold", float(scale))
        if key not in self._fold:
            self._fold[key] = ((W["post_quant_conv.w"] / scale).contiguous(), W["post_quant_conv.b"],
                               W["dec.out.w"].mean(0, keepdim=True).contiguous(),
                               W["dec.out.bias"].mean(0, keepdim=True).contiguous())
        w, b, wo, bo = self._fold[key]
        return self._decode(ops.conv1x1_nchw_small(latent, w, b), out_w=wo, out_b=bo)

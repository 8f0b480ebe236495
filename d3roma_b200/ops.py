"""Tensor-level wrappers over the C ABI (torch is used only for device memory and streams).

Every function takes CUDA tensors, extracts raw device pointers and calls ``libd3roma_b200.so``
on torch's current stream.  Activations are bf16 NHWC ``[B, H, W, C]`` (last dim contiguous; a
channel slice of a wider buffer is allowed — the pixel pitch is taken from ``stride(2)``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    """C-ABI dtype code of a 16-bit activation / weight tensor."""
    if t.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    if t.dtype == torch.float16:
        return _lib.DTYPE_F16
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    raise ValueError(f"expected a bf16, fp16 or fp32 tensor, got {t.dtype}")


def _check_nhwc(t: torch.Tensor, name: str) -> None:
    if not (t.is_cuda and t.dtype in (torch.bfloat16, torch.float16, torch.float32) and t.dim() == 4
            and t.stride(3) == 1):
        raise ValueError(f"{name}: expected a CUDA bf16/fp16/fp32 NHWC tensor with contiguous channels")
    B, H, W, _ = t.shape
    p = t.stride(2)
    if t.stride(1) != W * p or (B > 1 and t.stride(0) != H * W * p):
        raise ValueError(f"{name}: expected dense pixels (strides {t.stride()})")


def pack_conv_weight(w_oihw: torch.Tensor, shortcut_w: Optional[torch.Tensor] = None,
                     dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """OIHW fp32/bf16 conv weight (+ optional [Cout, Csc(,1,1)] 1x1 shortcut weight) -> bf16 [Cout, Ktot]
    with k = (kh*KW + kw)*Cin + c, shortcut channels appended (layout the igemm kernel expects)."""
    co, ci, kh, kw = w_oihw.shape
    w = w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
    if shortcut_w is not None:
        w = torch.cat([w, shortcut_w.reshape(co, -1).to(w.dtype)], dim=1)
    return w.to(dtype).contiguous()


def pack_upsample_weights(w_oihw: torch.Tensor, dtype: torch.dtype = torch.bfloat16):
    """Weights of the 4 sub-pixel phases of ``conv3x3(nearest_upsample_2x(x))``: for output parity (py, px) the 3x3
    kernel collapses onto a 2x2 input neighbourhood (rows {i+py-1, i+py}, cols {j+px-1, j+px}); taps that land on the
    same input pixel are summed in fp32.  Returns [phase] -> [Cout, 4*Cin], tap-major (dy, dx)."""
    w = w_oihw.float()
    groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}  # parity -> (taps hitting the first input row/col, the second)
    out = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for rows in groups[py]:
                for cols in groups[px]:
                    taps.append(w[:, :, rows][:, :, :, cols].sum(dim=(2, 3)))   # [Cout, Cin]
            out.append(torch.stack(taps, dim=1).reshape(w.shape[0], -1).to(dtype).contiguous())
    return out


def conv2d_igemm(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, *, ksize: int, stride: int = 1,
                 sc1: Optional[torch.Tensor] = None, sc2: Optional[torch.Tensor] = None,
                 residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 stats: Optional[torch.Tensor] = None, up2x_phase: int = 0, cout_valid: int = 0,
                 asym_pad: bool = False) -> torch.Tensor:
    """tcgen05 implicit-GEMM conv (see ``d3_conv2d_igemm`` in include/d3roma_b200.h).  ``cout_valid``: the output has
    only that many channels; the remaining rows of ``w_packed`` / ``bias`` are zero padding up to a multiple of 128."""
    _check_nhwc(x, "x")
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    Cv = cout_valid or Cout
    Ho, Wo = H // stride, W // stride
    if up2x_phase:
        Ho, Wo = 2 * H, 2 * W
    if out is None:
        out = torch.empty((B, Ho, Wo, Cv), dtype=x.dtype, device=x.device)
    if tuple(out.shape) != (B, Ho, Wo, Cv):
        raise ValueError(f"out has shape {tuple(out.shape)}, expected {(B, Ho, Wo, Cv)}")
    _check_nhwc(out, "out")
    p = _lib.ConvParams()
    p.x, p.B, p.H, p.W, p.Cin, p.x_pitch = x.data_ptr(), B, H, W, Cin, x.stride(2)
    p.ksize, p.stride = ksize, stride
    p.w, p.Cout = w_packed.data_ptr(), Cout
    c1 = c2 = 0
    if sc1 is not None:
        _check_nhwc(sc1, "sc1")
        c1 = sc1.shape[3]
        p.sc1, p.C1, p.sc1_pitch = sc1.data_ptr(), c1, sc1.stride(2)
    if sc2 is not None:
        _check_nhwc(sc2, "sc2")
        c2 = sc2.shape[3]
        p.sc2, p.C2, p.sc2_pitch = sc2.data_ptr(), c2, sc2.stride(2)
    ntaps = 4 if up2x_phase else ksize * ksize
    if w_packed.shape[1] != ntaps * Cin + c1 + c2 or w_packed.dtype != x.dtype:
        raise ValueError("w_packed shape/dtype mismatch")
    if bias.dtype != torch.float32 or bias.shape[-1] != Cout or bias.stride(-1) != 1:
        raise ValueError("bias must be fp32 [Cout] or [B, Cout]")
    p.bias, p.bias_bstride = bias.data_ptr(), (bias.stride(0) if bias.dim() == 2 and bias.shape[0] > 1 else 0)
    if residual is not None:
        _check_nhwc(residual, "residual")
        p.residual, p.res_pitch = residual.data_ptr(), residual.stride(2)
    p.out, p.out_pitch = out.data_ptr(), out.stride(2)
    if stats is not None:
        if stats.dtype != torch.float64 or stats.numel() != B * (Cout // 4) * 2:
            raise ValueError("stats must be fp64 [B, Cout/4, 2]")
        p.stats = stats.data_ptr()
    p.dtype = _dt(x)
    p.up2x_phase = int(up2x_phase)
    p.Cout_valid = int(cout_valid)
    p.asym_pad = int(bool(asym_pad))
    for t_ in (sc1, sc2, residual, out):
        if t_ is not None and t_.dtype != x.dtype:
            raise ValueError("all 16-bit operands of a conv must share one dtype")
    ws_bytes = _lib.lib().d3_conv2d_workspace_bytes(C.byref(p))
    if ws_bytes:   # split-K scratch for layers with too few tiles to fill the SMs (torch's allocator: graph-capturable)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
    _lib.check(_lib.lib().d3_conv2d_igemm(C.byref(p), _stream()))
    return out


def groupnorm_silu(x1: torch.Tensor, stats1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *,
                   x2: Optional[torch.Tensor] = None, stats2: Optional[torch.Tensor] = None, silu: bool = True,
                   groups: int = 32, eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over cat(x1, x2) using the producers' quad statistics (``d3_groupnorm_silu``)."""
    _check_nhwc(x1, "x1")
    B, H, W, C1 = x1.shape
    C2 = 0
    if x2 is not None:
        _check_nhwc(x2, "x2")
        C2 = x2.shape[3]
    if out is None:
        out = torch.empty((B, H, W, C1 + C2), dtype=x1.dtype, device=x1.device)
    _lib.check(_lib.lib().d3_groupnorm_silu(
        x1.data_ptr(), C1, x1.stride(2), stats1.data_ptr(), _ptr(x2), C2, x2.stride(2) if x2 is not None else 0,
        _ptr(stats2), gamma.data_ptr(), beta.data_ptr(), B, H * W, groups, eps, int(silu), out.data_ptr(),
        out.stride(2), _dt(x1), _stream()))
    return out


def groupnorm_silu_dual(x: torch.Tensor, stats: torch.Tensor, gamma_a: torch.Tensor, beta_a: torch.Tensor,
                        gamma_b: torch.Tensor, beta_b: torch.Tensor, out_b: torch.Tensor, *, groups_a: int = 32,
                        groups_b: int = 32, silu_a: bool = True, silu_b: bool = True, eps: float = 1e-5,
                        out_a: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One read of ``x``, two normalised outputs (``d3_groupnorm_silu_dual``): returns ``out_a``; ``out_b`` may be a
    channel-range view of a wider NHWC buffer (the up-block's concatenated GroupNorm output)."""
    _check_nhwc(x, "x")
    B, H, W, C = x.shape
    if out_a is None:
        out_a = torch.empty((B, H, W, C), dtype=x.dtype, device=x.device)
    if out_b.shape != (B, H, W, C) or out_b.dtype != x.dtype or out_b.stride(3) != 1:
        raise ValueError("groupnorm_silu_dual: out_b must be a [B,H,W,C] (view of an) NHWC tensor of x's dtype")
    if gamma_b.numel() != C or beta_b.numel() != C or not (gamma_b.is_contiguous() and beta_b.is_contiguous()):
        raise ValueError("groupnorm_silu_dual: gamma_b / beta_b must be contiguous [C] slices")
    _lib.check(_lib.lib().d3_groupnorm_silu_dual(
        x.data_ptr(), C, x.stride(2), stats.data_ptr(), B, H * W, eps, gamma_a.data_ptr(), beta_a.data_ptr(), groups_a,
        int(silu_a), out_a.data_ptr(), out_a.stride(2), gamma_b.data_ptr(), beta_b.data_ptr(), groups_b, int(silu_b),
        out_b.data_ptr(), out_b.stride(2), _dt(x), _stream()))
    return out_a


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    _check_nhwc(x, "x")
    B, H, W, Cc = x.shape
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().d3_upsample_nearest2x(x.data_ptr(), B, H, W, Cc, x.stride(2), out.data_ptr(),
                                                out.stride(2), _stream()))
    return out


def im2col_input(sample: torch.Tensor, cond: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                 dtype: torch.dtype = torch.bfloat16):
    """fp32 NCHW sample [B,Cs,H,W] (+ cond [B,Cc,H,W]) -> bf16 [B,H,W,128] im2col rows for conv_in."""
    B, Cs, H, W = sample.shape
    Cc = 0 if cond is None else cond.shape[1]
    for t in (sample, cond):
        if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("im2col_input expects contiguous fp32 CUDA NCHW tensors")
    if out is None:
        out = torch.empty((B, H, W, 128), dtype=dtype, device=sample.device)
    _lib.check(_lib.lib().d3_im2col_input(sample.data_ptr(), Cs, _ptr(cond), Cc, B, H, W, out.data_ptr(), _dt(out),
                                          _stream()))
    return out


def pack_conv_in_direct(w_oihw: torch.Tensor, c0: int, c1: int) -> torch.Tensor:
    """fp32 OIHW conv_in weight [128, Cin, 3, 3] -> [(c1-c0)*9][128] fp32, k = (c-c0)*9 + kh*3 + kw, for
    ``conv_in_direct`` over the input channels c0..c1-1."""
    w = w_oihw[:, c0:c1].float()
    return w.permute(1, 2, 3, 0).reshape((c1 - c0) * 9, w.shape[0]).contiguous()


def conv_in_direct(planes: torch.Tensor, wT: torch.Tensor, *, base: Optional[torch.Tensor] = None,
                   bias: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
                   out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Direct 3x3 conv of fp32 NCHW ``planes`` [B,C,H,W] with 128 output channels (``d3_conv_in_direct``).
    ``out_dtype=None``: fp32 NHWC result (``bias`` added).  ``out_dtype`` bf16/fp16: (``base`` or ``bias``) + conv,
    rounded, plus GroupNorm quad statistics into ``stats`` when given.  ``base`` is the per-loop conditioning term, NHWC
    [B,H,W,128] in fp32 or fp16 (the UNet keeps it in fp16: 4 instead of 6 bytes per element and step)."""
    if not (planes.is_cuda and planes.dtype == torch.float32 and planes.is_contiguous()):
        raise ValueError("conv_in_direct expects a contiguous fp32 CUDA NCHW tensor")
    B, C, H, W = planes.shape
    if wT.shape != (C * 9, 128) or wT.dtype != torch.float32:
        raise ValueError(f"conv_in_direct: weight must be fp32 [{C * 9}, 128]")
    base_dt = _lib.DTYPE_F32
    if base is not None:
        if base.shape != (B, H, W, 128) or base.dtype not in (torch.float32, torch.float16) or not base.is_contiguous():
            raise ValueError("conv_in_direct: base must be contiguous fp32 or fp16 [B,H,W,128]")
        base_dt = _dt(base)
    if out is not None:
        out_dtype = None if out.dtype == torch.float32 else out.dtype
    if out_dtype is None:
        if out is None:
            out = torch.empty((B, H, W, 128), dtype=torch.float32, device=planes.device)
        _lib.check(_lib.lib().d3_conv_in_direct(planes.data_ptr(), C, wT.data_ptr(), _ptr(base), base_dt, _ptr(bias), B,
                                                H, W, out.data_ptr(), None, 0, None, _stream()))
        return out
    if out is None:
        out = torch.empty((B, H, W, 128), dtype=out_dtype, device=planes.device)
    _lib.check(_lib.lib().d3_conv_in_direct(planes.data_ptr(), C, wT.data_ptr(), _ptr(base), base_dt, _ptr(bias), B, H,
                                            W, None, out.data_ptr(), _dt(out), _ptr(stats), _stream()))
    return out


def conv_out_fused(x: torch.Tensor, stats: torch.Tensor, gamma, beta, w, bias, dc: int, eps: float = 1e-5,
                   out: Optional[torch.Tensor] = None, ddim=None):
    """conv_norm_out + SiLU + conv_out.  ``ddim=(sample, coeffs)``: the DDIM update (eta = 0) is applied to the model
    output in the same pass (``d3_conv_out_fused_ddim``) and ``(prev_sample, pred_original)`` is returned instead of the
    model output."""
    _check_nhwc(x, "x")
    B, H, W, Cc = x.shape
    if Cc != 128:
        raise ValueError("conv_out_fused expects 128 input channels")
    ws = None
    if x.dtype != torch.float32:  # scratch for the 9*dc fp32 tap maps (torch's caching allocator; graph-capturable)
        ws = torch.empty((_lib.lib().d3_conv_out_workspace_bytes(B, H, W, dc),), dtype=torch.uint8, device=x.device)
    if ddim is not None:
        sample, coeffs = ddim
        s = _f32c(sample, "sample")
        if s.shape != (B, dc, H, W) or x.dtype == torch.float32:
            raise ValueError("conv_out_fused(ddim=): sample must be fp32 [B,dc,H,W] and the activations 16-bit")
        prev, orig = torch.empty_like(s), torch.empty_like(s)
        _lib.check(_lib.lib().d3_conv_out_fused_ddim(
            x.data_ptr(), x.stride(2), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w.data_ptr(),
            bias.data_ptr(), B, H, W, dc, eps, _ptr(out), _ptr(ws), _dt(x), s.data_ptr(), C.byref(coeffs),
            prev.data_ptr(), orig.data_ptr(), _stream()))
        return prev, orig
    if out is None:
        out = torch.empty((B, dc, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_conv_out_fused(x.data_ptr(), x.stride(2), stats.data_ptr(), gamma.data_ptr(),
                                            beta.data_ptr(), w.data_ptr(), bias.data_ptr(), B, H, W, dc, eps,
                                            out.data_ptr(), _ptr(ws), _dt(x), _stream()))
    return out


def time_embed(t: torch.Tensor, w1t, b1, w2t, b2) -> torch.Tensor:
    nt = t.numel()
    out = torch.empty((nt, 512), dtype=torch.float32, device=t.device)
    _lib.check(_lib.lib().d3_time_embed(t.data_ptr(), nt, w1t.data_ptr(), b1.data_ptr(), w2t.data_ptr(),
                                        b2.data_ptr(), out.data_ptr(), _stream()))
    return out


def temb_project_all(silu_emb: torch.Tensor, w_all: torch.Tensor, bias_all: torch.Tensor) -> torch.Tensor:
    nt, rows = silu_emb.shape[0], w_all.shape[0]
    out = torch.empty((nt, rows), dtype=torch.float32, device=silu_emb.device)
    _lib.check(_lib.lib().d3_temb_project_all(silu_emb.data_ptr(), nt, w_all.data_ptr(), bias_all.data_ptr(), rows,
                                              out.data_ptr(), _stream()))
    return out


def attention_hd8(qkv: torch.Tensor, C_: int) -> torch.Tensor:
    _check_nhwc(qkv, "qkv")
    B, H, W, _ = qkv.shape
    out = torch.empty((B, H, W, C_), dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.lib().d3_attention_hd8(qkv.data_ptr(), qkv.stride(2), B, H * W, C_, out.data_ptr(),
                                           out.stride(2), _dt(qkv), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# latent family (UNet2DConditionModel / AutoencoderKL): ops of csrc/ldm_ops.cu
# ------------------------------------------------------------------------------------------------
def pad_rows(w: torch.Tensor, multiple: int = 128) -> torch.Tensor:
    """Zero-pad dim 0 (output channels of a packed weight or bias) up to a multiple of the 128-channel MMA tile."""
    r = (-w.shape[0]) % multiple
    if r == 0:
        return w.contiguous()
    return torch.cat([w, w.new_zeros((r,) + tuple(w.shape[1:]))], 0).contiguous()


def gn_stats(x1: torch.Tensor, chan: torch.Tensor, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-channel fp64 sum / sum of squares of cat(x1, x2) accumulated into ``chan`` [B, C, 2] (zero on entry)."""
    _check_nhwc(x1, "x1")
    B, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[3]
    if x2 is not None:
        _check_nhwc(x2, "x2")
    if chan.dtype != torch.float64 or chan.numel() != B * (C1 + C2) * 2:
        raise ValueError("gn_stats: chan must be fp64 [B, C, 2]")
    _lib.check(_lib.lib().d3_gn_stats(x1.data_ptr(), C1, x1.stride(2), _ptr(x2), C2, x2.stride(2) if x2 is not None else 0,
                                      B, H * W, chan.data_ptr(), _dt(x1), _stream()))
    return chan


def gn_apply(x1: torch.Tensor, chan: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *,
             x2: Optional[torch.Tensor] = None, groups: int = 32, eps: float = 1e-5, silu: bool = True,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm (+SiLU) of cat(x1, x2) from the per-channel moments of ``gn_stats`` (any group width)."""
    B, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[3]
    if out is None:
        out = torch.empty((B, H, W, C1 + C2), dtype=x1.dtype, device=x1.device)
    _lib.check(_lib.lib().d3_gn_apply(x1.data_ptr(), C1, x1.stride(2), _ptr(x2), C2, x2.stride(2) if x2 is not None else 0,
                                      chan.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, H * W, groups, eps, int(silu),
                                      out.data_ptr(), out.stride(2), _dt(x1), _stream()))
    return out


def groupnorm_one_pass(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, x2: Optional[torch.Tensor] = None,
                       groups: int = 32, eps: float = 1e-5, silu: bool = True) -> torch.Tensor:
    """GroupNorm (+SiLU) of cat(x1, x2) in one launch (``d3_groupnorm_one_pass``; L2-resident tensors)."""
    _check_nhwc(x1, "x1")
    B, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[3]
    out = torch.empty((B, H, W, C1 + C2), dtype=x1.dtype, device=x1.device)
    _lib.check(_lib.lib().d3_groupnorm_one_pass(x1.data_ptr(), C1, x1.stride(2), _ptr(x2), C2,
                                                x2.stride(2) if x2 is not None else 0, gamma.data_ptr(), beta.data_ptr(),
                                                B, H * W, groups, eps, int(silu), out.data_ptr(), out.stride(2), _dt(x1),
                                                _stream()))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last dim of a 16-bit [..., C] tensor whose rows are ``x.stride(-2)`` apart."""
    C_ = x.shape[-1]
    M = x.numel() // C_
    out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().d3_layernorm(x.data_ptr(), x.stride(-2), gamma.data_ptr(), beta.data_ptr(), M, C_, eps,
                                       out.data_ptr(), C_, _dt(x), _stream()))
    return out


def geglu(x: torch.Tensor) -> torch.Tensor:
    """[..., 2D] -> [..., D]: x[..., :D] * gelu(x[..., D:])."""
    D = x.shape[-1] // 2
    M = x.numel() // (2 * D)
    out = torch.empty(tuple(x.shape[:-1]) + (D,), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().d3_geglu(x.data_ptr(), x.stride(-2), M, D, out.data_ptr(), D, _dt(x), _stream()))
    return out


def attention_hd64(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """q, k, v: [B, H, W, 64*heads] (channel-slice views of a fused QKV tensor are fine) -> [B, H, W, 64*heads]."""
    B, H, W, C_ = q.shape
    if C_ != 64 * heads:
        raise ValueError("attention_hd64: head_dim must be 64")
    out = torch.empty((B, H, W, C_), dtype=q.dtype, device=q.device)
    _lib.check(_lib.lib().d3_attention_hd64(q.data_ptr(), q.stride(2), k.data_ptr(), k.stride(2), v.data_ptr(),
                                            v.stride(2), B, H * W, heads, out.data_ptr(), C_, _dt(q), _stream()))
    return out


def attention_hd64_tc(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """``attention_hd64`` on the tcgen05 / TMEM / TMA path (``d3_attention_hd64_tc``)."""
    B, H, W, C_ = q.shape
    if C_ != 64 * heads:
        raise ValueError("attention_hd64_tc: head_dim must be 64")
    out = torch.empty((B, H, W, C_), dtype=q.dtype, device=q.device)
    ws = torch.empty((_lib.lib().d3_attention_hd64_tc_workspace_bytes(B, H * W, heads),), dtype=torch.uint8, device=q.device)
    _lib.check(_lib.lib().d3_attention_hd64_tc(q.data_ptr(), q.stride(2), k.data_ptr(), k.stride(2), v.data_ptr(),
                                               v.stride(2), B, H * W, heads, out.data_ptr(), C_, ws.data_ptr(), _dt(q),
                                               _stream()))
    return out


def cross_attention_small(q: torch.Tensor, kc: torch.Tensor, vc: torch.Tensor) -> torch.Tensor:
    """q 16-bit [B, H, W, C]; kc, vc fp32 [B, L, C] (L <= 8) -> [B, H, W, C]."""
    B, H, W, C_ = q.shape
    L = kc.shape[1]
    if kc.shape != (B, L, C_) or vc.shape != (B, L, C_) or kc.dtype != torch.float32 or not (kc.is_contiguous() and vc.is_contiguous()):
        raise ValueError("cross_attention_small: kc / vc must be contiguous fp32 [B, L, C]")
    out = torch.empty((B, H, W, C_), dtype=q.dtype, device=q.device)
    _lib.check(_lib.lib().d3_cross_attention_small(q.data_ptr(), q.stride(2), kc.data_ptr(), vc.data_ptr(), B, H * W, L,
                                                   C_, out.data_ptr(), C_, _dt(q), _stream()))
    return out


def linear_small(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act_in: int = 0,
                 act_out: int = 0) -> torch.Tensor:
    """fp32 [n, K] x fp32 [rows, K]^T (+ bias) -> [n, rows]; act 0 = identity, 1 = SiLU, 2 = GELU (erf), 3 = quick GELU."""
    n, K = x.shape
    rows = w.shape[0]
    if w.shape[1] != K or x.dtype != torch.float32 or w.dtype != torch.float32 or not (x.is_contiguous() and w.is_contiguous()):
        raise ValueError("linear_small: contiguous fp32 [n, K] and [rows, K] expected")
    out = torch.empty((n, rows), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_linear_small(x.data_ptr(), n, K, w.data_ptr(), _ptr(bias), rows, act_in, act_out,
                                          out.data_ptr(), rows, _stream()))
    return out


def embed_tokens(ids: torch.Tensor, token_embedding: torch.Tensor, position_embedding: torch.Tensor) -> torch.Tensor:
    """int64 [B, L] token ids -> fp32 [B*L, D]: token_embedding[id] + position_embedding[position]."""
    if ids.dtype != torch.int64 or ids.dim() != 2 or not ids.is_contiguous():
        raise ValueError("embed_tokens: contiguous int64 [B, L] ids expected")
    B, L = ids.shape
    V, D = token_embedding.shape
    lo, hi = int(ids.min()), int(ids.max())
    if lo < 0 or hi >= V:
        raise IndexError(f"embed_tokens: token id out of range [0, {V}) (got {lo} .. {hi})")
    out = torch.empty((B * L, D), dtype=torch.float32, device=ids.device)
    _lib.check(_lib.lib().d3_embed_tokens(ids.data_ptr(), B * L, L, _f32c(token_embedding, "token_embedding").data_ptr(),
                                          V, _f32c(position_embedding, "position_embedding").data_ptr(),
                                          position_embedding.shape[0], D, out.data_ptr(), _stream()))
    return out


def layernorm_f32(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the rows of a contiguous fp32 [M, C] matrix."""
    x = _f32c(x, "x")
    M, C_ = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib().d3_layernorm_f32(x.data_ptr(), M, C_, _f32c(gamma, "gamma").data_ptr(),
                                           _f32c(beta, "beta").data_ptr(), eps, out.data_ptr(), _stream()))
    return out


def causal_attention_f32(qkv: torch.Tensor, B: int, L: int, heads: int, scale: float) -> torch.Tensor:
    """Fused fp32 [B*L, 3*D] rows (q | k | v) -> fp32 [B*L, D]: per-head softmax over the keys j <= i."""
    qkv = _f32c(qkv, "qkv")
    D = qkv.shape[1] // 3
    if qkv.shape != (B * L, 3 * D) or D % heads:
        raise ValueError("causal_attention_f32: qkv must be [B*L, 3*D] with D divisible by heads")
    out = torch.empty((B * L, D), dtype=torch.float32, device=qkv.device)
    es = qkv.element_size()
    _lib.check(_lib.lib().d3_causal_attention_f32(qkv.data_ptr(), qkv.data_ptr() + D * es, qkv.data_ptr() + 2 * D * es,
                                                  3 * D, B, L, heads, D // heads, scale, out.data_ptr(), D, _stream()))
    return out


def timestep_sincos(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0) -> torch.Tensor:
    out = torch.empty((t.numel(), dim), dtype=torch.float32, device=t.device)
    _lib.check(_lib.lib().d3_timestep_sincos(t.data_ptr(), t.numel(), dim, int(flip_sin_to_cos), freq_shift,
                                             out.data_ptr(), _stream()))
    return out


def nchw_to_nhwc16(x: torch.Tensor, out: torch.Tensor, c_off: int = 0) -> torch.Tensor:
    """fp32 NCHW [B, C, H, W] -> channels [c_off, c_off + C) of the 16-bit NHWC tensor ``out``."""
    x = _f32c(x, "x")
    B, C_, H, W = x.shape
    _check_nhwc(out, "out")
    _lib.check(_lib.lib().d3_nchw_to_nhwc16(x.data_ptr(), B, C_, H, W, out.data_ptr(), out.stride(2), c_off, _dt(out),
                                            _stream()))
    return out


def nhwc16_to_nchw(x: torch.Tensor, C_: int, c_off: int = 0) -> torch.Tensor:
    _check_nhwc(x, "x")
    B, H, W, _ = x.shape
    out = torch.empty((B, C_, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_nhwc16_to_nchw(x.data_ptr(), x.stride(2), c_off, B, C_, H, W, out.data_ptr(), _dt(x),
                                            _stream()))
    return out


def resize_nearest(x: torch.Tensor, Ho: int, Wo: int, Hbuf: Optional[int] = None, Wbuf: Optional[int] = None):
    """F.interpolate(size=(Ho, Wo), mode="nearest") into the top-left window of a zero-filled [B, Hbuf, Wbuf, C] buffer."""
    _check_nhwc(x, "x")
    B, H, W, C_ = x.shape
    Hbuf, Wbuf = Hbuf or Ho, Wbuf or Wo
    out = torch.empty((B, Hbuf, Wbuf, C_), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().d3_resize_nearest_nhwc16(x.data_ptr(), B, H, W, C_, x.stride(2), out.data_ptr(), Ho, Wo, Hbuf,
                                                   Wbuf, C_, _stream()))
    return out


def softmax_rows_(s: torch.Tensor, n_cols: int, scale: float) -> torch.Tensor:
    """In-place softmax(scale * s) over the first ``n_cols`` columns of a 16-bit [rows, pitch] matrix (rest zeroed)."""
    rows, pitch = s.shape
    if not s.is_contiguous():
        raise ValueError("softmax_rows_: contiguous [rows, pitch] expected")
    _lib.check(_lib.lib().d3_softmax_rows(s.data_ptr(), rows, n_cols, pitch, scale, _dt(s), _stream()))
    return s


def transpose_pad(x: torch.Tensor, Np: int) -> torch.Tensor:
    """16-bit [N, C] (row pitch ``x.stride(0)``) -> [C, Np] with zero columns N..Np-1."""
    N, C_ = x.shape
    out = torch.empty((C_, Np), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().d3_transpose_pad16(x.data_ptr(), N, C_, x.stride(0), out.data_ptr(), Np, _stream()))
    return out


def conv1x1_nchw_small(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """fp32 NCHW 1x1 conv with <= 8 channels on both sides; ``w`` fp32 [Cout, Cin]."""
    x = _f32c(x, "x")
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_conv1x1_nchw_small(x.data_ptr(), B, Cin, Cout, H * W, w.data_ptr(), _ptr(bias),
                                                out.data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# scheduler / guidance / pre-post wrappers (fp32 NCHW tensors)
# ------------------------------------------------------------------------------------------------
def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float32):
        raise ValueError(f"{name}: expected an fp32 CUDA tensor (got {t.dtype} on {t.device}); "
                         "the d3roma_b200 path has no CPU fallback")
    return t if t.is_contiguous() else t.contiguous()


def sched_ddim_step(model_output, sample, coeffs: "_lib.DDIMCoeffs", variance_noise=None, want_mean=False,
                    want_dir=False):
    m, s = _f32c(model_output, "model_output"), _f32c(sample, "sample")
    noise = None if variance_noise is None else _f32c(variance_noise, "variance_noise")
    prev, orig = torch.empty_like(s), torch.empty_like(s)
    mean = torch.empty_like(s) if want_mean else None
    dirv = torch.empty_like(s) if want_dir else None
    _lib.check(_lib.lib().d3_sched_ddim_step(m.data_ptr(), s.data_ptr(), _ptr(noise), C.byref(coeffs), s.numel(),
                                             prev.data_ptr(), orig.data_ptr(), _ptr(mean), _ptr(dirv), _stream()))
    return prev, orig, mean, dirv


def sched_ddpm_step(model_output, sample, coeffs: "_lib.DDPMCoeffs", variance_noise=None, want_mean=False):
    m, s = _f32c(model_output, "model_output"), _f32c(sample, "sample")
    noise = None if variance_noise is None else _f32c(variance_noise, "variance_noise")
    prev, orig = torch.empty_like(s), torch.empty_like(s)
    mean = torch.empty_like(s) if want_mean else None
    _lib.check(_lib.lib().d3_sched_ddpm_step(m.data_ptr(), s.data_ptr(), _ptr(noise), C.byref(coeffs), s.numel(),
                                             prev.data_ptr(), orig.data_ptr(), _ptr(mean), _stream()))
    return prev, orig, mean


def lincomb(a: float, x, b: float = 1.0, y=None, c: float = 0.0, z=None):
    """a*x (+ b*y) (+ c*z) with every product / sum rounded separately."""
    x = _f32c(x, "x")
    y = None if y is None else _f32c(y, "y")
    z = None if z is None else _f32c(z, "z")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().d3_lincomb(a, x.data_ptr(), b, _ptr(y), c, _ptr(z), x.numel(), out.data_ptr(), _stream()))
    return out


def guidance_grad(disp_norm, left, right, min_disp: float, max_disp: float, shift: float = 0.0,
                  w_ssim: float = 0.95):
    """Reference ``FlowGuidance.grad`` semantic: d(loss)/d(camera-normalised disparity) (``d3_guidance_grad``)."""
    d, l, r = _f32c(disp_norm, "disp_norm"), _f32c(left, "left"), _f32c(right, "right")
    B, _, H, W = d.shape
    margin = r.shape[-1] - W
    if l.shape != r.shape or l.shape[1] != 3 or margin < 0:
        raise ValueError("guidance_grad: left/right must be [B,3,H,W+margin]")
    out = torch.empty_like(d)
    _lib.check(_lib.lib().d3_guidance_grad(d.data_ptr(), l.data_ptr(), r.data_ptr(), B, H, W, margin, min_disp,
                                           max_disp, shift, w_ssim, out.data_ptr(), _stream()))
    return out


def guidance_step(x0, left, right, disp_scale: float, disp_offset: float, min_disp: float, max_disp: float,
                  w_ssim: float = 0.95, *, step: float = 0.0, want_x0: bool = False, want_grad: bool = False,
                  ca: float = 0.0, aux1=None, cb: float = 0.0, aux2=None, cc: float = 0.0):
    """``d3_guidance_step``: photometric gradient w.r.t. x0 through ``disp = offset + scale * sum_c x0_c`` and,
    optionally, ``x0' = x0 - step * g`` and ``out2 = ca * x0' + cb * aux1 (+ cc * aux2)`` in the same launch.
    Returns ``(x0' or None, out2 or None, g or None)``."""
    x, l, r = _f32c(x0, "x0"), _f32c(left, "left"), _f32c(right, "right")
    B, Cc, H, W = x.shape
    margin = r.shape[-1] - W
    if l.shape != r.shape or l.shape[1] != 3 or l.shape[0] != B or l.shape[2] != H or margin < 0:
        raise ValueError("guidance_step: left/right must be [B,3,H,W+margin]")
    p = _lib.GuidanceParams()
    p.x0, p.left, p.right = x.data_ptr(), l.data_ptr(), r.data_ptr()
    g = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device) if want_grad else None
    xo = torch.empty_like(x) if want_x0 else None
    o2 = None
    if aux1 is not None:
        a1 = _f32c(aux1, "aux1")
        if a1.shape != x.shape:
            raise ValueError("guidance_step: aux1 must have the shape of x0")
        o2 = torch.empty_like(x)
        p.out2, p.aux1 = o2.data_ptr(), a1.data_ptr()
        if aux2 is not None:
            a2 = _f32c(aux2, "aux2")
            p.aux2 = a2.data_ptr()
    p.grad, p.x0_out = _ptr(g), _ptr(xo)
    p.B, p.C, p.H, p.W, p.margin = B, Cc, H, W, margin
    p.disp_scale, p.disp_offset, p.min_disp, p.max_disp, p.w_ssim = disp_scale, disp_offset, min_disp, max_disp, w_ssim
    p.step, p.ca, p.cb, p.cc = step, ca, cb, cc
    _lib.check(_lib.lib().d3_guidance_step(C.byref(p), _stream()))
    return xo, o2, g


def depth_to_disp(depth, fxb: float):
    d = _f32c(depth, "depth")
    out = torch.empty_like(d)
    _lib.check(_lib.lib().d3_depth_to_disp(d.data_ptr(), fxb, d.numel(), out.data_ptr(), _stream()))
    return out


def disp_to_depth(disp, fxb: float):
    d = _f32c(disp, "disp")
    out = torch.empty_like(d)
    _lib.check(_lib.lib().d3_disp_to_depth(d.data_ptr(), fxb, d.numel(), out.data_ptr(), _stream()))
    return out


def normalize_disp(disp, num_chs: int, bound: float, gamma: float, t: float, s: float):
    """disp [..., H, W] with leading dims flattened to N -> [N*num_chs, H, W] (caller reshapes)."""
    d = _f32c(disp, "disp")
    H, W = d.shape[-2:]
    N = d.numel() // (H * W)
    out = torch.empty((N * num_chs, H, W), dtype=torch.float32, device=d.device)
    _lib.check(_lib.lib().d3_normalize_disp(d.data_ptr(), N, H * W, num_chs, bound, gamma, t, s, out.data_ptr(),
                                            _stream()))
    return out


def denormalize_disp(y, num_chs: int, bound: float, gamma: float, t: float, s: float):
    """y [B, R*num_chs, H, W] -> [B, R, H, W]."""
    y = _f32c(y, "y")
    B, Cc, H, W = y.shape
    R = Cc // num_chs
    out = torch.empty((B, R, H, W), dtype=torch.float32, device=y.device)
    _lib.check(_lib.lib().d3_denormalize_disp(y.data_ptr(), B * R, H * W, num_chs, bound, gamma, t, s,
                                              out.data_ptr(), _stream()))
    return out


def _f32_array(vals):
    arr = (C.c_float * len(vals))(*[float(v) for v in vals])
    return arr


def normalize_piecewise(disp, bounds, gammas, t: float, s: float):
    """disp [..., H, W] -> [N * len(bounds), H, W] (``d3_normalize_piecewise``)."""
    d = _f32c(disp, "disp")
    H, W = d.shape[-2:]
    N = d.numel() // (H * W)
    n = len(bounds)
    out = torch.empty((N * n, H, W), dtype=torch.float32, device=d.device)
    _lib.check(_lib.lib().d3_normalize_piecewise(d.data_ptr(), N, H * W, n, _f32_array(bounds), _f32_array(gammas), t, s,
                                                 out.data_ptr(), _stream()))
    return out


def denormalize_piecewise(y, bounds, gammas, t: float, s: float):
    """y [B, R*n, H, W] -> [B, R, H, W] (``d3_denormalize_piecewise``)."""
    y = _f32c(y, "y")
    B, Cc, H, W = y.shape
    n = len(bounds)
    R = Cc // n
    out = torch.empty((B, R, H, W), dtype=torch.float32, device=y.device)
    _lib.check(_lib.lib().d3_denormalize_piecewise(y.data_ptr(), B * R, H * W, n, _f32_array(bounds), _f32_array(gammas),
                                                   t, s, out.data_ptr(), _stream()))
    return out


def pad_replicate(x, pads):
    """x [B,C,H,W] fp32, pads = [left, right, top, bottom] (InputPadder._pad order)."""
    x = _f32c(x, "x")
    B, Cc, H, W = x.shape
    l, r, t, b = pads
    out = torch.empty((B, Cc, H + t + b, W + l + r), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_pad_replicate(x.data_ptr(), B * Cc, H, W, l, r, t, b, out.data_ptr(), _stream()))
    return out


def pad_zero(x, pads):
    """x [B,C,H,W] fp32, zero border, pads = [left, right, top, bottom] (InputPadder.pad_zero)."""
    x = _f32c(x, "x")
    B, Cc, H, W = x.shape
    l, r, t, b = pads
    out = torch.empty((B, Cc, H + t + b, W + l + r), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_pad_zero(x.data_ptr(), B * Cc, H, W, l, r, t, b, out.data_ptr(), _stream()))
    return out


def clamp_target_(x, mask, clip_range: float):
    """In place: x <- clamp(x, -r, r), mask <- mask * (|x| < r)  (train.py:175-178); x [B,C,H,W], mask [B,1,H,W]."""
    B, Cc, H, W = x.shape
    if not (x.is_contiguous() and mask.is_contiguous() and x.dtype == mask.dtype == torch.float32 and x.is_cuda):
        raise ValueError("clamp_target_: contiguous fp32 CUDA tensors expected")
    _lib.check(_lib.lib().d3_clamp_target(x.data_ptr(), mask.data_ptr(), B, Cc, H * W, clip_range, _stream()))
    return x, mask


def crop(x, top: int, left: int, H: int, W: int):
    x = _f32c(x, "x")
    B, Cc, Hi, Wi = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_crop(x.data_ptr(), B * Cc, Hi, Wi, top, left, H, W, out.data_ptr(), _stream()))
    return out


def clamp_into(x, out, slot: int, lo: float = -1.0, hi: float = 1.0):
    """out[:, slot*C:(slot+1)*C] = clamp(x, lo, hi) for x [B,C,H,W], out [B,C*T,H,W]."""
    x = _f32c(x, "x")
    B, Cc, H, W = x.shape
    chw = Cc * H * W
    _lib.check(_lib.lib().d3_clamp_into(x.data_ptr(), B, chw, out.stride(0), slot * chw, lo, hi, out.data_ptr(),
                                        _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# evaluation post-processing
# ------------------------------------------------------------------------------------------------
def disp_metrics(gt_disp, pred_disp, mask, fxb=None) -> torch.Tensor:
    """[B,H,W] fp32 disparities (pixels) + mask -> [B, 11] fp64: epe, d1, d2, d3, d5, a1, a2, a3, rmse, rel, mae."""
    g, p = _f32c(gt_disp, "gt_disp"), _f32c(pred_disp, "pred_disp")
    m = _f32c(mask.to(torch.float32), "mask")
    B = g.shape[0]
    HW = g.numel() // max(B, 1)
    f = None
    if fxb is not None:
        f = torch.as_tensor(fxb, dtype=torch.float32, device=g.device).reshape(-1).contiguous()
    out = torch.empty((B, 11), dtype=torch.float64, device=g.device)
    _lib.check(_lib.lib().d3_disp_metrics(g.data_ptr(), p.data_ptr(), m.data_ptr(), _ptr(f), 0 if f is None else f.numel(),
                                          B, HW, out.data_ptr(), _stream()))
    return out


def scale_shift_lsq(pred, target, mask) -> torch.Tensor:
    """[R, HW] rows -> [R, 2] (scale, shift) of the masked least-squares fit pred * s + t ~ target."""
    p, t = _f32c(pred, "pred"), _f32c(target, "target")
    m = _f32c(mask.to(torch.float32), "mask")
    R = p.shape[0]
    out = torch.empty((R, 2), dtype=torch.float32, device=p.device)
    _lib.check(_lib.lib().d3_scale_shift_lsq(p.data_ptr(), t.data_ptr(), m.data_ptr(), R, p.numel() // max(R, 1),
                                             out.data_ptr(), _stream()))
    return out


def masked_minmax(x, mask=None) -> torch.Tensor:
    x = _f32c(x, "x")
    m = None if mask is None else _f32c(mask.to(torch.float32), "mask")
    R = x.shape[0]
    out = torch.empty((R, 2), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().d3_masked_minmax(x.data_ptr(), _ptr(m), R, x.numel() // max(R, 1), out.data_ptr(), _stream()))
    return out


def normalize_ssi(x, mask, low_up, t: float, s: float) -> torch.Tensor:
    x = _f32c(x, "x")
    m = None if mask is None else _f32c(mask.to(torch.float32), "mask")
    lu = _f32c(low_up, "low_up")
    R = x.shape[0]
    out = torch.empty_like(x)
    _lib.check(_lib.lib().d3_normalize_ssi(x.data_ptr(), _ptr(m), lu.data_ptr(), R, x.numel() // max(R, 1), t, s,
                                           out.data_ptr(), _stream()))
    return out


def affine_rows(y, scale_shift) -> torch.Tensor:
    y = _f32c(y, "y")
    st = _f32c(scale_shift, "scale_shift")
    R = st.shape[0]
    out = torch.empty_like(y)
    _lib.check(_lib.lib().d3_affine_rows(y.data_ptr(), st.data_ptr(), R, y.numel() // max(R, 1), out.data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# training step: backward / optimizer ops of csrc/conv_wgrad.cu, train_ops.cu, attention_bwd.cu
# ------------------------------------------------------------------------------------------------
def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, *, ksize: int, stride: int = 1,
                 col0: int = 0) -> torch.Tensor:
    """dw[co, col0 + tap*Cin + ci] += sum dy[b,y,x,co] * x[b, s*y+kh-1, s*x+kw-1, ci] (``d3_conv2d_wgrad``); dw fp32
    [Cout, Ktot] in the forward pack layout (a row-pitched view is allowed)."""
    _check_nhwc(x, "x")
    _check_nhwc(dy, "dy")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    if dy.shape[:3] != (B, H // stride, W // stride) or dy.dtype != x.dtype:
        raise ValueError(f"conv2d_wgrad: dy {tuple(dy.shape)} does not match x {tuple(x.shape)} at stride {stride}")
    if dw.dtype != torch.float32 or dw.dim() != 2 or dw.shape[0] != Cout or dw.stride(1) != 1:
        raise ValueError("conv2d_wgrad: dw must be fp32 [Cout, Ktot] with contiguous rows")
    p = _lib.WgradParams()
    p.x, p.B, p.H, p.W, p.Cin, p.x_pitch = x.data_ptr(), B, H, W, Cin, x.stride(2)
    p.dy, p.Cout, p.dy_pitch = dy.data_ptr(), Cout, dy.stride(2)
    p.ksize, p.stride = ksize, stride
    p.dw, p.dw_pitch, p.dw_col0 = dw.data_ptr(), dw.stride(0), col0
    p.dtype = _dt(x)
    _lib.check(_lib.lib().d3_conv2d_wgrad(C.byref(p), _stream()))
    return dw


def groupnorm_silu_bwd(x1: torch.Tensor, stats1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dy: torch.Tensor, *,
                       x2: Optional[torch.Tensor] = None, stats2: Optional[torch.Tensor] = None, silu: bool = True,
                       groups: int = 32, eps: float = 1e-5, addends=(), dgamma: Optional[torch.Tensor] = None,
                       dbeta: Optional[torch.Tensor] = None, colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Backward of ``groupnorm_silu``: returns dx over the concatenated channels (+ up to 3 ``addends``, dense NHWC
    tensors of dx's shape); ``dgamma`` / ``dbeta`` (fp32 [C]) are accumulated; ``colsum`` (fp32 [B, Cs] view with
    contiguous channels, or [Cs] = summed over the batch) receives the per-channel sums of dx over its first Cs channels."""
    _check_nhwc(x1, "x1")
    _check_nhwc(dy, "dy")
    B, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[3]
    Ct = C1 + C2
    if dy.shape != (B, H, W, Ct):
        raise ValueError("groupnorm_silu_bwd: dy must cover the concatenated channels")
    addends = [t for t in addends if t is not None]
    if len(addends) > 3:
        raise ValueError("groupnorm_silu_bwd: at most 3 addends")
    for t in addends:
        _check_nhwc(t, "addend")
        if t.shape != (B, H, W, Ct) or t.dtype != x1.dtype:
            raise ValueError("groupnorm_silu_bwd: addends must have dx's shape and dtype")
    dx = torch.empty((B, H, W, Ct), dtype=x1.dtype, device=x1.device)
    ws = torch.empty((B * Ct * 2,), dtype=torch.float32, device=x1.device)
    cs_ptr, cs_stride, cs_ch = None, 0, 0
    if colsum is not None:
        if colsum.dtype != torch.float32 or colsum.stride(-1) != 1 or colsum.shape[-1] > Ct or colsum.shape[-1] % 8:
            raise ValueError("groupnorm_silu_bwd: colsum must be fp32 [..., Cs <= C] with contiguous channels")
        cs_ptr, cs_stride, cs_ch = colsum.data_ptr(), (colsum.stride(0) if colsum.dim() == 2 else 0), colsum.shape[-1]
    n = len(addends)
    ptrs = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in addends])
    pitches = (C.c_int * max(n, 1))(*[t.stride(2) for t in addends])
    _lib.check(_lib.lib().d3_groupnorm_silu_bwd(
        x1.data_ptr(), C1, x1.stride(2), stats1.data_ptr(), _ptr(x2), C2, x2.stride(2) if x2 is not None else 0,
        _ptr(stats2), gamma.data_ptr(), beta.data_ptr(), B, H * W, groups, eps, int(silu), dy.data_ptr(), dy.stride(2),
        ptrs, pitches, n, dx.data_ptr(), dx.stride(2), _ptr(dgamma), _ptr(dbeta), cs_ptr, cs_stride, cs_ch,
        ws.data_ptr(), _dt(x1), _stream()))
    return dx


def colsum_nhwc16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (fp32 [B, C] view or [C]) += per-channel sums of x over the pixels (and over the batch for a 1-D out)."""
    _check_nhwc(x, "x")
    B, H, W, Cc = x.shape
    if out.dtype != torch.float32 or out.shape[-1] != Cc or out.stride(-1) != 1:
        raise ValueError("colsum_nhwc16: out must be fp32 [..., C] with contiguous channels")
    _lib.check(_lib.lib().d3_colsum_nhwc16(x.data_ptr(), x.stride(2), B, H * W, Cc, out.data_ptr(),
                                           out.stride(0) if out.dim() == 2 else 0, _dt(x), _stream()))
    return out


def add_nhwc16(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
               colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a + b; ``colsum`` (fp32 [B, C] view or [C]) additionally receives the per-channel sums of the result."""
    _check_nhwc(a, "a")
    _check_nhwc(b, "b")
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("add_nhwc16: shape / dtype mismatch")
    B, H, W, Cc = a.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=a.dtype, device=a.device)
    if colsum is not None:
        if colsum.dtype != torch.float32 or colsum.shape[-1] != Cc or colsum.stride(-1) != 1:
            raise ValueError("add_nhwc16: colsum must be fp32 [..., C] with contiguous channels")
        _lib.check(_lib.lib().d3_add_colsum_nhwc16(a.data_ptr(), a.stride(2), b.data_ptr(), b.stride(2), out.data_ptr(),
                                                   out.stride(2), B, H * W, Cc, colsum.data_ptr(),
                                                   colsum.stride(0) if colsum.dim() == 2 else 0, _dt(a), _stream()))
        return out
    _lib.check(_lib.lib().d3_add_nhwc16(a.data_ptr(), a.stride(2), b.data_ptr(), b.stride(2), out.data_ptr(),
                                        out.stride(2), B * H * W, Cc, _dt(a), _stream()))
    return out


def sumpool2x_nhwc16(x: torch.Tensor, addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, 2H, 2W, C] -> [B, H, W, C] block sums (+ addend): backward of ``upsample_nearest2x``."""
    _check_nhwc(x, "x")
    B, H2, W2, Cc = x.shape
    out = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().d3_sumpool2x_nhwc16(x.data_ptr(), x.stride(2), _ptr(addend),
                                              addend.stride(2) if addend is not None else 0, out.data_ptr(),
                                              out.stride(2), B, H2 // 2, W2 // 2, Cc, _dt(x), _stream()))
    return out


def conv_out_bwd(h: torch.Tensor, stats: torch.Tensor, gamma, beta, w, dy: torch.Tensor, dw: torch.Tensor,
                 eps: float = 1e-5) -> torch.Tensor:
    """Backward of ``conv_out_fused`` up to SiLU(GroupNorm(h)): returns da (16-bit NHWC [B,H,W,128]); dw fp32 [dc,9,128]
    is accumulated."""
    _check_nhwc(h, "h")
    B, H, W, Cc = h.shape
    dc = dy.shape[1]
    dy = _f32c(dy, "dy")
    da = torch.empty((B, H, W, 128), dtype=h.dtype, device=h.device)
    _lib.check(_lib.lib().d3_conv_out_bwd(h.data_ptr(), h.stride(2), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                          w.data_ptr(), dy.data_ptr(), B, H, W, dc, eps, da.data_ptr(), da.stride(2),
                                          dw.data_ptr(), _dt(h), _stream()))
    return da


def attention_hd8_bwd(qkv: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor) -> torch.Tensor:
    _check_nhwc(qkv, "qkv")
    B, H, W, C3 = qkv.shape
    C_ = C3 // 3
    dqkv = torch.empty((B, H, W, C3), dtype=qkv.dtype, device=qkv.device)
    ws = torch.empty((_lib.lib().d3_attention_hd8_bwd_workspace_bytes(B, H * W, C_),), dtype=torch.uint8, device=qkv.device)
    _lib.check(_lib.lib().d3_attention_hd8_bwd(qkv.data_ptr(), qkv.stride(2), o.data_ptr(), o.stride(2), d_o.data_ptr(),
                                               d_o.stride(2), B, H * W, C_, dqkv.data_ptr(), dqkv.stride(2),
                                               ws.data_ptr(), _dt(qkv), _stream()))
    return dqkv


def masked_loss_grad(pred: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor],
                     weights: Optional[torch.Tensor], *, l1: bool = False, grad_scale: float = 1.0,
                     dbias: Optional[torch.Tensor] = None):
    """Returns (dy fp32 like pred, loss_sum fp64 [B]): see ``d3_masked_loss_grad``."""
    pred, target = _f32c(pred, "pred"), _f32c(target, "target")
    B, dc, H, W = pred.shape
    dy = torch.empty_like(pred)
    loss_sum = torch.zeros((B,), dtype=torch.float64, device=pred.device)
    _lib.check(_lib.lib().d3_masked_loss_grad(pred.data_ptr(), target.data_ptr(), _ptr(mask), _ptr(weights), B, dc, H * W,
                                              int(l1), grad_scale, dy.data_ptr(), loss_sum.data_ptr(), _ptr(dbias),
                                              _stream()))
    return dy, loss_sum


def linear_small_bwd(x: torch.Tensor, w: Optional[torch.Tensor], dy: torch.Tensor, *, act_in: int = 0,
                     dw: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None, want_dx: bool = True):
    """Backward of ``linear_small`` (y = act_in(x) W^T + b): accumulates dw / db, returns dx (or None)."""
    n, K = x.shape
    rows = dy.shape[1]
    dx = torch.empty_like(x) if want_dx else None
    _lib.check(_lib.lib().d3_linear_small_bwd(x.data_ptr(), n, K, act_in, _ptr(w), rows, dy.data_ptr(), _ptr(dw), _ptr(db),
                                              _ptr(dx), 0, _stream()))
    return dx


def bilinear_upsample_acc_(dst: torch.Tensor, src: torch.Tensor, mul: float) -> torch.Tensor:
    """dst [B,C,H,W] += mul * Upsample(size=(H, W), mode="bilinear")(src [B,C,h,w]), fp32, in place."""
    dst_, src = _f32c(dst, "dst"), _f32c(src, "src")
    B, Cc, H, W = dst.shape
    _lib.check(_lib.lib().d3_bilinear_upsample_acc(src.data_ptr(), B * Cc, src.shape[2], src.shape[3], dst_.data_ptr(), H, W,
                                                   mul, _stream()))
    return dst


def div_by_std_(x: torch.Tensor) -> torch.Tensor:
    """x /= x.std() (unbiased), in place."""
    x = _f32c(x, "x")
    ws = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().d3_div_by_std(x.data_ptr(), x.numel(), ws.data_ptr(), _stream()))
    return x


def sumsq(g: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _lib.check(_lib.lib().d3_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), _stream()))
    return out


def adamw_step(p, g, m, v, *, p16=None, n16: int = 0, lr: float, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-8, weight_decay: float = 0.01, step: int, sumsq_buf=None, inv_loss_scale: float = 1.0,
               max_norm: float = 0.0, norm_out=None) -> None:
    _lib.check(_lib.lib().d3_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), _ptr(p16), n16,
                                        _dt(p16) if p16 is not None else 0, lr, beta1, beta2, eps, weight_decay, step,
                                        _ptr(sumsq_buf), inv_loss_scale, max_norm, _ptr(norm_out), _stream()))


def cast_f32_to_16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _lib.check(_lib.lib().d3_cast_f32_to_16(x.data_ptr(), out.data_ptr(), x.numel(), _dt(out), _stream()))
    return out


def gather_add(src: torch.Tensor, ia: torch.Tensor, ib: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _lib.check(_lib.lib().d3_gather_add(src.data_ptr(), ia.data_ptr(), ib.data_ptr(), out.data_ptr(), out.numel(), _stream()))
    return out


def pack_taps16(w: torch.Tensor, col0: int, Cin: int, tap_map, out: torch.Tensor) -> torch.Tensor:
    """out[ci, j, co] (16-bit) = w[co, col0 + tap_map[j]*Cin + ci] (0 where tap_map[j] < 0); w fp32 [Cout, Ktot]."""
    Cout = w.shape[0]
    tm = (C.c_int32 * len(tap_map))(*tap_map)
    _lib.check(_lib.lib().d3_pack_taps16(w.data_ptr(), w.stride(0), col0, Cin, Cout, tm, len(tap_map), out.data_ptr(),
                                         _dt(out), _stream()))
    return out


def make_pack_table(entries, device):
    """Device-side table for ``pack_taps16_batched``: ``entries`` = [(w_offset, w_pitch, out_offset, col0, Cin, Cout,
    tap_map)] (offsets in elements).  Returns (entries tensor (bytes), tile-prefix tensor, n_entries, total_tiles)."""
    import numpy as np
    arr = (_lib.PackEntry * len(entries))()
    prefix, total = [], 0
    for i, (wo, wp, oo, col0, cin, cout, tm) in enumerate(entries):
        if cin % 32 or cout % 32 or not 1 <= len(tm) <= 16:
            raise ValueError("make_pack_table: Cin / Cout must be multiples of 32, 1..16 output taps")
        e = arr[i]
        e.w_offset, e.w_pitch, e.out_offset = wo, wp, oo
        e.col0, e.Cin, e.Cout, e.n_out_taps = col0, cin, cout, len(tm)
        for j, t in enumerate(tm):
            e.tap_map[j] = t
        prefix.append(total)
        total += (cout // 32) * (cin // 32) * len(tm)
    raw = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).to(device)
    pre = torch.tensor(prefix, dtype=torch.int32, device=device)
    return raw, pre, len(entries), total


def pack_taps16_batched(table, w_base: torch.Tensor, out_base: torch.Tensor) -> torch.Tensor:
    raw, pre, n, total = table
    _lib.check(_lib.lib().d3_pack_taps16_batched(raw.data_ptr(), pre.data_ptr(), n, total, w_base.data_ptr(),
                                                 out_base.data_ptr(), _dt(out_base), _stream()))
    return out_base

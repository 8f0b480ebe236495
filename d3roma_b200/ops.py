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


def _check_nhwc(t: torch.Tensor, name: str) -> None:
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4 and t.stride(3) == 1):
        raise ValueError(f"{name}: expected a CUDA bf16 NHWC tensor with contiguous channels")
    B, H, W, _ = t.shape
    p = t.stride(2)
    if t.stride(1) != W * p or (B > 1 and t.stride(0) != H * W * p):
        raise ValueError(f"{name}: expected dense pixels (strides {t.stride()})")


def pack_conv_weight(w_oihw: torch.Tensor, shortcut_w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """OIHW fp32/bf16 conv weight (+ optional [Cout, Csc(,1,1)] 1x1 shortcut weight) -> bf16 [Cout, Ktot]
    with k = (kh*KW + kw)*Cin + c, shortcut channels appended (layout the igemm kernel expects)."""
    co, ci, kh, kw = w_oihw.shape
    w = w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
    if shortcut_w is not None:
        w = torch.cat([w, shortcut_w.reshape(co, -1).to(w.dtype)], dim=1)
    return w.to(torch.bfloat16).contiguous()


def conv2d_igemm(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, *, ksize: int, stride: int = 1,
                 sc1: Optional[torch.Tensor] = None, sc2: Optional[torch.Tensor] = None,
                 residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tcgen05 implicit-GEMM conv (see ``d3_conv2d_igemm`` in include/d3roma_b200.h)."""
    _check_nhwc(x, "x")
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    Ho, Wo = H // stride, W // stride
    if out is None:
        out = torch.empty((B, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
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
    if w_packed.shape[1] != ksize * ksize * Cin + c1 + c2 or w_packed.dtype != torch.bfloat16:
        raise ValueError("w_packed shape/dtype mismatch")
    if bias.dtype != torch.float32 or bias.shape[-1] != Cout:
        raise ValueError("bias must be fp32 [Cout] or [B, Cout]")
    p.bias, p.bias_bstride = bias.data_ptr(), (Cout if bias.dim() == 2 else 0)
    if residual is not None:
        _check_nhwc(residual, "residual")
        p.residual, p.res_pitch = residual.data_ptr(), residual.stride(2)
    p.out, p.out_pitch = out.data_ptr(), out.stride(2)
    if stats is not None:
        if stats.dtype != torch.float64 or stats.numel() != B * (Cout // 4) * 2:
            raise ValueError("stats must be fp64 [B, Cout/4, 2]")
        p.stats = stats.data_ptr()
    _lib.check(_lib.lib().d3_conv2d_igemm(C.byref(p), _stream()))
    return out

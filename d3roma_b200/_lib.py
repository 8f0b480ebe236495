"""ctypes binding of ``libd3roma_b200.so`` (the C ABI declared in ``include/d3roma_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception
is raised.  The product path never routes through ``oracle/`` or a PyTorch re-implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libd3roma_b200.so"


class D3Error(RuntimeError):
    pass


DTYPE_BF16 = 0
DTYPE_F16 = 1
DTYPE_F32 = 2


class ConvParams(C.Structure):
    """Mirror of ``d3_conv_params`` (include/d3roma_b200.h)."""

    _fields_ = [
        ("x", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("x_pitch", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("w", C.c_void_p),
        ("Cout", C.c_int32),
        ("sc1", C.c_void_p), ("C1", C.c_int32), ("sc1_pitch", C.c_int32),
        ("sc2", C.c_void_p), ("C2", C.c_int32), ("sc2_pitch", C.c_int32),
        ("bias", C.c_void_p), ("bias_bstride", C.c_int32),
        ("residual", C.c_void_p), ("res_pitch", C.c_int32),
        ("out", C.c_void_p), ("out_pitch", C.c_int32),
        ("stats", C.c_void_p),
        ("dtype", C.c_int32),
        ("up2x_phase", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("Cout_valid", C.c_int32), ("asym_pad", C.c_int32),
    ]


class WgradParams(C.Structure):
    """Mirror of ``d3_wgrad_params``."""

    _fields_ = [
        ("x", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("x_pitch", C.c_int32),
        ("dy", C.c_void_p),
        ("Cout", C.c_int32), ("dy_pitch", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("dw", C.c_void_p),
        ("dw_pitch", C.c_int64),
        ("dw_col0", C.c_int32), ("dtype", C.c_int32),
    ]


class PackEntry(C.Structure):
    """Mirror of ``d3_pack_entry``."""

    _fields_ = [
        ("w_offset", C.c_int64), ("w_pitch", C.c_int64), ("out_offset", C.c_int64),
        ("col0", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32), ("n_out_taps", C.c_int32),
        ("tap_map", C.c_int32 * 16),
    ]


class DDIMCoeffs(C.Structure):
    """Mirror of ``d3_ddim_coeffs``."""

    _fields_ = [
        ("prediction_type", C.c_int32),
        ("sqrt_alpha_t", C.c_float), ("sqrt_beta_t", C.c_float), ("sqrt_alpha_prev", C.c_float),
        ("dir_coeff", C.c_float), ("std_dev", C.c_float),
        ("clip_mode", C.c_int32), ("clip_range", C.c_float),
        ("use_clipped_model_output", C.c_int32),
    ]


class DDPMCoeffs(C.Structure):
    """Mirror of ``d3_ddpm_coeffs``."""

    _fields_ = [
        ("prediction_type", C.c_int32),
        ("sqrt_alpha_t", C.c_float), ("sqrt_beta_t", C.c_float),
        ("x0_coeff", C.c_float), ("xt_coeff", C.c_float), ("sigma", C.c_float),
        ("clip_mode", C.c_int32), ("clip_range", C.c_float),
    ]


class GuidanceParams(C.Structure):
    """Mirror of ``d3_guidance_params``."""

    _fields_ = [
        ("x0", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p),
        ("grad", C.c_void_p), ("x0_out", C.c_void_p), ("out2", C.c_void_p),
        ("aux1", C.c_void_p), ("aux2", C.c_void_p),
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("margin", C.c_int32),
        ("disp_scale", C.c_float), ("disp_offset", C.c_float), ("min_disp", C.c_float), ("max_disp", C.c_float),
        ("w_ssim", C.c_float),
        ("step", C.c_float), ("ca", C.c_float), ("cb", C.c_float), ("cc", C.c_float),
    ]


_lib = None


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    sigs = {
        "d3_last_error": (C.c_char_p, []),
        "d3_version": (i32, []),
        "d3_launch_count": (i64, []),
        "d3_reset_launch_count": (None, []),
        "d3_set_pdl": (i32, [i32]),
        "d3_conv2d_igemm": (i32, [C.POINTER(ConvParams), vp]),
        "d3_conv2d_workspace_bytes": (C.c_size_t, [C.POINTER(ConvParams)]),
        "d3_groupnorm_silu": (i32, [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, f32, i32, vp, i32, i32,
                                    vp]),
        "d3_groupnorm_silu_dual": (i32, [vp, i32, i32, vp, i32, i32, f32, vp, vp, i32, i32, vp, i32, vp, vp, i32, i32,
                                         vp, i32, i32, vp]),
        "d3_upsample_nearest2x": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
        "d3_im2col_input": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, i32, vp]),
        "d3_conv_in_direct": (i32, [vp, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, i32, vp, vp]),
        "d3_conv_out_fused": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, i32, vp]),
        "d3_conv_out_fused_ddim": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, i32, vp,
                                         C.POINTER(DDIMCoeffs), vp, vp, vp]),
        "d3_conv_out_workspace_bytes": (C.c_size_t, [i32, i32, i32, i32]),
        "d3_time_embed": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
        "d3_temb_project_all": (i32, [vp, i32, vp, vp, i32, vp, vp]),
        "d3_attention_hd8": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, vp]),
        "d3_sched_ddim_step": (i32, [vp, vp, vp, C.POINTER(DDIMCoeffs), i64, vp, vp, vp, vp, vp]),
        "d3_sched_ddpm_step": (i32, [vp, vp, vp, C.POINTER(DDPMCoeffs), i64, vp, vp, vp, vp]),
        "d3_lincomb": (i32, [f32, vp, f32, vp, f32, vp, i64, vp, vp]),
        "d3_guidance_grad": (i32, [vp, vp, vp, i32, i32, i32, i32, f32, f32, f32, f32, vp, vp]),
        "d3_guidance_step": (i32, [C.POINTER(GuidanceParams), vp]),
        "d3_depth_to_rgb_frame_workspace_bytes": (C.c_size_t, [i32, i32]),
        "d3_depth_to_rgb_frame": (i32, [vp, i32, i32, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), i32, i32, i32, i32,
                                        i32, i32, i32, vp, vp, vp]),
        "d3_depth_to_disp": (i32, [vp, f32, i64, vp, vp]),
        "d3_disp_to_depth": (i32, [vp, f32, i64, vp, vp]),
        "d3_normalize_disp": (i32, [vp, i64, i64, i32, f32, f32, f32, f32, vp, vp]),
        "d3_denormalize_disp": (i32, [vp, i64, i64, i32, f32, f32, f32, f32, vp, vp]),
        "d3_normalize_piecewise": (i32, [vp, i64, i64, i32, C.POINTER(f32), C.POINTER(f32), f32, f32, vp, vp]),
        "d3_denormalize_piecewise": (i32, [vp, i64, i64, i32, C.POINTER(f32), C.POINTER(f32), f32, f32, vp, vp]),
        "d3_pad_replicate": (i32, [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp]),
        "d3_crop": (i32, [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp]),
        "d3_pad_zero": (i32, [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp]),
        "d3_clamp_target": (i32, [vp, vp, i64, i32, i64, f32, vp]),
        "d3_clamp_into": (i32, [vp, i64, i64, i64, i64, f32, f32, vp, vp]),
        "d3_disp_metrics": (i32, [vp, vp, vp, vp, i32, i32, i64, vp, vp]),
        "d3_scale_shift_lsq": (i32, [vp, vp, vp, i32, i64, vp, vp]),
        "d3_masked_minmax": (i32, [vp, vp, i32, i64, vp, vp]),
        "d3_normalize_ssi": (i32, [vp, vp, vp, i32, i64, f32, f32, vp, vp]),
        "d3_affine_rows": (i32, [vp, vp, i32, i64, vp, vp]),
        "d3_ransac_workspace_bytes": (C.c_size_t, [i32, i64]),
        "d3_ransac_iteration": (i32, [vp, vp, vp, vp, i32, i32, i64, f32, f32, vp, vp, vp, vp, vp]),
        "d3_frame_to_input": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
        "d3_depth_to_input": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, f32, f32, f32, f32, f32, vp, vp,
                                    vp, vp]),
        "d3_gn_stats": (i32, [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, vp]),
        "d3_gn_apply": (i32, [vp, i32, i32, vp, i32, i32, vp, vp, vp, i32, i32, i32, f32, i32, vp, i32, i32, vp]),
        "d3_groupnorm_one_pass": (i32, [vp, i32, i32, vp, i32, i32, vp, vp, i32, i32, i32, f32, i32, vp, i32, i32, vp]),
        "d3_layernorm": (i32, [vp, i32, vp, vp, i64, i32, f32, vp, i32, i32, vp]),
        "d3_geglu": (i32, [vp, i32, i64, i32, vp, i32, i32, vp]),
        "d3_attention_hd64": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp]),
        "d3_attention_hd64_tc": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
        "d3_attention_hd64_tc_workspace_bytes": (C.c_size_t, [i32, i32, i32]),
        "d3_cross_attention_small": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, vp, i32, i32, vp]),
        "d3_linear_small": (i32, [vp, i32, i32, vp, vp, i32, i32, i32, vp, i32, vp]),
        "d3_timestep_sincos": (i32, [vp, i32, i32, i32, f32, vp, vp]),
        "d3_peer_alloc": (i32, [C.c_size_t, C.POINTER(vp), vp]),
        "d3_peer_open": (i32, [vp, C.POINTER(vp)]),
        "d3_peer_close": (i32, [vp]),
        "d3_peer_free": (i32, [vp]),
        "d3_denormalize_gather": (i32, [vp, i64, i64, i32, f32, f32, f32, f32, C.POINTER(vp), i32, i64, vp]),
        "d3_peer_push": (i32, [vp, C.c_size_t, C.POINTER(vp), i32, C.c_size_t, vp]),
        "d3_peer_signal": (i32, [C.POINTER(vp), i32, i32, C.c_uint32, vp]),
        "d3_peer_wait": (i32, [vp, i32, C.c_uint32, vp]),
        "d3_ensemble_workspace_bytes": (C.c_size_t, []),
        "d3_ensemble_objective": (i32, [vp, i32, i64, vp, vp, i32, vp, vp, vp]),
        "d3_ensemble_reduce": (i32, [vp, i32, i64, vp, vp, i32, vp, vp, vp, vp]),
        "d3_embed_tokens": (i32, [vp, i32, i32, vp, i32, vp, i32, i32, vp, vp]),
        "d3_layernorm_f32": (i32, [vp, i64, i32, vp, vp, f32, vp, vp]),
        "d3_causal_attention_f32": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, i32, vp]),
        "d3_nchw_to_nhwc16": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
        "d3_nhwc16_to_nchw": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
        "d3_resize_nearest_nhwc16": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, i32, i32, vp]),
        "d3_softmax_rows": (i32, [vp, i32, i32, i32, f32, i32, vp]),
        "d3_transpose_pad16": (i32, [vp, i32, i32, i32, vp, i32, vp]),
        "d3_conv1x1_nchw_small": (i32, [vp, i32, i32, i32, i64, vp, vp, vp, vp]),
        # training step
        "d3_conv2d_wgrad": (i32, [C.POINTER(WgradParams), vp]),
        "d3_groupnorm_bwd_workspace_bytes": (C.c_size_t, [i32, i32]),
        "d3_groupnorm_silu_bwd": (i32, [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, f32, i32, vp, i32,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_int), i32, vp, i32, vp, vp, vp, i64, i32, vp,
                                        i32, vp]),
        "d3_colsum_nhwc16": (i32, [vp, i32, i32, i32, i32, vp, i64, i32, vp]),
        "d3_add_nhwc16": (i32, [vp, i32, vp, i32, vp, i32, i64, i32, i32, vp]),
        "d3_add_colsum_nhwc16": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i64, i32, vp]),
        "d3_sumpool2x_nhwc16": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
        "d3_conv_out_bwd": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, i32, vp, i32, vp]),
        "d3_attention_hd8_bwd_workspace_bytes": (C.c_size_t, [i32, i32, i32]),
        "d3_attention_hd8_bwd": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
        "d3_masked_loss_grad": (i32, [vp, vp, vp, vp, i32, i32, i64, i32, f32, vp, vp, vp, vp]),
        "d3_linear_small_bwd": (i32, [vp, i32, i32, i32, vp, i32, vp, vp, vp, vp, i32, vp]),
        "d3_bilinear_upsample_acc": (i32, [vp, i64, i32, i32, vp, i32, i32, f32, vp]),
        "d3_div_by_std": (i32, [vp, i64, vp, vp]),
        "d3_sumsq": (i32, [vp, i64, vp, vp]),
        "d3_adamw_step": (i32, [vp, vp, vp, vp, i64, vp, i64, i32, f32, f32, f32, f32, f32, i32, vp, f32, f32, vp, vp]),
        "d3_cast_f32_to_16": (i32, [vp, vp, i64, i32, vp]),
        "d3_gather_add": (i32, [vp, vp, vp, vp, i32, vp]),
        "d3_pack_taps16": (i32, [vp, i64, i32, i32, i32, C.POINTER(C.c_int32), i32, vp, i32, vp]),
        "d3_pack_taps16_batched": (i32, [vp, vp, i32, i32, vp, vp, i32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == header and library out of sync
        fn.restype = res
        fn.argtypes = args


def lib() -> C.CDLL:
    """Return the loaded library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise D3Error(
                f"{LIB_PATH} is missing: build it with `python -m d3roma_b200.build` "
                "(there is no CPU / PyTorch fallback for the CUDA path)")
        handle = C.CDLL(str(LIB_PATH))
        _declare(handle)      # raises when header and library are out of sync: never hand out a half-declared library
        _lib = handle         # (calls through undeclared ctypes prototypes would truncate 64-bit pointers)
    return _lib


_SYNC_CHECK = bool(int(os.environ.get("D3_SYNC_CHECK", "0")))   # debugging aid: synchronise after every library call


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().d3_last_error()
        raise D3Error(f"d3roma_b200 error {rc}: {msg.decode() if msg else '?'}")
    if _SYNC_CHECK:
        import torch
        torch.cuda.synchronize()

"""ORACLE (test infrastructure, not product code): CPU restatement of the pre/post-processing either side
of the denoising loop — ``Normalizer`` non-SSI "average" mode (/root/reference/utils_d3roma/utils.py:171-196,
231-273), ``InputPadder`` (:297-325), depth<->disparity (inference.py:128-131, 246-249) and
``normalize_disp/unnormalize_disp`` (utils_d3roma/camera.py:771-780).

Pinned against ``tests/golden/prepost.npz`` (reference outputs + the reference's own 13 known answers,
camera.py:1027-1044) in tests/test_oracle_prepost.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def normalize_average(x: torch.Tensor, bound: float, gamma: float = 1.0, num_chs: int = 1, t: float = 0.5,
                      s: float = 2.0) -> torch.Tensor:
    """utils.py:183-184,186-196: [1,H,W] -> [num_chs,H,W]; clamp(max=bound) / bound, ** gamma, replicate, (y-t)*s."""
    x = torch.clamp(x, max=torch.tensor(bound, dtype=torch.float32))
    y = torch.cat([(x / bound) ** gamma] * num_chs, dim=0)
    return (y - t) * s


def denormalize_average(y: torch.Tensor, bound: float, gamma: float = 1.0, num_chs: int = 1, t: float = 0.5,
                        s: float = 2.0) -> torch.Tensor:
    """utils.py:252-258,260-273: [B,R*num_chs,H,W] -> [B,R,H,W]; sum_c (y_c/s+t)^(1/gamma) * bound/num_chs."""
    B, C, H, W = y.shape
    R = C // num_chs
    v = (y.view(B * R, num_chs, H, W) / s + t)
    z = 0
    for i in range(num_chs):
        z = z + v[:, i:i + 1] ** (1 / gamma) * (bound / num_chs)
    return z.view(B, R, H, W)


def normalize_piecewise(x: torch.Tensor, bounds, gammas, t: float = 0.5, s: float = 2.0) -> torch.Tensor:
    """utils.py:186-189, 198-229: [1,H,W] -> [len(bounds),H,W]; channel c carries the part of the disparity inside
    [sum(bounds[:c]), sum(bounds[:c+1])), divided by its bound, ** gamma_c; then (y - t) * s."""
    x = torch.clamp(x, max=torch.sum(torch.tensor(bounds, dtype=torch.float32)))
    chans, res = [], x
    for b, g in zip(bounds, gammas):
        ch = torch.minimum(res, torch.ones_like(res) * b) / b
        res = torch.where(ch < 1.0, torch.zeros_like(res), res - b)
        chans.append(ch ** g)
    return (torch.cat(chans, dim=0) - t) * s


def denormalize_piecewise(y: torch.Tensor, bounds, gammas, t: float = 0.5, s: float = 2.0) -> torch.Tensor:
    """utils.py:252-258, 275-293: [B,R*n,H,W] -> [B,R,H,W]; sum_c (y_c/s+t)^(1/gamma_c) * bound_c."""
    B, C, H, W = y.shape
    n = len(bounds)
    R = C // n
    v = y.view(B * R, n, H, W) / s + t
    z = v[:, 0:1] ** (1 / gammas[0]) * bounds[0]
    for c in range(1, n):
        z = z + v[:, c:c + 1] ** (1 / gammas[c]) * bounds[c]
    return z.view(B, R, H, W)


def pad_amounts(h: int, w: int, divis_by: int = 32):
    """utils.py:299-304 ('sintel' mode): [left, right, top, bottom]."""
    ph = (((h // divis_by) + 1) * divis_by - h) % divis_by
    pw = (((w // divis_by) + 1) * divis_by - w) % divis_by
    return [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]


def pad_replicate(x: torch.Tensor, divis_by: int = 32) -> torch.Tensor:
    """utils.py:313-315."""
    return F.pad(x, pad_amounts(x.shape[-2], x.shape[-1], divis_by), mode="replicate")


def unpad(x: torch.Tensor, pads) -> torch.Tensor:
    """utils.py:321-325."""
    ht, wd = x.shape[-2:]
    return x[..., pads[2]:ht - pads[3], pads[0]:wd - pads[1]]


def depth_to_disp(depth: torch.Tensor, fxb: float) -> torch.Tensor:
    """inference.py:128-131: disp = fxb / depth where depth > 0 else 0."""
    out = torch.zeros_like(depth)
    v = depth > 0
    out[v] = fxb / depth[v]
    return out


def disp_to_depth(disp, fxb: float):
    """inference.py:246-249 (numpy on the way out)."""
    disp = np.asarray(disp)
    out = np.zeros_like(disp)
    v = disp > 0
    out[v] = fxb / disp[v]
    return out


def normalize_disp(disp, min_disp=5, max_disp=120, shift=0):
    """camera.py:771-776."""
    return (disp - min_disp) / (max_disp - min_disp) + shift


def unnormalize_disp(disp, min_disp=5, max_disp=120, shift=0):
    """camera.py:778-780."""
    return min_disp + (disp - shift) * (max_disp - min_disp)

"""ORACLE (test infrastructure, not product code): CPU restatement (numpy / torch) of the evaluation post-processing:
``compute_errors`` space="disp" (/root/reference/utils_d3roma/utils.py:554-653), ``compute_scale_and_shift``
(:488-528) and the ``Normalizer(ssi=True)`` branch (:171-182, 231-251).

Pinned against ``tests/golden/postproc.npz`` (outputs of the reference's own functions, oracle/gen_golden.py) in
tests/test_oracle_postproc.py.
"""
from __future__ import annotations

import numpy as np
import torch


def compute_errors_disp(gt_disps: np.ndarray, pred_disps: np.ndarray, masks=None, fxb=None):
    """utils.py:554-653 for space == "disp". Returns ([B,5] epe,d1,d2,d3,d5 ; [B,6] a1,a2,a3,rmse,rel,mae)."""
    valid = lambda x: ~(np.isnan(x) | np.isinf(x) | (x == 0))  # noqa: E731
    if masks is None:
        masks = valid(gt_disps)
    pred_disps = pred_disps.copy()
    pred_disps[~masks] = 0
    md, mz = [], []
    for b in range(gt_disps.shape[0]):
        g, p = gt_disps[b][masks[b]], pred_disps[b][masks[b]]
        if np.min(p) <= 0:
            p = np.clip(p, g.min(), None)                      # :596-601
        gz, pz = (fxb[b] / g, fxb[b] / p) if fxb is not None else (1 / g, 1 / p)
        e = np.abs(g - p)
        scale = g.shape[-1] / 480                              # :629 (sic: number of valid pixels / 480)
        md.append([e.mean().item()] + [np.mean(e > k * scale).item() for k in (1.0, 2.0, 3.0, 5.0)])
        th = np.maximum(gz / pz, pz / gz)
        mz.append([np.mean(th < 1.05).item(), np.mean(th < 1.10).item(), np.mean(th < 1.25).item(),
                   np.sqrt(np.mean((gz - pz) ** 2)).item(), np.mean(np.abs(gz - pz) / gz).item(),
                   np.mean(np.abs(gz - pz)).item()])
    return np.array(md), np.array(mz)


def compute_scale_and_shift(prediction: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils.py:488-528: closed-form 2x2 least squares per row."""
    prediction, target, mask = prediction.clone(), target.clone(), mask.clone()
    prediction[~mask.bool()] = 0.0
    target[~mask.bool()] = 0.0
    a00, a01, a11 = (mask * prediction * prediction).sum(1), (mask * prediction).sum(1), mask.sum(1)
    b0, b1 = (mask * prediction * target).sum(1), (mask * target).sum(1)
    det = a00 * a11 - a01 * a01
    if not torch.all(det > 0.):
        det = det + 1e-4
    return torch.stack([(a11 * b0 - a01 * b1) / det, (-a01 * b0 + a00 * b1) / det]).transpose(0, 1)


def normalize_ssi(x: torch.Tensor, mask: torch.Tensor, t: float = 0.5, s: float = 2.0):
    """utils.py:172-180 with low_p = 0, high_p = 1 (min / max of the valid pixels)."""
    m = mask.bool()
    low, up = x[m].min(), x[m].max()
    y = torch.zeros_like(x)
    y[m] = (torch.clamp((x[m] - low) / (up - low), 0, 1) - t) * s
    return y, low, up


def denormalize_ssi(y: torch.Tensor, raw_disp: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils.py:232-251 with safe_ssi=False."""
    B, R, H, W = y.shape
    st = compute_scale_and_shift(y.reshape(-1, H * W), raw_disp.repeat(1, R, 1, 1).reshape(-1, H * W),
                                 mask.repeat(1, R, 1, 1).reshape(-1, H * W))
    s, t = torch.split(st.view(B, R, 1, 2), 1, dim=-1)
    return y * s + t

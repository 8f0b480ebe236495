"""Evaluation metrics and scale-shift alignment on the device, with the reference's function surfaces:

* ``compute_errors(gt_disps, pred_disps, space="disp", masks=None, fxb=None)`` — /root/reference/utils_d3roma/utils.py:554-653
* ``compute_scale_and_shift(prediction, target, mask=None)``                   — utils_d3roma/utils.py:488-528
* ``metrics_to_dict``                                                          — utils_d3roma/utils.py:657-661

The reference computes these in numpy on the host after ``.cpu()`` (evaluate.py:122-138), one Python loop iteration per
sample; here each is one reduction kernel over the whole batch and only the [B, 11] metric rows leave the GPU (they are
also what ``distributed_evaluate.py:162-163`` all-gathers).  Reference quirks are kept: the D-k thresholds are scaled by
``n_valid / 480`` (utils.py:629) and non-positive predictions are clipped to the smallest valid
ground-truth disparity (utils.py:596-601).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops

DISP_KEYS = ["epe", "d1", "d2", "d3", "d5"]
DEPTH_KEYS = ["a1", "a2", "a3", "rmse", "rel", "mae"]


def compute_errors_device(gt_disps: torch.Tensor, pred_disps: torch.Tensor, masks: Optional[torch.Tensor] = None,
                          fxb=None) -> torch.Tensor:
    """[B,H,W] CUDA tensors -> [B, 11] fp64 CUDA tensor (columns DISP_KEYS + DEPTH_KEYS)."""
    assert pred_disps.dim() == gt_disps.dim() == 3, "pred_disps should be BxHxW"
    if masks is None:
        masks = torch.isfinite(gt_disps) & (gt_disps != 0)
    return ops.disp_metrics(gt_disps, pred_disps, masks, fxb)


def compute_errors(gt_disps, pred_disps, space: str = "disp", masks=None, fxb=None):
    """Reference-shaped wrapper: accepts numpy arrays or tensors, returns (metrics_disp [B,5], metrics_depth [B,6])."""
    if space != "disp":
        raise NotImplementedError("only prediction_space == 'disp' (every pixel-space config) is implemented")
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.as_tensor(np.asarray(gt_disps) if not torch.is_tensor(gt_disps) else gt_disps).to(dev, torch.float32)
    p = torch.as_tensor(np.asarray(pred_disps) if not torch.is_tensor(pred_disps) else pred_disps).to(dev, torch.float32)
    m = None if masks is None else torch.as_tensor(np.asarray(masks) if not torch.is_tensor(masks) else masks).to(dev)
    out = compute_errors_device(g, p, m, fxb).cpu().numpy()
    return out[:, :5], out[:, 5:]


def compute_scale_and_shift(prediction: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """[B, HW] -> [B, 2] (scale, shift) minimising ||mask * (prediction * s + t - target)||^2."""
    assert prediction.shape == target.shape, "shape mismatch"
    B = prediction.shape[0]
    if mask is None:
        mask = torch.ones_like(target)
    return ops.scale_shift_lsq(prediction.reshape(B, -1), target.reshape(B, -1), mask.reshape(B, -1))


def metrics_to_dict(metrics_disp, metrics_depth):
    return {"disp": dict(zip(DISP_KEYS, np.array(metrics_disp).mean(0))),
            "depth": dict(zip(DEPTH_KEYS, np.array(metrics_depth).mean(0)))}

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

import ctypes

import numpy as np
import torch

from . import _lib, ops

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


class RANSAC:
    """Robust scale/shift fit of /root/reference/utils_d3roma/ransac.py:39-96 (``RANSAC(n, k, t, d).fit(X, Y, mask)``
    -> ``best_fit`` [B, 2]), one fused C-ABI call per iteration (``d3_ransac_iteration``).  The random permutation of each
    iteration is drawn like the reference draws it, ``torch.randperm(HW, device=X.device)``; ``perm_fn`` lets a test
    inject the permutations of a recorded reference run."""

    def __init__(self, n=0.1, k=100, t=0.05, d=0.5, perm_fn=None):
        self.n, self.k, self.t, self.d = n, k, t, d
        self.perm_fn = perm_fn
        self.best_fit = self.best_error = self.best_num_inlier = None

    def fit(self, X: torch.Tensor, Y: torch.Tensor, mask: torch.Tensor):
        assert X.shape == Y.shape == mask.shape
        B, HW = X.shape
        N, D = int(self.n * HW), int(self.d * HW)
        assert D < HW and N < HW, "N, D must be less than HW"
        dev = X.device
        Xc, Yc = X.to(torch.float32).contiguous(), Y.to(torch.float32).contiguous()
        Mc = mask.to(torch.float32).contiguous()
        lib = _lib.lib()
        ws = torch.empty((lib.d3_ransac_workspace_bytes(B, HW),), dtype=torch.uint8, device=dev)
        self.best_fit = torch.tensor([1.0, 0.0], device=dev).repeat(B, 1).contiguous()     # s = 1, t = 0
        self.best_error = torch.full((B, 1), float("inf"), device=dev)
        best_num = torch.zeros((B, 1), device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(self.k):
            perm = self.perm_fn(HW, dev) if self.perm_fn is not None else torch.randperm(HW, device=dev)
            perm = perm[:N].to(device=dev, dtype=torch.int32).contiguous()
            _lib.check(lib.d3_ransac_iteration(Xc.data_ptr(), Yc.data_ptr(), Mc.data_ptr(), perm.data_ptr(), N, B, HW,
                                               float(self.t), float(D), ws.data_ptr(), self.best_fit.data_ptr(),
                                               self.best_error.data_ptr(), best_num.data_ptr(), stream))
        self.best_num_inlier = best_num.to(torch.int32)
        return self


def metrics_to_dict(metrics_disp, metrics_depth):
    return {"disp": dict(zip(DISP_KEYS, np.array(metrics_disp).mean(0))),
            "depth": dict(zip(DEPTH_KEYS, np.array(metrics_depth).mean(0)))}

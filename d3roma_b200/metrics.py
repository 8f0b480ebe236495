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


_ENSEMBLE_OBJECTIVE = None      # tests/: a torch emulation of d3_ensemble_objective / d3_ensemble_reduce (host logic on CPU)
_ENSEMBLE_REDUCE = None


def _ens_objective(images: torch.Tensor, s: np.ndarray, t: np.ndarray, median: bool, ws: torch.Tensor,
                   out3: torch.Tensor) -> torch.Tensor:
    if _ENSEMBLE_OBJECTIVE is not None:
        return _ENSEMBLE_OBJECTIVE(images, s, t, median)
    n, P = images.shape[0], images[0].numel()
    f = ctypes.POINTER(ctypes.c_float)
    _lib.check(_lib.lib().d3_ensemble_objective(images.data_ptr(), n, P, s.ctypes.data_as(f), t.ctypes.data_as(f),
                                                int(median), ws.data_ptr(), out3.data_ptr(), ops._stream()))
    return out3


def ensemble_depths(input_images: torch.Tensor, regularizer_strength: float = 0.02, max_iter: int = 2, tol: float = 1e-3,
                    reduction: str = "median", max_res: int = None):
    """``ensemble_depths`` of /root/reference/utils_d3roma/utils.py:28-120 (``config.ensemble``, evaluate.py:117-120):
    ``input_images`` fp32 CUDA [n, H, W] (the R rounds of one sample) -> ``(aligned_images [H, W], uncertainty [H, W])``.

    The host side is the reference's: the same initial guess (per-map 1 / (max - min) and -min / (max - min), float32), the
    same ``scipy.optimize.minimize(closure, x, method="BFGS", tol=tol, options={"maxiter": max_iter, "disp": False})``
    call with a closure that returns a float32 scalar.  The closure body (transform, pair distances, median / mean, near
    and far error) is ONE kernel launch over the maps (``d3_ensemble_objective``) and 24 bytes back, the final prediction
    (median + median absolute deviation, or mean + std, rescaled to [0, 1]) is ``d3_ensemble_reduce``.

    Note on what the optimiser does: the iterate is float32, so scipy's forward differences (step ~1.5e-8 * |x|) fall
    below its resolution; with the scipy versions that keep the float32 iterate (this image's among them) the gradient is
    0 / 0 and BFGS returns the initial guess after 0 iterations - the result is then "min-max normalise every map,
    reduce, min-max normalise again", and bit-exact parity with the reference holds (tests/golden/ensemble.npz).  With a
    scipy that upcasts the iterate both implementations descend on the rounding noise of their float32 objectives."""
    from scipy.optimize import minimize
    if reduction not in ("mean", "median"):
        raise ValueError(f"Unknown reduction method: {reduction}")
    if max_res is not None:
        # utils.py:48-52 calls torch.from_numpy on a tensor here and raises for every input
        raise NotImplementedError("ensemble_depths: max_res is not supported (the reference's branch raises a TypeError)")
    emulated = _ENSEMBLE_OBJECTIVE is not None
    if input_images.dim() != 3:
        raise ValueError("ensemble_depths: [n, H, W] expected")
    images = input_images.to(torch.float32).contiguous() if emulated else ops._f32c(input_images, "input_images")
    n = images.shape[0]
    if n < 2:
        # the reference's closure concatenates the pair differences of torch.combinations: with one map the list is empty
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors (ensemble_depths needs at least two maps)")
    if n > 32:
        raise ValueError(f"ensemble_depths: at most 32 maps (got {n})")
    median = reduction == "median"
    P = images[0].numel()
    dev = images.device
    ws = out3 = None
    if not emulated:
        ws = torch.empty((_lib.lib().d3_ensemble_workspace_bytes(),), dtype=torch.uint8, device=dev)
        out3 = torch.empty((3,), dtype=torch.float64, device=dev)
    # init guess (utils.py:54-59), float32 like np.min / np.max of a float32 array
    mm = (torch.stack([images.reshape(n, -1).min(1).values, images.reshape(n, -1).max(1).values], 1) if emulated
          else ops.masked_minmax(images.reshape(n, -1))).cpu().numpy().astype(np.float32)
    _min, _max = mm[:, 0], mm[:, 1]
    s_init = (1.0 / (_max - _min)).reshape((-1, 1, 1))
    t_init = (-1 * s_init.flatten() * _min.flatten()).reshape((-1, 1, 1))
    x = np.concatenate([s_init, t_init]).reshape(-1).astype(np.float32)
    n_pairs = n * (n - 1) // 2

    def closure(x):                                   # utils.py:64-87
        l = len(x)
        s = np.ascontiguousarray(x[: int(l / 2)], dtype=np.float32)     # torch.from_numpy(s).to(dtype=float32)
        t = np.ascontiguousarray(x[int(l / 2):], dtype=np.float32)
        acc, lo, hi = (float(v) for v in _ens_objective(images, s, t, median, ws, out3).cpu())
        sqrt_dist = np.float32(np.sqrt(np.float32(acc / (n_pairs * P))))
        near_err = np.float32(abs(0.0 - lo))          # sqrt((0 - min(pred)) ** 2)
        far_err = np.float32(abs(np.float32(1.0) - np.float32(hi)))
        err = np.float32(sqrt_dist + np.float32((near_err + far_err) * np.float32(regularizer_strength)))
        return np.asarray(err, dtype=np.float32)

    res = minimize(closure, x, method="BFGS", tol=tol, options={"maxiter": max_iter, "disp": False})
    x = res.x
    l = len(x)
    s = np.ascontiguousarray(x[: int(l / 2)], dtype=np.float32)
    t = np.ascontiguousarray(x[int(l / 2):], dtype=np.float32)
    if emulated:
        return _ENSEMBLE_REDUCE(images, s, t, median)
    aligned = torch.empty(images.shape[1:], dtype=torch.float32, device=dev)
    uncertainty = torch.empty_like(aligned)
    f = ctypes.POINTER(ctypes.c_float)
    _lib.check(_lib.lib().d3_ensemble_reduce(images.data_ptr(), n, P, s.ctypes.data_as(f), t.ctypes.data_as(f),
                                             int(median), ws.data_ptr(), aligned.data_ptr(), uncertainty.data_ptr(),
                                             ops._stream()))
    return aligned, uncertainty

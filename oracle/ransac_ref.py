"""ORACLE (test infrastructure, not product code): CPU restatement of the robust scale/shift fit of
/root/reference/utils_d3roma/ransac.py:39-96 (``RANSAC.fit``) with the permutations passed in, pinned to the reference's
own run in ``tests/golden/ransac.npz`` (``oracle/gen_golden.py::gen_ransac`` records the permutations the reference
drew).  Only ``tests/`` imports this.
"""
from __future__ import annotations

import torch

from .postproc_ref import compute_scale_and_shift


def ransac_fit(X: torch.Tensor, Y: torch.Tensor, mask: torch.Tensor, perms, n=0.1, t=0.05, d=0.5):
    """ransac.py:51-94 with ``perms[k]`` standing for ``torch.randperm(HW)`` of iteration k.
    Returns (best_fit [B,2], best_error [B,1], best_num_inlier [B,1])."""
    B, HW = X.shape
    N, D = int(n * HW), int(d * HW)
    best_num = torch.zeros((B, 1), dtype=torch.int64)
    best_err = torch.full((B, 1), float("inf"))
    best_fit = torch.tensor([1.0, 0.0]).repeat(B, 1)
    for perm in perms:
        ids = torch.as_tensor(perm, dtype=torch.int64)[:N].repeat(B, 1)
        maybe = compute_scale_and_shift(torch.gather(X, 1, ids), torch.gather(Y, 1, ids), torch.gather(mask, 1, ids))
        X_ = X * maybe[:, 0:1] + maybe[:, 1:]
        thr = ((Y - X_) ** 2 < t) & mask.bool()
        better = compute_scale_and_shift(X, Y, thr.float())
        X__ = X * better[:, 0:1] + better[:, 1:]
        acc = torch.maximum(Y / X__, X__ / Y)
        this_err = 1 / torch.mean((acc < 1.25).float())          # no dim: one score for all rows (sic)
        num = thr.sum(1, keepdim=True)
        sel = (num > D) & (this_err < best_err)
        best_num = torch.where(sel, num, best_num)
        best_fit = torch.where(sel, better, best_fit)
        best_err = torch.where(sel, this_err.expand_as(best_err), best_err)
    return best_fit, best_err, best_num

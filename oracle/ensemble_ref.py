"""ORACLE (test infrastructure only - never imported by the product path): CPU restatement of ``ensemble_depths``
(/root/reference/utils_d3roma/utils.py:28-120, with ``inter_distances`` :15-25), the ``config.ensemble`` branch of
``eval_batch`` (evaluate.py:117-120).  Pinned: ``oracle/gen_golden.py::gen_ensemble`` records the reference's own
function on seeded inputs (tests/golden/ensemble.npz) and ``tests/test_oracle_postproc.py`` compares this file with it
bit for bit.

The closure and the final prediction are split out (``objective`` / ``reduce``) because those are the two device entry
points of the product (d3_ensemble_objective / d3_ensemble_reduce); ``ensemble_depths`` glues them with the reference's
scipy call."""
from __future__ import annotations

import numpy as np
import torch


def make_inputs(n: int, hw=(24, 40), seed: int = 0, ties: bool = False) -> torch.Tensor:
    """n affine-distorted, noisy copies of one random map (what R sampling rounds of one sample look like)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(*hw, generator=g)
    maps = [base * (0.5 + torch.rand(1, generator=g)) + 0.3 * torch.rand(1, generator=g) +
            0.02 * torch.randn(*hw, generator=g) for _ in range(n)]
    if ties and n >= 3:
        maps[2] = maps[0].clone()       # two identical maps: equal values inside the median
    return torch.stack(maps)


def transform(images: torch.Tensor, s, t) -> torch.Tensor:
    s = torch.as_tensor(np.asarray(s, dtype=np.float32))
    t = torch.as_tensor(np.asarray(t, dtype=np.float32))
    return images * s.view(-1, 1, 1) + t.view(-1, 1, 1)                    # utils.py:70


def objective(images: torch.Tensor, s, t, median: bool):
    """-> (sum of squared pair distances, min of pred, max of pred): the three reductions of the closure (utils.py:64-87)."""
    tr = transform(images, s, t)
    n = tr.shape[0]
    acc = 0.0
    for i in range(n):                                                     # torch.combinations order (:20)
        for j in range(i + 1, n):
            d = (tr[i] - tr[j]).double()
            acc += float((d * d).sum())
    pred = torch.median(tr, dim=0).values if median else torch.mean(tr, dim=0)
    return acc, float(pred.min()), float(pred.max())


def reduce(images: torch.Tensor, s, t, median: bool):
    """utils.py:99-118."""
    tr = transform(images, s, t)
    if median:
        aligned = torch.median(tr, dim=0).values
        unc = torch.median(torch.abs(tr - aligned), dim=0).values
    else:
        aligned = torch.mean(tr, dim=0)
        unc = torch.std(tr, dim=0)
    _min, _max = torch.min(aligned), torch.max(aligned)
    return (aligned - _min) / (_max - _min), unc / (_max - _min)


def ensemble_depths(images: torch.Tensor, regularizer_strength=0.02, max_iter=2, tol=1e-3, reduction="median"):
    from scipy.optimize import minimize
    n = images.shape[0]
    flat = images.reshape(n, -1).numpy()
    _min, _max = np.min(flat, axis=1), np.max(flat, axis=1)
    s_init = 1.0 / (_max - _min).reshape((-1, 1, 1))
    t_init = (-1 * s_init.flatten() * _min.flatten()).reshape((-1, 1, 1))
    x = np.concatenate([s_init, t_init]).reshape(-1).astype(np.float32)
    median = reduction == "median"
    P, pairs = images[0].numel(), n * (n - 1) // 2

    def closure(x):
        l = len(x)
        acc, lo, hi = objective(images, x[: int(l / 2)], x[int(l / 2):], median)
        err = np.float32(np.sqrt(np.float32(acc / (pairs * P)))) + \
            np.float32((np.float32(abs(lo)) + np.float32(abs(np.float32(1.0) - np.float32(hi)))) * np.float32(regularizer_strength))
        return np.asarray(err, dtype=np.float32)

    res = minimize(closure, x, method="BFGS", tol=tol, options={"maxiter": max_iter, "disp": False})
    l = len(res.x)
    return reduce(images, res.x[: int(l / 2)], res.x[int(l / 2):], median) + (res,)

"""Pin oracle/postproc_ref.py against the reference's own compute_errors / compute_scale_and_shift / Normalizer(ssi)
outputs (tests/golden/postproc.npz)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import postproc_ref as P
from tests._golden import SOURCES, load, same

G = Path(__file__).parent / "golden"


def test_compute_errors_matches_reference():
    z = np.load(G / "postproc.npz")
    md, mz = P.compute_errors_disp(z["err/gt"], z["err/pred"], z["err/mask"], [24.5471958] * 3)
    np.testing.assert_array_equal(md, z["err/disp"])
    np.testing.assert_array_equal(mz, z["err/depth"])
    pred2 = np.where(z["err/mask"], np.abs(z["err/pred"]) + 0.1, 0).astype(np.float32)
    md, mz = P.compute_errors_disp(z["err/gt"], pred2, None, None)
    np.testing.assert_array_equal(md, z["err/disp_nofxb"])
    np.testing.assert_array_equal(mz, z["err/depth_nofxb"])


def test_scale_shift_and_ssi_normalizer_match_reference():
    z = np.load(G / "postproc.npz")
    st = P.compute_scale_and_shift(torch.from_numpy(z["ssi/pred"]), torch.from_numpy(z["ssi/target"]),
                                   torch.from_numpy(z["ssi/mask"]).float())
    np.testing.assert_array_equal(st.numpy(), z["ssi/st"])
    y, low, up = P.normalize_ssi(torch.from_numpy(z["ssi/norm_in"]), torch.from_numpy(z["ssi/norm_mask"]))
    np.testing.assert_array_equal(y.numpy(), z["ssi/norm_out"])
    np.testing.assert_array_equal(np.array([float(low), float(up)], dtype=np.float32), z["ssi/low_up"])
    d = P.denormalize_ssi(torch.from_numpy(z["ssi/den_in"]), torch.from_numpy(z["ssi/den_raw"]),
                          torch.from_numpy(z["ssi/den_mask"]).float())
    np.testing.assert_array_equal(d.numpy(), z["ssi/den_out"])


def test_ransac_oracle_matches_reference_run():
    """oracle/ransac_ref.py fed with the permutations the reference drew reproduces the reference's RANSAC.fit
    (utils_d3roma/ransac.py:51-96) on the golden problem."""
    import numpy as np
    import torch
    from pathlib import Path
    from oracle.ransac_ref import ransac_fit
    g = np.load(Path(__file__).parent / "golden" / "ransac.npz")
    fit, err, num = ransac_fit(torch.from_numpy(g["pred"]), torch.from_numpy(g["gt"]), torch.from_numpy(g["mask"]),
                               g["perms_fit"], n=0.1, t=0.6, d=0.2)
    np.testing.assert_allclose(fit.numpy(), g["best_fit"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(err.numpy(), g["best_error"], rtol=1e-6)
    np.testing.assert_array_equal(num.numpy(), g["best_num_inlier"])
    assert (g["best_num_inlier"] > 0).all()          # the problem is one where RANSAC does select models


# ---------------------------------------------------------------------------------------------- ensemble_depths
ENSEMBLE_CASES = (("med5", 5, 0, False, "median"), ("med4", 4, 1, False, "median"), ("med6t", 6, 2, True, "median"),
                  ("mean3", 3, 3, False, "mean"), ("mean5t", 5, 4, True, "mean"), ("med2", 2, 5, False, "median"))


@pytest.mark.parametrize("source", SOURCES)
@pytest.mark.parametrize("tag,n,seed,ties,red", ENSEMBLE_CASES)
def test_ensemble_depths_restatement(source, tag, n, seed, ties, red):
    """oracle/ensemble_ref.py against the reference's own ensemble_depths (utils.py:28-120)."""
    import warnings
    from oracle import ensemble_ref as E
    G = load("ensemble", source)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aligned, unc, res = E.ensemble_depths(E.make_inputs(n, seed=seed, ties=ties), reduction=red)
    assert same(aligned.numpy(), G[f"{tag}/aligned"], source, rtol=1e-6, atol=1e-6)
    assert same(unc.numpy(), G[f"{tag}/uncertainty"], source, rtol=1e-6, atol=1e-6)
    assert float(aligned.min()) == 0.0 and float(aligned.max()) == 1.0


@pytest.mark.parametrize("tag,n,seed,ties,red", ENSEMBLE_CASES)
def test_ensemble_depths_host_logic(monkeypatch, tag, n, seed, ties, red):
    """d3roma_b200.metrics.ensemble_depths (initial guess, scipy call, closure arithmetic) with the two device entry
    points swapped for the oracle's torch forms: the reference's result."""
    import warnings
    from d3roma_b200 import metrics
    from oracle import ensemble_ref as E
    calls = []

    def objective(images, s, t, median):
        calls.append(1)
        return torch.tensor(E.objective(images, s, t, median), dtype=torch.float64)

    monkeypatch.setattr(metrics, "_ENSEMBLE_OBJECTIVE", objective)
    monkeypatch.setattr(metrics, "_ENSEMBLE_REDUCE", lambda images, s, t, median: E.reduce(images, s, t, median))
    G = load("ensemble")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aligned, unc = metrics.ensemble_depths(E.make_inputs(n, seed=seed, ties=ties), reduction=red)
    assert len(calls) >= 1 + 2 * n                      # the value at x0 and one forward difference per variable
    assert same(aligned.numpy(), G[f"{tag}/aligned"], rtol=1e-6, atol=1e-6)
    assert same(unc.numpy(), G[f"{tag}/uncertainty"], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        metrics.ensemble_depths(E.make_inputs(3), reduction="mode")
    with pytest.raises(NotImplementedError):
        metrics.ensemble_depths(E.make_inputs(3), max_res=64)
    with pytest.raises(RuntimeError):                   # one map: the reference's torch.cat of an empty list
        metrics.ensemble_depths(E.make_inputs(1))

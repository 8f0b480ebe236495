"""GPU parity of the device-side evaluation post-processing against the reference's outputs (tests/golden/postproc.npz)
and against the oracle at full size."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"


def test_compute_errors_vs_reference_golden():
    import d3roma_b200 as d
    z = np.load(G / "postproc.npz")
    md, mz = d.compute_errors(z["err/gt"], z["err/pred"], "disp", z["err/mask"], [24.5471958] * 3)
    np.testing.assert_allclose(md, z["err/disp"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mz, z["err/depth"], rtol=1e-5, atol=1e-7)
    pred2 = np.where(z["err/mask"], np.abs(z["err/pred"]) + 0.1, 0).astype(np.float32)
    md, mz = d.compute_errors(z["err/gt"], pred2, "disp", None, None)
    np.testing.assert_allclose(md, z["err/disp_nofxb"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mz, z["err/depth_nofxb"], rtol=2e-5, atol=1e-7)
    dd = d.metrics_to_dict(md, mz)
    assert set(dd["disp"]) == {"epe", "d1", "d2", "d3", "d5"} and set(dd["depth"]) == {"a1", "a2", "a3", "rmse", "rel", "mae"}
    with pytest.raises(NotImplementedError):
        d.compute_errors(z["err/gt"], z["err/pred"], "depth", z["err/mask"], [1.0] * 3)


def test_compute_errors_full_size_vs_oracle():
    import d3roma_b200 as d
    from oracle import postproc_ref as P
    g = torch.Generator().manual_seed(2)
    B, H, W = 8, 360, 640
    gt = torch.rand(B, H, W, generator=g) * 100 + 12
    gt[torch.rand(B, H, W, generator=g) < 0.2] = 0
    pred = gt + torch.randn(B, H, W, generator=g) * 2
    md, mz = d.compute_errors(gt, pred, "disp", None, [24.5] * B)
    rmd, rmz = P.compute_errors_disp(gt.numpy(), pred.numpy(), None, [24.5] * B)
    np.testing.assert_allclose(md, rmd, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(mz, rmz, rtol=2e-5, atol=1e-7)


def test_scale_shift_and_ssi_normalizer_vs_reference_golden():
    import d3roma_b200 as d
    z = np.load(G / "postproc.npz")
    st = d.compute_scale_and_shift(torch.from_numpy(z["ssi/pred"]).cuda(), torch.from_numpy(z["ssi/target"]).cuda(),
                                   torch.from_numpy(z["ssi/mask"]).float().cuda())
    np.testing.assert_allclose(st.cpu().numpy(), z["ssi/st"], rtol=2e-5, atol=2e-5)
    n = d.Normalizer(ssi=True, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0, safe_ssi=False)
    y, low, up = n.normalize(torch.from_numpy(z["ssi/norm_in"]).cuda(), torch.from_numpy(z["ssi/norm_mask"]).cuda())
    assert np.array_equal(np.array([float(low), float(up)], dtype=np.float32), z["ssi/low_up"])
    assert np.array_equal(y.cpu().numpy(), z["ssi/norm_out"])        # elementwise part is bit-exact
    out = n.denormalize(torch.from_numpy(z["ssi/den_in"]).cuda(), torch.from_numpy(z["ssi/den_raw"]).cuda(),
                        torch.from_numpy(z["ssi/den_mask"]).float().cuda())
    np.testing.assert_allclose(out.cpu().numpy(), z["ssi/den_out"], rtol=3e-5, atol=3e-4)


def test_ransac_and_safe_ssi_denormalize_match_reference_golden():
    """metrics.RANSAC (d3_ransac_iteration) and Normalizer(ssi=True, safe_ssi=True).denormalize fed with the
    permutations of the recorded reference run (tests/golden/ransac.npz, utils_d3roma/ransac.py + utils.py:240-251)."""
    from pathlib import Path
    import d3roma_b200 as d
    from d3roma_b200.metrics import RANSAC
    g = np.load(Path(__file__).parent / "golden" / "ransac.npz")
    X, Y, Mk = (torch.from_numpy(g[k]).cuda() for k in ("pred", "gt", "mask"))
    perms = iter(g["perms_fit"])
    reg = RANSAC(n=0.1, k=10, d=0.2, t=0.6, perm_fn=lambda hw, dev: torch.from_numpy(next(perms))).fit(X, Y, Mk)
    np.testing.assert_allclose(reg.best_fit.cpu().numpy(), g["best_fit"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(reg.best_error.cpu().numpy(), g["best_error"], rtol=1e-6)
    np.testing.assert_array_equal(reg.best_num_inlier.cpu().numpy(), g["best_num_inlier"])
    n = d.Normalizer(ssi=True, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0, safe_ssi=True,
                     ransac_error_threshold=0.6)
    perms2 = iter(g["perms_den"])
    n.perm_fn = lambda hw, dev: torch.from_numpy(next(perms2))
    B, _, H, W = g["den_out"].shape
    out = n.denormalize(X.view(B, 1, H, W), Y.view(B, 1, H, W), Mk.view(B, 1, H, W))
    np.testing.assert_allclose(out.cpu().numpy(), g["den_out"], rtol=1e-4, atol=1e-3)
    # default permutation source: runs and returns a sane fit
    reg2 = RANSAC(n=0.1, k=10, d=0.2, t=0.6).fit(X, Y, Mk)
    assert torch.isfinite(reg2.best_fit).all() and (reg2.best_num_inlier > 0).all()

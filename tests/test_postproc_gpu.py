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


# ---------------------------------------------------------------------------------------------- ensemble_depths
ENSEMBLE_CASES = (("med5", 5, 0, False, "median"), ("med4", 4, 1, False, "median"), ("med6t", 6, 2, True, "median"),
                  ("mean3", 3, 3, False, "mean"), ("mean5t", 5, 4, True, "mean"), ("med2", 2, 5, False, "median"))


def _ens_call(name, images, s, t, median, *outs):
    import ctypes
    from d3roma_b200 import _lib, ops
    f = ctypes.POINTER(ctypes.c_float)
    ws = torch.empty((_lib.lib().d3_ensemble_workspace_bytes(),), dtype=torch.uint8, device="cuda")
    s, t = np.ascontiguousarray(s, dtype=np.float32), np.ascontiguousarray(t, dtype=np.float32)
    _lib.check(getattr(_lib.lib(), name)(images.data_ptr(), images.shape[0], images[0].numel(), s.ctypes.data_as(f),
                                         t.ctypes.data_as(f), int(median), ws.data_ptr(), *[o.data_ptr() for o in outs],
                                         ops._stream()))


@pytest.mark.parametrize("n,hw,median", [(5, (24, 40), True), (4, (33, 17), True), (3, (24, 40), False),
                                         (8, (360, 640), True), (32, (8, 9), True), (1, (5, 7), True), (7, (360, 640), False)])
def test_ensemble_kernels_vs_oracle(n, hw, median):
    """d3_ensemble_objective / d3_ensemble_reduce against the oracle's torch forms of the closure and of the final
    prediction (utils.py:64-87, 99-118) for arbitrary scale / shift vectors."""
    from oracle import ensemble_ref as E
    x = E.make_inputs(n, hw=hw, seed=n)
    rng = np.random.default_rng(n)
    s, t = (0.5 + rng.random(n)).astype(np.float32), (rng.random(n) - 0.5).astype(np.float32)
    xd = x.cuda()
    out3 = torch.empty(3, dtype=torch.float64, device="cuda")
    _ens_call("d3_ensemble_objective", xd, s, t, median, out3)
    acc, lo, hi = E.objective(x, s, t, median)
    got = out3.cpu().numpy()
    assert abs(got[0] - acc) <= 1e-6 * max(acc, 1e-30)
    if median:
        assert got[1] == lo and got[2] == hi                    # the median is one of the inputs: exact
    else:
        assert abs(got[1] - lo) < 1e-6 and abs(got[2] - hi) < 1e-6
    out3b = torch.empty(3, dtype=torch.float64, device="cuda")
    _ens_call("d3_ensemble_objective", xd, s, t, median, out3b)
    assert torch.equal(out3, out3b)                             # deterministic (BFGS differentiates it numerically)
    if n == 1 and not median:
        return
    al, un = torch.empty(hw, device="cuda"), torch.empty(hw, device="cuda")
    _ens_call("d3_ensemble_reduce", xd, s, t, median, al, un)
    wal, wun = E.reduce(x, s, t, median)
    if median:
        assert torch.equal(al.cpu(), wal) and torch.equal(un.cpu(), wun)
    else:
        assert float((al.cpu() - wal).abs().max()) < 1e-6 and float((un.cpu() - wun).abs().max()) < 1e-6


@pytest.mark.parametrize("tag,n,seed,ties,red", ENSEMBLE_CASES)
def test_ensemble_depths_vs_reference_golden(tag, n, seed, ties, red):
    """d3roma_b200.metrics.ensemble_depths against the reference's own function (tests/golden/ensemble.npz)."""
    import warnings
    from d3roma_b200 import metrics
    from oracle import ensemble_ref as E
    z = np.load(G / "ensemble.npz")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aligned, unc = metrics.ensemble_depths(E.make_inputs(n, seed=seed, ties=ties).cuda(), reduction=red)
    assert aligned.is_cuda and aligned.shape == (24, 40)
    tol = 0.0 if red == "median" else 1e-6
    assert np.abs(aligned.cpu().numpy() - z[f"{tag}/aligned"]).max() <= tol
    assert np.abs(unc.cpu().numpy() - z[f"{tag}/uncertainty"]).max() <= tol
    with pytest.raises(ValueError):                                 # no CPU fallback
        metrics.ensemble_depths(E.make_inputs(3))


def test_ensemble_depths_recovers_common_map_at_full_size():
    """Size-independent property at 360 x 640: n affine copies of one map align to that map rescaled to [0, 1] with zero
    spread, whatever their scales and shifts."""
    import warnings
    from d3roma_b200 import metrics
    g = torch.Generator().manual_seed(9)
    base = torch.rand(360, 640, generator=g)
    x = torch.stack([base * a + b for a, b in ((1.0, 0.0), (2.5, -1.0), (0.3, 4.0), (7.0, 0.5), (1.7, -3.0))]).cuda()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        aligned, unc = metrics.ensemble_depths(x)
    want = (base - base.min()) / (base.max() - base.min())
    assert float((aligned.cpu() - want).abs().max()) < 5e-6
    assert float(unc.max()) < 5e-6


@pytest.mark.parametrize("dataset", ["synthetic", "SceneFlow"])
def test_evaluate_intermediate_metrics_vs_oracle(dataset):
    """evaluate.py:28-57: metric rows of the plotted x0 estimates, against the oracle's compute_errors over the same
    maps prepared the reference's way (including its fxb[i % B] indexing)."""
    from types import SimpleNamespace
    from d3roma_b200 import evaluate
    from oracle import postproc_ref as P
    g = torch.Generator().manual_seed(13)
    B, N, H, W = 3, 4, 40, 56
    gt = torch.rand(B, 1, H, W, generator=g) * 60 + 1
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.3).float()
    pred = gt.repeat(1, N, 1, 1) + torch.randn(B, N, H, W, generator=g) * 3
    pred[0, 1, :4] = -2.0                                     # clamped to 0.25 / the SceneFlow floor
    fxb = torch.tensor([24.5, 30.0, 18.0])
    cfg = evaluate.default_config(eval_dataset=[dataset], camera_resolution="480x270")
    md, mz = evaluate.evaluate_intermediate_metrics(cfg, fxb.cuda(), SimpleNamespace(images_pred_orig=pred.cuda()),
                                                    gt.cuda(), mask.cuda(), None)
    assert md.shape == (B, N, 5) and mz.shape == (B, N, 6)
    p, m = pred.clone(), mask.repeat(1, N, 1, 1).bool()
    if dataset == "SceneFlow":
        lo, hi = 0.5 / 2.0, 192.0 / 2.0
        m = (p >= lo) & (p < hi) & m
        p = p.clamp(lo, hi)
    else:
        p = p.clamp(min=0.25)
    fx = fxb.view(B, 1).repeat(B, N, 1).view(-1, 1).numpy()
    rd, rz = P.compute_errors_disp(gt.repeat(1, N, 1, 1).view(-1, H, W).numpy(), p.view(-1, H, W).numpy(),
                                   m.view(-1, H, W).numpy(), fx[:, 0])
    np.testing.assert_allclose(md.reshape(-1, 5), rd, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(mz.reshape(-1, 6), rz, rtol=2e-5, atol=1e-6)

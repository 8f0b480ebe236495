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
    with pytest.raises(NotImplementedError):
        d.Normalizer(ssi=True, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.]).denormalize(
            torch.zeros(1, 1, 4, 4, device="cuda"), torch.ones(1, 1, 4, 4, device="cuda"), torch.ones(1, 1, 4, 4, device="cuda"))

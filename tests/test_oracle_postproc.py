"""Pin oracle/postproc_ref.py against the reference's own compute_errors / compute_scale_and_shift / Normalizer(ssi)
outputs (tests/golden/postproc.npz)."""
from pathlib import Path

import numpy as np
import torch

from oracle import postproc_ref as P

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

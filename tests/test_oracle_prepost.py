"""Pin oracle/prepost_ref.py and oracle/guidance_ref.py against reference outputs (tests/golden/prepost.npz,
guidance.npz) including the reference's own 13 known answers (utils_d3roma/camera.py:1027-1044)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import guidance_ref, prepost_ref as P

from tests._golden import SOURCES, load, same

G = Path(__file__).parent / "golden"


def test_camera_known_answers():
    z = np.load(G / "prepost.npz")
    for (d, a, b, s), e in zip(z["camera/normalize_in"], z["camera/normalize_out"]):
        assert P.normalize_disp(d, a, b, s) == e
    for (d, a, b, s), e in zip(z["camera/unnormalize_in"], z["camera/unnormalize_out"]):
        assert P.unnormalize_disp(d, a, b, s) == e
    # the literal asserts of the reference's __main__
    assert P.normalize_disp(70, 20, 120) == 0.5 and P.normalize_disp(0, 20, 120) == -0.2
    assert P.unnormalize_disp(0.5, 20, 120) == 70 and P.unnormalize_disp(0.2 + 0.5, 20, 120, 0.2) == 70
    fxb, mind, maxd, shift = z["camera/sim"]
    np.testing.assert_allclose([fxb, mind, maxd], [24.5471958, 12.2735979, 122.735979], rtol=1e-7)
    disp = P.depth_to_disp(torch.from_numpy(z["camera/depth_in"]), float(fxb))
    assert np.array_equal(disp.numpy(), z["camera/disp_out"])


@pytest.mark.parametrize("source", SOURCES)
def test_normalizer_bit_exact(source):
    """gamma = 1 (every shipped config): bit-exact against the committed fixtures on any host.  gamma != 1 goes through
    torch's SIMD ``pow``: bit-exact against the reference run on this host, an ulp against the committed fixtures."""
    z = load("prepost", source)
    for tag, (bound, gamma, chs) in {"avg1": (128., 1., 1), "avg3": (128., 1., 3), "avg1g": (64., 0.5, 1)}.items():
        src = source if gamma != 1. else "live"          # "live" = exact comparison
        y = P.normalize_average(torch.from_numpy(z[f"norm/{tag}/in"]), bound, gamma, chs)
        assert same(y.numpy(), z[f"norm/{tag}/out"], src, atol=1e-6), tag
        d = P.denormalize_average(torch.from_numpy(z[f"norm/{tag}/den_in"]), bound, gamma, chs)
        assert same(d.numpy(), z[f"norm/{tag}/den_out"], src, atol=1e-5), tag


PIECEWISE = {"pw3": ([64., 32., 32.], [1., 1., 1.]), "pw2g": ([64., 64.], [0.5, 1.]), "pw1": ([128.], [1.])}


@pytest.mark.parametrize("source", SOURCES)
def test_normalizer_piecewise_bit_exact(source):
    z = load("piecewise", source)
    for tag, (bounds, gammas) in PIECEWISE.items():
        src = source if any(g != 1. for g in gammas) else "live"
        y = P.normalize_piecewise(torch.from_numpy(z[f"{tag}/in"]), bounds, gammas)
        assert same(y.numpy(), z[f"{tag}/out"], src, atol=1e-6), tag
        d = P.denormalize_piecewise(torch.from_numpy(z[f"{tag}/den_in"]), bounds, gammas)
        assert same(d.numpy(), z[f"{tag}/den_out"], src, atol=1e-5), tag


def test_input_padder():
    z = np.load(G / "prepost.npz")
    for hw in ((360, 640), (540, 960), (37, 50), (64, 96)):
        key = f"pad/{hw[0]}x{hw[1]}"
        pads = P.pad_amounts(*hw, 32)
        assert pads == list(z[f"{key}/pad"])
        if f"{key}/in" in z.files:
            x = torch.from_numpy(z[f"{key}/in"])
            y = P.pad_replicate(x, 32)
            assert np.array_equal(y.numpy(), z[f"{key}/out"])
            assert torch.equal(P.unpad(y, pads), x)
    assert P.pad_amounts(360, 640) == [0, 0, 12, 12] and P.pad_amounts(540, 960) == [0, 0, 2, 2]


def test_guidance_gradient_matches_reference():
    z = np.load(G / "guidance.npz")
    mind, maxd, shift = z["cam"]
    for tag in "ab":
        g, warp = guidance_ref.guidance_grad(torch.from_numpy(z[f"{tag}/disp"]), torch.from_numpy(z[f"{tag}/left"]),
                                             torch.from_numpy(z[f"{tag}/right"]), float(mind), float(maxd),
                                             float(shift))
        np.testing.assert_allclose(warp.numpy(), z[f"{tag}/warp"], atol=1e-6)
        np.testing.assert_allclose(g.numpy(), z[f"{tag}/grad"], rtol=1e-4, atol=1e-4)

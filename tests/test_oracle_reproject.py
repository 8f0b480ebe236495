"""Pin oracle/reproject_ref.py against tests/golden/reproject.npz (outputs of the reference's own camera object,
utils_d3roma/camera.py:610-701)."""
from pathlib import Path

import numpy as np

from oracle import reproject_ref as R

G = Path(__file__).resolve().parent / "golden"


def load():
    z = np.load(G / "reproject.npz")
    depth = z["depth_mm"].astype(np.float32) / np.float32(1000.0)
    Kd, Kc = z["K_depth"], z["K_color"]
    arr = lambda k: np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]])   # noqa: E731
    return z, depth, arr(Kd), arr(Kc)


def mm(o):
    return np.round(o.astype(np.float64) * 1000.0).astype(np.uint16)


def test_reprojection_matches_reference_camera():
    z, depth, Kd, Kc = load()
    H, W = z["hw"]
    T = z["T_fc"] @ z["color_to_depth"] @ z["T_fc"].T
    assert np.allclose(T, z["T_cl"])
    out_a = R.reproject(depth, np.linalg.inv(Kd), T, Kc, (H, W), True)
    out_r = R.reproject(depth, np.linalg.inv(Kd), T, Kc, (H, W), False)
    assert np.array_equal(mm(out_a), z["out_aggressive_mm"]) and np.array_equal(mm(out_r), z["out_round_mm"])
    oy, ox, ch, cw = z["crop"]
    out_c = R.reproject(depth[oy:oy + ch, ox:ox + cw], np.linalg.inv(z["cropped_K"]), T, Kc, (H, W), False,
                        crop=(int(Kc[1, 2] - z["cropped_K"][1, 2]), int(Kc[0, 2] - z["cropped_K"][0, 2]), ch, cw))
    # (Realsense.K is the colour intrinsics, camera.py:574-577: the crop offsets come from K_color)
    assert np.array_equal(mm(out_c), z["out_cropped_mm"])

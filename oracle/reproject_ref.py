"""ORACLE (test infrastructure, not product code): numpy restatement of the depth -> colour-frame reprojection of the
reference, ``DepthCamera.transform_depth_to_rgb_frame`` (/root/reference/utils_d3roma/camera.py:654-701) and
``transform_cropped_depth_to_rgb_frame`` (:610-652): float64 pinhole algebra and numpy's fancy-index scatter (for
colliding targets the last source pixel in row-major order wins; "aggressive_fill" = four corner fills in sequence).

Pinned against ``tests/golden/reproject.npz`` (the reference's own ``Realsense.default_real("wsl")`` camera object run on
synthetic depth maps; oracle/gen_golden.py ``gen_reproject``) in tests/test_oracle_reproject.py."""
from __future__ import annotations

import numpy as np


def reproject(depth: np.ndarray, Kinv_depth: np.ndarray, T_cl: np.ndarray, K_color: np.ndarray, canvas_hw,
              aggressive_fill: bool, crop=None) -> np.ndarray:
    H, W = depth.shape[:2]
    CH, CW = canvas_hw
    depth_1d = depth.reshape(1, H * W)
    xs, ys = np.meshgrid(range(W), range(H), indexing="xy")
    pix = np.concatenate((xs.reshape(1, -1).astype(np.float32), ys.reshape(1, -1).astype(np.float32),
                          np.ones((1, H * W), dtype=np.float32)), axis=0)
    cam = (Kinv_depth @ pix) * depth_1d
    valid = cam[2] > 0.
    cam = cam[:, valid]
    cam = T_cl[:3, :3] @ cam + T_cl[:3, 3:]
    q = K_color @ cam
    q[:2] /= q[2:3]
    out = np.zeros((CH, CW), dtype=np.float32)
    u, v = q[:2]

    def fill(uu, vv):
        uu, vv = np.round(uu).astype(np.uint32), np.round(vv).astype(np.uint32)
        ok = (uu < CW) & (vv < CH)
        out[vv[ok], uu[ok]] = depth_1d[0, valid][ok]

    if aggressive_fill:
        ul, ur = np.floor(u).astype(np.uint32), np.ceil(u).astype(np.uint32)
        vu, vb = np.floor(v).astype(np.uint32), np.ceil(v).astype(np.uint32)
        fill(ul, vu)
        fill(ul, vb)
        fill(ur, vu)
        fill(ur, vb)
    else:
        fill(u, v)
    if crop is not None:
        oy, ox, oh, ow = crop
        return out[oy:oy + oh, ox:ox + ow]
    return out

"""Pre/post-processing either side of the loop with the reference's class surfaces:

* ``Normalizer`` (non-SSI, ``mode="average"``) — /root/reference/utils_d3roma/utils.py:137-293
* ``InputPadder``                               — utils_d3roma/utils.py:297-325
* ``Camera`` scalars + ``depth_to_disp`` / ``disp_to_depth`` — utils_d3roma/camera.py:115-125, 599-602, 771-780;
  inference.py:128-131, 246-249

Every tensor op runs in ``libd3roma_b200.so`` (one fused pass each).  ``ssi=True`` (DREDS / ClearPose configs) is
supported with the direct least-squares alignment (``safe_ssi=False``) and the RANSAC variant (``safe_ssi=True``,
``metrics.RANSAC``); both channel layouts are implemented: "average" (the pixel-space configs) and "piecewise" (up
to three range-splitting channels, the LDM-family configs; utils.py:198-229, 275-293).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import ops


@dataclass
class Camera:
    """The camera scalars the path needs (utils_d3roma/camera.py:115-125: fxb = baseline*fx, min/max disparity)."""

    fxb: float
    min_disp: float
    max_disp: float
    shift: float = 0.0

    @property
    def fxb_depth(self) -> float:
        return self.fxb

    def normalize_disp(self, disp):
        return (disp - self.min_disp) / (self.max_disp - self.min_disp) + self.shift  # camera.py:771-776

    def unnormalize_disp(self, disp):
        return self.min_disp + (disp - self.shift) * (self.max_disp - self.min_disp)  # camera.py:778-780

    @staticmethod
    def from_intrinsics(fx: float, baseline: float, min_depth: float, max_depth: float, shift: float = 0.0):
        fxb = fx * baseline
        return Camera(fxb, fxb / max_depth, fxb / min_depth, shift)


def _sim_camera() -> Camera:
    # Realsense.default_sim at 640x360 (camera.py:379, 535-555): hfov 71.28 deg, baseline 0.055 m, depth 0.2-2 m
    fx = 640.0 / (2.0 * math.tan(math.radians(71.28) / 2.0))
    return Camera.from_intrinsics(fx, 0.055, 0.2, 2.0)


SIM_CAMERA = _sim_camera()


def normalize_disp(disp, min_disp=5, max_disp=120, shift=0):
    """camera.py:771-776 (scalar / array form; known answers at camera.py:1027-1044)."""
    return (disp - min_disp) / (max_disp - min_disp) + shift


def unnormalize_disp(disp, min_disp=5, max_disp=120, shift=0):
    """camera.py:778-780."""
    return min_disp + (disp - shift) * (max_disp - min_disp)


def depth_to_disp(depth: torch.Tensor, fxb: float) -> torch.Tensor:
    """inference.py:128-131: ``disp = fxb / depth where depth > 0 else 0``."""
    return ops.depth_to_disp(depth, float(fxb))


def disp_to_depth(disp: torch.Tensor, fxb: float) -> torch.Tensor:
    """inference.py:246-249."""
    return ops.disp_to_depth(disp, float(fxb))


class Normalizer:
    def __init__(self, ssi=False, mode="piecewise", num_chs=3, ch_bounds=[64, 32, 32], ch_gammas=[1., 1., 1.], t=0.5,
                 s=2.0, safe_ssi=True, ransac_error_threshold=0.6):
        self.ssi, self.mode, self.num_chs = ssi, mode, num_chs
        self.ch_bounds, self.ch_gammas = ch_bounds, ch_gammas
        self.t, self.s = t, s
        self.safe_ssi, self.ransac_error_threshold = safe_ssi, ransac_error_threshold
        self.low_p, self.high_p = 0.0, 1.0
        self.perm_fn = None   # test hook: source of RANSAC permutations (default torch.randperm on the device)

    @staticmethod
    def from_config(config):
        return Normalizer(config.ssi, config.normalize_mode, config.num_chs, config.ch_bounds, config.ch_gammas,
                          config.norm_t, config.norm_s, config.safe_ssi, config.ransac_error_threshold)

    def _check(self):
        if self.ssi:
            return
        if self.mode == "average":
            assert len(self.ch_bounds) == len(self.ch_gammas) == 1, "inconsistent params"
        else:   # "piecewise" (utils.py:198-229): one bound / gamma per channel
            assert len(self.ch_bounds) == len(self.ch_gammas) == self.num_chs, "inconsistent params"
            if not 1 <= self.num_chs <= 3:
                raise RuntimeError("not implemented")

    def normalize(self, x: torch.Tensor, mask=None, low=None, up=None):
        """[1,H,W] raw disparity -> ([num_chs,H,W] in [-1,1], None, None)   (utils.py:171-196)."""
        self._check()
        assert len(x.shape) == 3
        if self.ssi:
            # utils.py:172-180: low/up = the 0- and 1-quantiles (min / max) of the valid pixels, unless given
            m = None if mask is None else mask.to(torch.float32)
            if low is None and up is None:
                lu = ops.masked_minmax(x.reshape(1, -1), None if m is None else m.reshape(1, -1))
            else:
                lu = torch.stack([torch.as_tensor(low, dtype=torch.float32, device=x.device).reshape(()),
                                  torch.as_tensor(up, dtype=torch.float32, device=x.device).reshape(())]).view(1, 2)
            y = ops.normalize_ssi(x.reshape(1, -1), None if m is None else m.reshape(1, -1), lu, float(self.t),
                                  float(self.s)).view_as(x)
            return y, lu[0, 0], lu[0, 1]
        if self.mode != "average":
            y = ops.normalize_piecewise(x, self.ch_bounds, self.ch_gammas, float(self.t), float(self.s))
        else:
            y = ops.normalize_disp(x, self.num_chs, float(self.ch_bounds[0]), float(self.ch_gammas[0]), float(self.t),
                                   float(self.s))
        return y.view(x.shape[0] * self.num_chs, *x.shape[1:]), None, None

    def denormalize(self, y: torch.Tensor, raw_disp=None, mask=None):
        """[B, R*num_chs, H, W] -> [B, R, H, W] disparity in pixels (utils.py:231-273)."""
        self._check()
        assert len(y.shape) == 4, "should be of shape B,C,H,W"
        if self.ssi:
            # utils.py:232-251: per (sample, round) least-squares scale/shift against the raw disparity
            assert raw_disp is not None and mask is not None
            B, R, H, W = y.shape
            pred = y.reshape(B * R, H * W)
            gt = raw_disp.to(torch.float32).repeat(1, R, 1, 1).reshape(B * R, H * W)
            mk = mask.to(torch.float32).repeat(1, R, 1, 1).reshape(B * R, H * W)
            if self.safe_ssi:
                from .metrics import RANSAC                       # utils.py:240-243
                st = RANSAC(n=0.1, k=10, d=0.2, t=self.ransac_error_threshold, perm_fn=self.perm_fn).fit(
                    pred, gt, mk).best_fit
            else:
                st = ops.scale_shift_lsq(pred, gt, mk)
            return ops.affine_rows(pred, st).view(B, R, H, W)
        if self.mode != "average":
            return ops.denormalize_piecewise(y, self.ch_bounds, self.ch_gammas, float(self.t), float(self.s))
        return ops.denormalize_disp(y, self.num_chs, float(self.ch_bounds[0]), float(self.ch_gammas[0]),
                                    float(self.t), float(self.s))


class InputPadder:
    """Pads images such that dimensions are divisible by ``divis_by`` (utils.py:297-325; replicate, 'sintel')."""

    def __init__(self, dims: Sequence[int], mode: str = 'sintel', divis_by: int = 8):
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // divis_by) + 1) * divis_by - self.ht) % divis_by
        pad_wd = (((self.wd // divis_by) + 1) * divis_by - self.wd) % divis_by
        if mode == 'sintel':
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    @property
    def padded_size(self):
        return self._pad[2] + self._pad[3] + self.ht, self._pad[0] + self._pad[1] + self.wd

    def pad(self, *inputs):
        assert all((x.ndim == 4) for x in inputs if x is not None)
        return [ops.pad_replicate(x, self._pad) if x is not None else x for x in inputs]

    def pad_zero(self, *inputs):
        """utils.py:317-319 (the training mask is padded with zeros, train.py:186)."""
        assert all((x.ndim == 4) for x in inputs if x is not None)
        return [ops.pad_zero(x, self._pad) if x is not None else x for x in inputs]

    def unpad(self, x: torch.Tensor):
        assert x.ndim == 4, "B,C,H,W"
        ht, wd = x.shape[-2:]
        return ops.crop(x, self._pad[2], self._pad[0], ht - self._pad[2] - self._pad[3],
                        wd - self._pad[0] - self._pad[1])

"""``DepthCamera`` — the part of the reference's camera model the input side of the path needs on the device
(/root/reference/utils_d3roma/camera.py): intrinsics ``K`` (:12-38), ``K_depth`` / ``K_color`` / ``T_cl`` (:98-140),
and the depth -> colour-frame reprojection ``transform_depth_to_rgb_frame`` (:654-701) /
``transform_cropped_depth_to_rgb_frame`` (:610-652) that the robot path runs on every raw depth frame before
inference (d3roma_processor.py:248-286, inference.py:291).  The reprojection is a scatter with numpy's collision order
reproduced exactly on the GPU (``d3_depth_to_rgb_frame``); the 3x3 / 4x4 matrix preparation stays in numpy float64 like
the reference's.  Calibration tables, plotting and open3d helpers of the reference class are out of scope."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np
import torch

from . import _lib


class K:
    """camera.py:12-38."""

    def __init__(self, fx, fy, cx, cy):
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy

    @property
    def vec(self) -> np.ndarray:
        return np.array([self.fx, self.fy, self.cx, self.cy])

    @property
    def arr(self) -> np.ndarray:
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]])

    @property
    def inv(self) -> np.ndarray:
        return np.linalg.inv(self.arr)

    def __mul__(self, scale):
        return K(self.fx * scale, self.fy * scale, self.cx * scale, self.cy * scale)


class DepthCamera:
    """Intrinsics / extrinsics holder with the reference's attribute names (camera.py:40-140)."""

    def __init__(self, resolution: str, scale: float = 1, intrinsics=None, extrinsics=None, config=None,
                 device: str = "unknown", T_fc=None):
        self.scale = scale
        self.resolution_str = resolution            # "WxH"
        self.intrinsics = dict(intrinsics or {})    # "color", "depth": (fx, fy, cx, cy)
        self.extrinsics = dict(extrinsics or {})    # "color_to_depth": 4x4
        self._T_fc = np.eye(4) if T_fc is None else np.asarray(T_fc, dtype=np.float64)
        self.config = dict(config or {"auto_scale": True, "aggressive_fill": True})
        self.device = device

    @property
    def resolution(self):
        W, H = self.resolution_str.split("x")
        return int(H), int(W)

    H = property(lambda self: self.resolution[0])
    W = property(lambda self: self.resolution[1])

    @property
    def K_color(self) -> K:
        return K(*self.intrinsics["color"]) * self.scale

    @property
    def K_depth(self) -> K:
        return K(*self.intrinsics["depth"]) * self.scale

    @property
    def T_cl(self) -> np.ndarray:
        return self._T_fc @ np.asarray(self.extrinsics["color_to_depth"], dtype=np.float64) @ self._T_fc.T

    # ------------------------------------------------------------------ reprojection
    def _reproject(self, depth, Kinv: np.ndarray, canvas, aggressive: bool, crop):
        was_np = isinstance(depth, np.ndarray)
        d = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).cuda() if was_np else depth
        if not (d.is_cuda and d.dtype == torch.float32 and d.dim() == 2):
            raise ValueError("depth must be a float32 [H, W] map (numpy array or CUDA tensor; no CPU fallback)")
        d = d.contiguous()
        H, W = d.shape
        CH, CW = canvas
        oy, ox, oh, ow = crop
        lib = _lib.lib()
        ws = torch.empty((lib.d3_depth_to_rgb_frame_workspace_bytes(CH, CW),), dtype=torch.uint8, device=d.device)
        out = torch.empty((oh, ow), dtype=torch.float32, device=d.device)
        f64 = lambda m: np.ascontiguousarray(m, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        kinv, T, kc = (np.ascontiguousarray(m, dtype=np.float64) for m in (Kinv, self.T_cl[:3, :4], self.K_color.arr))
        _lib.check(lib.d3_depth_to_rgb_frame(d.data_ptr(), H, W, f64(kinv), f64(T), f64(kc), CH, CW, int(aggressive), oy,
                                             ox, oh, ow, ws.data_ptr(), out.data_ptr(),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out.cpu().numpy() if was_np else out

    def transform_depth_to_rgb_frame(self, depth: Union[np.ndarray, torch.Tensor]):
        """camera.py:654-701: depth [H, W] in metres (sensor frame) -> the same map seen from the colour camera."""
        if np.allclose(self.T_cl, np.eye(4)) or self.device == "d435":   # already aligned (:655-657)
            return depth
        H, W = self.H, self.W
        if self.config.get("auto_scale", False) and tuple(depth.shape[:2]) != (H, W):
            import cv2
            if not isinstance(depth, np.ndarray):
                raise ValueError("auto_scale resize (cv2.INTER_NEAREST, :660-662) needs a numpy depth map")
            depth = cv2.resize(depth, (W, H), interpolation=cv2.INTER_NEAREST)
        assert H == depth.shape[0] and W == depth.shape[1], "depth resolution mismatch"
        return self._reproject(depth, self.K_depth.inv, (H, W), bool(self.config.get("aggressive_fill", False)),
                               (0, 0, H, W))

    def transform_cropped_depth_to_rgb_frame(self, depth: Union[np.ndarray, torch.Tensor], cropped_K: np.ndarray):
        """camera.py:610-652: a cropped depth map with its own intrinsics; single rounded fill, result cropped back."""
        if np.allclose(self.T_cl, np.eye(4)):
            return depth
        H, W = depth.shape[:2]
        off_x = int(self.K.cx - cropped_K[0, 2])
        off_y = int(self.K.cy - cropped_K[1, 2])
        return self._reproject(depth, np.linalg.inv(cropped_K), (self.H, self.W), False, (off_y, off_x, H, W))

    @property
    def K(self) -> K:
        """``Realsense.K`` is the COLOUR intrinsics ("caution", camera.py:574-577); the cropped variant derives its crop
        offsets from it (:616-617)."""
        return self.K_color

"""Single-frame inference front end with the reference's call surface (``D3RoMa.infer`` / ``infer_with_rgb_raw`` /
``run_pipeline``, /root/reference/inference.py:97-250): numpy camera frames in, restored depth in metres out.

Everything between the numpy arrays and the pipeline call — cv2.resize, HWC->CHW, ``(x/255-0.5)*2``, depth -> disparity,
``Normalizer.normalize``, the validity mask, the replicate padding the pixel-space UNet needs — is two kernels
(``d3_frame_to_input``, ``d3_depth_to_input``); everything after it (unpad, denormalise, mean over rounds, disparity ->
depth) stays on the device until the final copy.  Construction: pass a ready pipeline, camera and config namespace, or
``D3RoMa.from_config(config, camera)`` which loads the pipeline from ``config.resume_pretrained`` like the reference's
``_load_pipeline`` (inference.py:66-95; no hydra); the public methods (``infer``, ``infer_with_rgb_raw``) are the
reference's.  The internal ``run_pipeline`` takes the RAW frames and depth (the reference's takes tensors its callers
preprocessed, inference.py:191) because resize / normalise / pad are fused into the two kernels above.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib, metrics as M, ops, prepost
from .evaluate import default_config
from .prepost import Camera, InputPadder, Normalizer


def _stream():
    import ctypes as C
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def frames_to_input(frames: np.ndarray | torch.Tensor, size_hw, padder: InputPadder, device) -> torch.Tensor:
    """uint8 [H,W], [H,W,C] or [B,H,W,C] frames -> fp32 [B,3,Hp,Wp] in [-1,1] (resized to ``size_hw``, padded)."""
    t = torch.as_tensor(frames)
    if t.dtype != torch.uint8:
        raise AssertionError("frames must be uint8 (inference.py:108,150)")
    if t.dim() == 2:
        t = t[None, :, :, None]
    elif t.dim() == 3:
        t = t[None] if t.shape[-1] in (1, 3, 4) else t[..., None]    # [H,W,C] frame, else a [B,H,W] stack of gray frames
    t = t.contiguous().to(device, non_blocking=True)
    B, Hs, Ws, C = t.shape
    H, W = size_hw
    pl, pr, pt, pb = padder._pad
    Hp, Wp = H + pt + pb, W + pl + pr
    out = torch.empty((B, 3, Hp, Wp), dtype=torch.float32, device=device)
    _lib.check(_lib.lib().d3_frame_to_input(t.data_ptr(), B, Hs, Ws, C, H, W, pt, pl, Hp, Wp, out.data_ptr(), _stream()))
    return out


def depth_to_input(depth: np.ndarray | torch.Tensor, size_hw, padder: InputPadder, fxb: float, normalizer: Normalizer,
                   device):
    """fp32 depth in metres [H,W] / [H,W,1] / [B,H,W] -> (raw_disp [B,1,H,W], mask [B,1,H,W], normalised padded
    conditioning [B,1,Hp,Wp])."""
    t = torch.as_tensor(depth)
    if t.dim() == 3 and t.shape[-1] in (1, 3):      # [H,W,1] or a 3-channel depth image (inference.py:176-179)
        t = t[..., 0]
    if t.dim() == 2:
        t = t[None]
    t = t.to(torch.float32).contiguous().to(device, non_blocking=True)
    B, Hs, Ws = t.shape
    H, W = size_hw
    pl, pr, pt, pb = padder._pad
    Hp, Wp = H + pt + pb, W + pl + pr
    if normalizer.ssi or normalizer.mode != "average" or normalizer.num_chs != 1:
        raise NotImplementedError("depth_to_input covers the non-SSI single-channel 'average' normaliser")
    raw = torch.empty((B, 1, H, W), dtype=torch.float32, device=device)
    mask = torch.empty_like(raw)
    norm = torch.empty((B, 1, Hp, Wp), dtype=torch.float32, device=device)
    _lib.check(_lib.lib().d3_depth_to_input(t.data_ptr(), B, Hs, Ws, H, W, pt, pl, Hp, Wp, float(fxb),
                                            float(normalizer.ch_bounds[0]), float(normalizer.ch_gammas[0]),
                                            float(normalizer.t), float(normalizer.s), raw.data_ptr(), mask.data_ptr(),
                                            norm.data_ptr(), _stream()))
    return raw, mask, norm


class D3RoMa:
    """``D3RoMa`` of /root/reference/inference.py:28-250 over a ready pipeline.

    ``camera.resolution`` = (H, W) the frames are resized to (``self.camera.resolution[::-1]`` is cv2's (W, H) dsize,
    inference.py:121,165); ``camera.fxb`` plays ``camera.fxb_depth``.  ``config``: the fields of
    ``evaluate.default_config``."""

    def __init__(self, pipeline, camera: Camera, config=None, resolution_hw=None):
        self.pipeline = pipeline
        self.camera = camera
        self.config = config if config is not None else default_config()
        self.normer = Normalizer.from_config(self.config)
        if resolution_hw is None:
            w, h = (int(v) for v in self.config.camera_resolution.split("x"))
            resolution_hw = (h, w)
        self.resolution = tuple(resolution_hw)
        self.last_metrics = None

    # ---- inference.py:66-95
    @staticmethod
    def load_pipeline(config, compute_dtype=None, device="cuda"):
        """``D3RoMa._load_pipeline`` (inference.py:66-95): the pipeline class picked by ``config.ldm``, loaded from the
        diffusers-layout directory ``config.resume_pretrained``, ``flow_guidance_mode`` taken from the config, the
        sampler replaced by ``MyDDIMScheduler`` when ``config.sampler == "my_ddim"`` (same scheduler fields)."""
        import os
        from .latent_pipeline import GuidedLatentDiffusionPipeline
        from .pipelines import GuidedDiffusionPipeline
        from .scheduler_ddim import MyDDIMScheduler
        path = f"{config.resume_pretrained}"
        if not os.path.exists(path):
            raise ValueError(f"patrained path not exists: {path}")          # (sic) inference.py:93
        clazz = GuidedLatentDiffusionPipeline if getattr(config, "ldm", False) else GuidedDiffusionPipeline
        pipeline = clazz.from_pretrained(path, compute_dtype=compute_dtype, device=device).to(device)
        pipeline.guidance.flow_guidance_mode = config.flow_guidance_mode
        if getattr(config, "sampler", None) == "my_ddim":
            pipeline.scheduler = MyDDIMScheduler.from_config(dict(
                beta_schedule=config.beta_schedule, beta_start=config.beta_start, beta_end=config.beta_end,
                clip_sample=config.clip_sample, num_train_timesteps=config.num_train_timesteps,
                prediction_type=config.prediction_type, set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1,
                trained_betas=None))
        return pipeline

    @classmethod
    def from_config(cls, config, camera: Camera, compute_dtype=None, device="cuda"):
        """The reference's constructor path (inference.py:40-64) without hydra: ``config`` carries the fields of the
        task YAML (``resume_pretrained``, ``ldm``, ``sampler``, scheduler and normaliser fields)."""
        return cls(cls.load_pipeline(config, compute_dtype, device), camera, config)

    # ---- inference.py:97-135
    @torch.no_grad()
    def infer_with_rgb_raw(self, rgb: np.ndarray, raw_depth: np.ndarray):
        """Depth restoration with RGB and raw depth (aligned).  Returns the restored depth in metres (numpy [H,W])."""
        assert rgb.dtype == np.uint8
        return self.run_pipeline(None, None, raw_depth, rgb)

    # ---- inference.py:137-189
    @torch.no_grad()
    def infer(self, left: np.ndarray, right: np.ndarray, raw_depth: np.ndarray = None, rgb: np.ndarray = None):
        """Depth restoration with left / right (IR) images and raw depth.  Returns depth in metres (numpy [H,W])."""
        assert len(left.shape) == len(right.shape)
        assert left.dtype == right.dtype == np.uint8
        if raw_depth is None or rgb is None:
            raise NotImplementedError("no worry, i will implement this soon")   # inference.py:153-154
        return self.run_pipeline(left, right, raw_depth, rgb)

    # ---- inference.py:191-250
    @torch.no_grad()
    def run_pipeline(self, left_image, right_image, raw_depth, rgb):
        cfg = self.config
        dev = self.pipeline.device
        H, W = self.resolution
        R = int(cfg.num_inference_rounds)
        padder = InputPadder((1, 1, H, W), divis_by=cfg.divis_by)
        rep = lambda t: None if t is None else t.repeat(R, 1, 1, 1)   # noqa: E731  batch = inference rounds (:198-213)
        n_rgb = None if rgb is None else frames_to_input(rgb, (H, W), padder, dev)
        n_left = None if left_image is None else frames_to_input(left_image, (H, W), padder, dev)
        n_right = None if right_image is None else frames_to_input(right_image, (H, W), padder, dev)
        raw_disp, mask, n_raw = depth_to_input(raw_depth, (H, W), padder, self.camera.fxb, self.normer, dev)
        out = self.pipeline(rep(n_rgb), rep(n_left), rep(n_right), rep(n_raw), rep(raw_disp), rep(mask),
                            num_inference_steps=cfg.num_inference_timesteps,
                            num_intermediate_images=cfg.num_intermediate_images, add_noise_rgb=cfg.noise_rgb,
                            depth_channels=cfg.depth_channels, cond_channels=cfg.cond_channels,
                            denorm=self.normer.denormalize)                      # inference.py:214-223
        pred = padder.unpad(out.images)                                          # [R, dc, H, W]
        disp = self.normer.denormalize(pred, rep(raw_disp), rep(mask))           # inference.py:232
        disp = disp.mean(dim=0, keepdim=True).mean(dim=1)                        # rounds (and dc replicas) -> [1, H, W]
        self.uncertainty = pred.std(dim=0, unbiased=False)[0] * mask[0, 0] if R > 1 else None   # :226-230
        fxb = torch.tensor([self.camera.fxb], dtype=torch.float32, device=dev)
        rows = M.compute_errors_device(raw_disp[:, 0], disp, mask[:, 0] > 0, fxb)   # vs the RAW disparity (:236-243)
        self.last_metrics = M.metrics_to_dict(rows[:, :5].cpu().numpy(), rows[:, 5:].cpu().numpy())
        depth = prepost.disp_to_depth(disp, self.camera.fxb)                     # inference.py:246-249
        return depth[0].cpu().numpy()

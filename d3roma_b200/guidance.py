"""``FlowGuidance`` — the guidance hook the schedulers call (``guidance.perturb`` gated on
``guidance.flow_guidance_weight > 0``; /root/reference/core/scheduler_ddim.py:461, core/scheduler_ddpm.py:506).

Reference behaviour (/root/reference/core/guidance.py): the constructor signature (:95-107) and ``perturb`` signature
(:214) are kept. Of the reference's modes, ``"imputation"`` raises ``NotImplementedError`` (:218-219) and
``"gradient"`` is a 10-iteration Adam loop through the *latent* VAE decoder (:224-240) that belongs to the LDM family
(out of scope).  The stereo photometric gradient (``grad``, :293-331) is computed by the reference but
never wired into the pixel pipeline; here it is exposed as mode ``"stereo"``:

    x0' = x0 - flow_guidance_weight * d(loss)/d(x0)

where ``loss`` is the reference's own photometric loss (guidance.py:318-321) evaluated on the disparity the sample
stands for.  ``x0`` lives in the Normalizer's space (utils_d3roma/utils.py:183-196: ``(disp/bound - t) * s``), the
reference's loss takes CAMERA-normalised disparity (``camera.unnormalize_disp``, utils_d3roma/camera.py:599-602,
771-780), so the chain is

    x0 --denormer--> disparity [px] --camera.normalize_disp--> d_hat --grad() of the reference--> d(loss)/d(d_hat)
    d(loss)/d(x0_c) = d(loss)/d(d_hat) * d(d_hat)/d(px) * d(px)/d(x0_c) = G / (max_d - min_d) * bound / (s * C)

``denormer`` is the reference's hook argument for exactly this (``perturb(..., denormer, ...)``, guidance.py:214; the
pipeline's ``denorm=`` argument, custom_pipelines.py:104).  The whole chain, the gradient, the update and the
scheduler's recombination run in ONE kernel (``d3_guidance_step``).  ``perturb_start_ratio`` gates the update to the
last part of the trajectory.
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import torch

from . import ops
from .prepost import SIM_CAMERA, Normalizer


class FlowGuidance:
    def __init__(self, flow_guidance_weight=1., perturb_start_ratio=0., flow_guidance_mode="imputation", camera=None):
        self.flow_guidance_weight = flow_guidance_weight
        self.perturb_start_ratio = perturb_start_ratio
        self.flow_guidance_mode = flow_guidance_mode
        self.no_ssim = False
        self.camera = camera if camera is not None else SIM_CAMERA
        self.num_train_timesteps: Optional[int] = None  # set by the pipeline for perturb_start_ratio gating
        self.denormer = None        # default x0 -> pixels map when the scheduler passes none (the DDPM hook, see perturb)
        self._affine_cache = {}

    # ModelMixin-ish surface used by callers (`pipeline.to`, save/load of an empty module)
    def to(self, *a, **k):
        return self

    def parameters(self):
        return iter(())

    @property
    def _w_ssim(self) -> float:
        return 0.0 if self.no_ssim else 0.95

    def grad(self, disp_raw: torch.Tensor, left_images: torch.Tensor, right_images: torch.Tensor, timestep=None,
             th=None) -> torch.Tensor:
        """d(photometric loss)/d(camera-normalised disparity), [B,1,H,W] — guidance.py:298-323 as one CUDA kernel."""
        cam = self.camera
        return ops.guidance_grad(disp_raw, left_images, right_images, float(cam.min_disp), float(cam.max_disp),
                                 float(cam.shift), self._w_ssim)

    # ------------------------------------------------------------------ x0 -> pixels
    def denormer_affine(self, denormer, C: int, device=None) -> Tuple[float, float]:
        """``(scale, offset)`` such that ``denormer(x0) = offset + scale * sum_c x0_c`` (pixels).  Accepts a
        ``Normalizer``, its bound ``denormalize`` method, ``functools.partial(denormalize, config)`` as the reference's
        callers build it (inference.py:214), or any affine callable (probed once with three constant tensors)."""
        if denormer is None:
            denormer = self.denormer
        if denormer is None:
            raise ValueError("stereo guidance needs the x0 -> disparity map: pass denorm= to the pipeline / denormer= "
                             "to scheduler.step (the reference's hook argument), or set guidance.denormer")
        n = None
        if isinstance(denormer, Normalizer):
            n = denormer
        elif isinstance(getattr(denormer, "__self__", None), Normalizer):
            n = denormer.__self__
        elif isinstance(denormer, functools.partial):
            for arg in tuple(denormer.args) + tuple(denormer.keywords.values()):
                if all(hasattr(arg, k) for k in ("ssi", "normalize_mode", "num_chs", "ch_bounds", "ch_gammas")):
                    n = Normalizer.from_config(arg)
                    break
        if n is not None:
            if n.ssi:
                raise NotImplementedError("stereo guidance with a scale-shift-invariant normaliser: the x0 -> pixel map "
                                          "depends on the prediction itself")
            n._check()
            if float(n.ch_gammas[0]) != 1.0:
                raise NotImplementedError("stereo guidance needs gamma = 1 (an affine x0 -> pixel map)")
            if C != n.num_chs:
                raise ValueError(f"x0 has {C} channels, the normaliser {n.num_chs}")
            bound = float(n.ch_bounds[0])
            return bound / (float(n.s) * C), float(n.t) * bound
        key = id(denormer)
        hit = self._affine_cache.get(key)
        if hit is not None and hit[0] is denormer and hit[1] == C:
            return hit[2]
        dev = device if device is not None else "cuda"
        probe = torch.stack([torch.full((C, 2, 2), v, dtype=torch.float32, device=dev) for v in (0.0, 1.0, -1.0)])
        d = torch.as_tensor(denormer(probe)).to(torch.float32).reshape(3, -1).mean(dim=1).cpu()
        d0, d1, d2 = (float(v) for v in d)
        if abs((d1 - d0) - (d0 - d2)) > 1e-4 * max(1.0, abs(d1 - d2)):
            raise ValueError("stereo guidance needs an affine denormer (x0 -> pixels)")
        res = ((d1 - d0) / C, d0)
        self._affine_cache[key] = (denormer, C, res)
        return res

    def _gate(self, t) -> bool:
        """False while the trajectory is still before ``perturb_start_ratio``."""
        if self.num_train_timesteps is not None and self.perturb_start_ratio > 0.0:
            return int(t) <= (1.0 - self.perturb_start_ratio) * self.num_train_timesteps
        return True

    def _check_mode(self, left_image, right_image):
        if self.flow_guidance_mode == "imputation":
            raise NotImplementedError
        if self.flow_guidance_mode == "gradient":
            raise NotImplementedError("'gradient' mode optimises VAE latents (guidance.py:224-240): LDM family only")
        if self.flow_guidance_mode != "stereo":
            raise NotImplementedError
        if left_image is None or right_image is None:
            raise ValueError("stereo guidance needs left_image and right_image")

    def perturb(self, pred_original_sample, t, decoder, denormer, raw_mask=None, left_image=None, right_image=None,
                raw_depth=None):
        self._check_mode(left_image, right_image)
        if not self._gate(t):
            return pred_original_sample
        x0 = pred_original_sample
        sc, off = self.denormer_affine(denormer, x0.shape[1], x0.device)
        cam = self.camera
        x0p, _, _ = ops.guidance_step(x0, left_image, right_image, sc, off, float(cam.min_disp), float(cam.max_disp),
                                      self._w_ssim, step=float(self.flow_guidance_weight), want_x0=True)
        return x0p

    def perturb_fused(self, pred_original_sample, t, decoder, denormer, ca: float, aux1: torch.Tensor, cb: float,
                      raw_mask=None, left_image=None, right_image=None, raw_depth=None):
        """``perturb`` plus the scheduler's recombination ``ca * x0' + cb * aux1`` (scheduler_ddim.py:468,
        scheduler_ddpm.py:512) in the same launch.  Returns ``(x0', recombined)``."""
        self._check_mode(left_image, right_image)
        x0 = pred_original_sample
        if not self._gate(t):
            return x0, ops.lincomb(ca, x0, cb, aux1)
        sc, off = self.denormer_affine(denormer, x0.shape[1], x0.device)
        cam = self.camera
        x0p, out2, _ = ops.guidance_step(x0, left_image, right_image, sc, off, float(cam.min_disp),
                                         float(cam.max_disp), self._w_ssim, step=float(self.flow_guidance_weight),
                                         want_x0=True, ca=ca, aux1=aux1, cb=cb)
        return x0p, out2

    def loss_grad_x0(self, x0, denormer, left_image, right_image) -> torch.Tensor:
        """d(loss)/d(x0_c) [B,1,H,W] (the same for every channel) — the quantity ``perturb`` steps along."""
        sc, off = self.denormer_affine(denormer, x0.shape[1], x0.device)
        cam = self.camera
        _, _, g = ops.guidance_step(x0, left_image, right_image, sc, off, float(cam.min_disp), float(cam.max_disp),
                                    self._w_ssim, want_grad=True)
        return g

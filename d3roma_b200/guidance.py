"""``FlowGuidance`` — the guidance hook the schedulers call (``guidance.perturb`` gated on
``guidance.flow_guidance_weight > 0``; /root/reference/core/scheduler_ddim.py:461, core/scheduler_ddpm.py:506).

Reference behaviour (/root/reference/core/guidance.py): the constructor signature (:95-107) and ``perturb`` signature
(:214) are kept. Of the reference's modes, ``"imputation"`` raises ``NotImplementedError`` (:218-219) and
``"gradient"`` is a 10-iteration Adam loop through the *latent* VAE decoder (:224-240) that belongs to the LDM family
(out of scope).  The stereo photometric gradient (``grad``, :293-331) is computed by the reference but
never wired into the pixel pipeline; here it is exposed as mode ``"stereo"``:

    x0' = x0 - flow_guidance_weight * d(loss)/d(x0)        (loss of guidance.py:318-321, gradient by d3_guidance_grad)

applied from ``perturb_start_ratio`` of the way through sampling on, which is what the hook and the weight are for.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .prepost import SIM_CAMERA


class FlowGuidance:
    def __init__(self, flow_guidance_weight=1., perturb_start_ratio=0., flow_guidance_mode="imputation", camera=None):
        self.flow_guidance_weight = flow_guidance_weight
        self.perturb_start_ratio = perturb_start_ratio
        self.flow_guidance_mode = flow_guidance_mode
        self.no_ssim = False
        self.camera = camera if camera is not None else SIM_CAMERA
        self.num_train_timesteps: Optional[int] = None  # set by the pipeline for perturb_start_ratio gating

    # ModelMixin-ish surface used by callers (`pipeline.to`, save/load of an empty module)
    def to(self, *a, **k):
        return self

    def parameters(self):
        return iter(())

    def grad(self, disp_raw: torch.Tensor, left_images: torch.Tensor, right_images: torch.Tensor, timestep=None,
             th=None) -> torch.Tensor:
        """d(photometric loss)/d(normalised disparity), [B,1,H,W] — guidance.py:298-323 as one CUDA kernel."""
        cam = self.camera
        return ops.guidance_grad(disp_raw, left_images, right_images, float(cam.min_disp), float(cam.max_disp),
                                 float(cam.shift), 0.0 if self.no_ssim else 0.95)

    def perturb(self, pred_original_sample, t, decoder, denormer, raw_mask=None, left_image=None, right_image=None,
                raw_depth=None):
        if self.flow_guidance_mode == "imputation":
            raise NotImplementedError
        if self.flow_guidance_mode == "gradient":
            raise NotImplementedError("'gradient' mode optimises VAE latents (guidance.py:224-240): LDM family only")
        if self.flow_guidance_mode != "stereo":
            raise NotImplementedError
        if left_image is None or right_image is None:
            raise ValueError("stereo guidance needs left_image and right_image")
        if self.num_train_timesteps is not None and self.perturb_start_ratio > 0.0:
            if int(t) > (1.0 - self.perturb_start_ratio) * self.num_train_timesteps:
                return pred_original_sample
        x0 = pred_original_sample
        if x0.shape[1] != 1:
            # SceneFlow variant: the target is the same disparity replicated over 3 channels (utils.py:196)
            g = self.grad(x0.mean(dim=1, keepdim=True), left_image, right_image, t)
            return ops.lincomb(1.0, x0, -float(self.flow_guidance_weight) / x0.shape[1],
                               g.expand_as(x0).contiguous())
        g = self.grad(x0, left_image, right_image, t)
        return ops.lincomb(1.0, x0, -float(self.flow_guidance_weight), g)

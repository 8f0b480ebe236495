"""ORACLE (test infrastructure, not product code): CPU restatement of the GUIDED pixel-space sampling loop.

The reference computes the stereo photometric loss (``FlowGuidance.grad``, /root/reference/core/guidance.py:298-323,
``SSIM`` :28-42) and gives both schedulers a guidance hook (core/scheduler_ddim.py:461-472,
core/scheduler_ddpm.py:506-516) whose ``denormer`` argument maps the sample to pixels, but never connects the two in
the pixel pipeline (``custom_pipelines.py:174``).  The connection built here and in
``d3roma_b200.FlowGuidance`` is the chain rule over the reference's own maps:

    x0 --Normalizer.denormalize (utils_d3roma/utils.py:260-273)--> disparity [px]
       --camera.normalize_disp (utils_d3roma/camera.py:771-776)--> d_hat
       --camera.unnormalize_disp + clip (:599-602) / grid_sample / SSIM + L1 (guidance.py:298-321)--> loss
    x0' = x0 - flow_guidance_weight * d(loss)/d(x0)            (autograd, exactly as the reference differentiates)

followed by the hook's algebra ``prev' = sqrt(a_prev) * x0' + direction`` (DDIM) / ``c0 * x0' + ct * x_t`` (DDPM).

Pinned against ``tests/golden/guided.npz`` (the reference's scheduler classes, ``SSIM`` module, ``Normalizer`` and
camera functions driving this loop; oracle/gen_golden.py ``gen_guided``) in tests/test_oracle_guided.py.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from . import schedulers_ref as S
from .guidance_ref import ssim_loss
from .pipeline_ref import COND_ORDER


def photometric_loss(disp_px: torch.Tensor, left: torch.Tensor, right: torch.Tensor, min_disp: float, max_disp: float,
                     w: float = 0.95) -> torch.Tensor:
    """guidance.py:298-321 for a disparity map in PIXELS [B,H,W] (clipped to the camera range like
    ``camera.unnormalize_disp``): warp right -> left, 0.95 * sum SSIM_loss + 0.05 * sum L1."""
    B, H, W = disp_px.shape
    disp = disp_px.clip(min_disp, max_disp)
    margin = right.shape[-1] - W
    xx, yy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    xx = xx[None].repeat(B, 1, 1)
    yy = yy[None].repeat(B, 1, 1)
    gx = (xx - disp + margin) / ((W + margin - 1) / 2.0) - 1
    gy = yy / ((H - 1) / 2.0) - 1
    warp = F.grid_sample(right, torch.stack([gx, gy], dim=-1), align_corners=True, mode="bilinear",
                         padding_mode="border")
    lcrop = left[..., margin:]
    l1 = F.l1_loss(lcrop, warp, reduction="sum")
    return w * ssim_loss(lcrop, warp).sum() + (1 - w) * l1


def denormalize_average(y: torch.Tensor, bound: float, t: float = 0.5, s: float = 2.0) -> torch.Tensor:
    """utils.py:260-273, mode "average", gamma 1, one round: [B,C,H,W] -> [B,H,W] pixels."""
    C = y.shape[1]
    return ((y / s + t) * (bound / C)).sum(dim=1)


def x0_loss_and_grad(x0: torch.Tensor, left, right, bound: float, min_disp: float, max_disp: float, shift: float = 0.0,
                     w: float = 0.95, t: float = 0.5, s: float = 2.0):
    """loss(x0) through denormalise -> camera-normalise -> unnormalise + clip -> warp, and d(loss)/d(x0) [B,C,H,W]."""
    x = x0.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        px = denormalize_average(x, bound, t, s)
        d_hat = (px - min_disp) / (max_disp - min_disp) + shift                 # camera.py:771-776
        disp = min_disp + (d_hat - shift) * (max_disp - min_disp)                # camera.py:778-780
        loss = photometric_loss(disp, left, right, min_disp, max_disp, w)
        (g,) = torch.autograd.grad(loss, x)
    return loss.detach(), g


def run_guided_pipeline(unet, x_T: torch.Tensor, conds: dict, cond_channels: str, left, right, *, weight: float,
                        bound: float, min_disp: float, max_disp: float, shift: float = 0.0, sampler: str = "ddim",
                        num_inference_steps: int = 10, num_train_timesteps: int = 128,
                        beta_schedule: str = "squaredcos_cap_v2", prediction_type: str = "sample",
                        steps_offset: int = 1, ddpm_noise=None, record: Optional[List[dict]] = None,
                        teacher_outputs: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """custom_pipelines.py:134-196 with the guidance kwargs of :174 enabled.  ``record`` receives per-step dicts
    (model_output, sample, x0, x0_perturbed, prev); ``teacher_outputs`` replaces the UNet by recorded outputs."""
    T, n = num_train_timesteps, num_inference_steps
    ac = torch.cumprod(1.0 - S.make_betas(beta_schedule, T), dim=0)
    off = steps_offset if sampler == "ddim" else 0
    timesteps = S.make_timesteps(T, n, "leading", off)
    cond = torch.cat([conds[k] for k in COND_ORDER[cond_channels]], dim=1)
    image = x_T
    for i, t in enumerate(timesteps):
        if teacher_outputs is not None:
            m = teacher_outputs[i]
        else:
            with torch.no_grad():
                m = unet(torch.cat([image, cond], dim=1), torch.tensor(int(t))).sample
        rec = {"model_output": m.clone(), "sample": image.clone()}

        def perturb(x0):
            _, g = x0_loss_and_grad(x0, left, right, bound, min_disp, max_disp, shift)
            rec["x0"] = x0.clone()
            rec["x0_perturbed"] = x0 - weight * g
            return rec["x0_perturbed"]

        if sampler == "ddim":
            image = guided_ddim_step(m, int(t), image, ac, ac[0], T, n, perturb, prediction_type=prediction_type)
        else:
            noise = ddpm_noise(i, m.shape) if int(t) > 0 else None
            image = guided_ddpm_step(m, int(t), image, ac, T, n, noise, perturb, prediction_type=prediction_type)
        rec["prev"] = image.clone()
        if record is not None:
            record.append(rec)
    return image.clamp(-1, 1)


def guided_ddim_step(model_output, t, sample, ac, final_ac, T, n, perturb, prediction_type="sample", clip_sample=True):
    """scheduler_ddim.py:413-472 (eta = 0): the unguided step gives x0 and prev = sqrt(a_prev) x0 + direction; the hook
    replaces x0 by perturb(x0) in the first term only (:468)."""
    prev, x0 = S.ddim_step(model_output, t, sample, ac, final_ac, T, n, prediction_type=prediction_type,
                           clip_sample=clip_sample)
    prev_t = t - T // n
    a_prev = ac[prev_t] if prev_t >= 0 else final_ac
    direction = prev - a_prev ** 0.5 * x0
    return a_prev ** 0.5 * perturb(x0) + direction


def guided_ddpm_step(model_output, t, sample, ac, T, n, noise, perturb, prediction_type="sample"):
    """scheduler_ddpm.py:460-518: mean' = c0 * perturb(x0) + ct * x_t, then + sqrt(var) * noise."""
    prev, x0, mean = S.ddpm_step(model_output, t, sample, ac, T, n, noise, prediction_type=prediction_type,
                                 thresholding=True)
    prev_t = t - T // n
    a_t = ac[t]
    a_prev = ac[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    c0 = (a_prev ** 0.5 * (1 - a_t / a_prev)) / (1 - a_t)
    return prev + c0 * (perturb(x0) - x0)

"""ORACLE (test infrastructure, not product code): the reference's training step for the pixel-space model, restated.

Follows ``train_step`` of ``/root/reference/train.py:157-538`` with ``config.ldm = False``:
  * forward diffusion ``noisy = noise_scheduler.add_noise(x_clean, noise, timesteps)`` (:278)
  * ``inputs = cat([noisy, left, right, sim_disp], 1)`` etc. (:283-297), ``model(inputs, timesteps)`` (:308)
  * ``network_loss = loss_fn(model_output * mask, target * mask)`` for epsilon (:311) / per channel for ``sample``
    (:379-384), plain ``loss_fn(model_output, v)`` for v_prediction (:320); ``loss_fn`` = MSELoss / L1Loss with
    ``reduction='none'`` (:255-260); ``loss = (network_loss * weights[:, None, None, None]).mean()`` (:418)
  * ``accelerator.backward(loss)`` (:455) = autograd; ``clip_grad_norm_(model.parameters(), config.clip_grad_norm)`` (:518);
    ``optimizer.step()`` with ``torch.optim.AdamW(model.parameters(), lr=config.learning_rate)`` (:524, :862).

``accelerate`` (the wrapper around backward / clipping / DDP) is a third-party dependency absent from this image, so the
function itself cannot be imported: **parity unpinned** for the step as a whole.  Its pieces are pinned: the scheduler
arithmetic against the reference's own classes (tests/golden/ddpm.npz), autograd / ``clip_grad_norm_`` / ``AdamW`` ARE
torch's, and the UNet is ``oracle/unet_ref.py``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import schedulers_ref as S


def loss_and_grads(unet: torch.nn.Module, noisy: torch.Tensor, cond: Optional[torch.Tensor], timesteps: torch.Tensor,
                   target: torch.Tensor, mask: Optional[torch.Tensor], weights: Optional[torch.Tensor],
                   loss_type: str = "mse"):
    """Forward + loss + autograd backward; returns (loss, model_output, {name: grad})."""
    unet.zero_grad(set_to_none=True)
    inputs = noisy if cond is None else torch.cat([noisy, cond], dim=1)
    out = unet(inputs, timesteps)
    out = out.sample if hasattr(out, "sample") else out
    fn = F.mse_loss if loss_type == "mse" else F.l1_loss
    m = 1.0 if mask is None else mask
    network_loss = fn(out * m, target * m, reduction="none")
    w = torch.ones(out.shape[0]) if weights is None else weights
    loss = (network_loss * w.view(-1, 1, 1, 1).to(network_loss)).mean()
    loss.backward()
    return loss.detach(), out.detach(), {k: p.grad.detach().clone() for k, p in unet.named_parameters()}


def train_step(unet: torch.nn.Module, optimizer: torch.optim.Optimizer, clean: torch.Tensor, cond: Optional[torch.Tensor],
               noise: torch.Tensor, timesteps: torch.Tensor, alphas_cumprod: torch.Tensor, mask=None, weights=None,
               prediction_type: str = "sample", loss_type: str = "mse", clip_grad_norm: float = 1.0):
    """One optimisation step; returns (loss, grad_norm)."""
    noisy = S.add_noise(clean, noise, timesteps, alphas_cumprod)
    if prediction_type == "sample":
        target = clean
    elif prediction_type == "epsilon":
        target = noise
    elif prediction_type == "v_prediction":
        target, mask = S.get_velocity(clean, noise, timesteps, alphas_cumprod), None
    else:
        raise NotImplementedError(prediction_type)
    loss, _, _ = loss_and_grads(unet, noisy, cond, timesteps, target, mask, weights, loss_type)
    gn = torch.nn.utils.clip_grad_norm_(unet.parameters(), clip_grad_norm)
    optimizer.step()
    optimizer.zero_grad()
    return loss, gn


def train_step_full(config, unet, optimizer, alphas_cumprod, t_sampler, normalized_disp, normalized_rgb=None, mask=None,
                    left_image=None, right_image=None, sim_disp=None):
    """The whole pixel-space ``train_step`` including its front end: out-of-range clamp + mask update (train.py:175-178),
    ``InputPadder.pad`` / ``pad_zero`` (:183-186), ``torch.randn`` noise (:266), ``t_sampler.sample`` (:261), the
    ``cond_channels`` concatenation (:283-297).  Returns (loss, per-sample losses, grad_norm, timesteps)."""
    from . import prepost_ref as P
    if config.clip_sample and (normalized_disp.max() > config.clip_sample_range
                               or normalized_disp.min() < -config.clip_sample_range):
        normalized_disp = torch.clamp(normalized_disp, -config.clip_sample_range, config.clip_sample_range)
        mask = mask * (normalized_disp.abs() < config.clip_sample_range).to(torch.float32)
    pads = P.pad_amounts(normalized_disp.shape[-2], normalized_disp.shape[-1], config.divis_by)
    rep = lambda x: None if x is None else F.pad(x, pads, mode="replicate")   # noqa: E731
    normalized_disp, normalized_rgb, left_image, right_image, sim_disp = map(
        rep, (normalized_disp, normalized_rgb, left_image, right_image, sim_disp))
    mask = F.pad(mask, pads, mode="constant", value=0)
    noise = torch.randn(normalized_disp.shape)
    timesteps, weights = t_sampler.sample(noise.shape[0], noise.device)
    if "rgb" in config.cond_channels and getattr(config, "noise_rgb", False):   # :271-276
        noise_rgb = torch.randn(normalized_rgb.shape)
        normalized_rgb = normalized_rgb * 0.5 + S.add_noise(normalized_rgb, noise_rgb, timesteps, alphas_cumprod) * 0.5
    parts = {"rgb": normalized_rgb, "left": left_image, "right": right_image, "raw": sim_disp}
    order = {"rgb": ("rgb",), "rgb+raw": ("rgb", "raw"), "left+right": ("left", "right"),
             "left+right+raw": ("left", "right", "raw"), "rgb+left+right": ("rgb", "left", "right"),
             "rgb+left+right+raw": ("rgb", "left", "right", "raw")}[config.cond_channels]
    cond = torch.cat([parts[k] for k in order], dim=1)
    noisy = S.add_noise(normalized_disp, noise, timesteps, alphas_cumprod)
    if config.prediction_type == "sample":
        target, m = normalized_disp, mask
    elif config.prediction_type == "epsilon":
        target, m = noise, mask
    else:
        target, m = S.get_velocity(normalized_disp, noise, timesteps, alphas_cumprod), None
    unet.zero_grad(set_to_none=True)
    out = unet(torch.cat([noisy, cond], 1), timesteps)
    out = out.sample if hasattr(out, "sample") else out
    fn = F.mse_loss if config.loss_type == "mse" else F.l1_loss
    mm = 1.0 if m is None else m
    network_loss = fn(out * mm, target * mm, reduction="none")
    loss = (network_loss * weights.view(-1, 1, 1, 1)).mean()
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(unet.parameters(), config.clip_grad_norm)
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach(), network_loss.mean(dim=(1, 2, 3)).detach(), gn, timesteps

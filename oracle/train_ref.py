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

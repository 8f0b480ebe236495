"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's DDIM / DDPM
scheduler arithmetic in plain torch-CPU fp32 (the reference computes with 0-dim fp32 CPU tensors).

Follows /root/reference/core/scheduler_ddim.py (``__init__`` :207-238, ``set_timesteps`` :301-344,
``_get_variance`` :257-265, ``step`` :346-498) and core/scheduler_ddpm.py (``__init__`` :167-206,
``set_timesteps`` :225-298, ``_get_variance`` :300-338, ``previous_timestep`` :388-401, ``step`` :403-526).

Pinned: every function here is checked against ``tests/golden/ddim.npz`` / ``ddpm.npz``, which were
produced by the reference's own classes (``oracle/gen_golden.py``) — see tests/test_oracle_schedulers.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def make_betas(schedule: str, T: int, beta_start: float = 1e-4, beta_end: float = 0.02) -> torch.Tensor:
    """scheduler_ddim.py:207-221 / scheduler_ddpm.py:167-186 (+ betas_for_alpha_bar :53-94)."""
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
    if schedule == "squaredcos_cap_v2":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        return torch.tensor([min(1 - f((i + 1) / T) / f(i / T), 0.999) for i in range(T)], dtype=torch.float32)
    if schedule == "sigmoid":
        return torch.sigmoid(torch.linspace(-6, 6, T)) * (beta_end - beta_start) + beta_start
    raise NotImplementedError(schedule)


def make_timesteps(T: int, n: int, spacing: str = "leading", steps_offset: int = 0) -> np.ndarray:
    """scheduler_ddim.py:320-342 / scheduler_ddpm.py:270-296."""
    if spacing == "linspace":
        return np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
    if spacing == "leading":
        return (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + steps_offset
    if spacing == "trailing":
        return np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
    raise ValueError(spacing)


def _clip(x0, thresholding, clip_sample, clip_range):
    # thresholding with sample_max_value == 1.0 is exactly clamp(-1, 1)
    if thresholding:
        return x0.clamp(-1.0, 1.0) / 1.0
    if clip_sample:
        return x0.clamp(-clip_range, clip_range)
    return x0


def ddim_step(model_output, t: int, sample, alphas_cumprod, final_alpha_cumprod, T: int, n: int,
              prediction_type="epsilon", clip_sample=True, thresholding=False, clip_sample_range=1.0, eta=0.0,
              use_clipped_model_output=False, variance_noise=None):
    """scheduler_ddim.py:413-498 without the guidance hook. Returns (prev_sample, pred_original_sample)."""
    prev_t = t - T // n
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else final_alpha_cumprod
    b_t = 1 - a_t
    if prediction_type == "epsilon":
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        eps = model_output
    elif prediction_type == "sample":
        x0 = model_output
        eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
    elif prediction_type == "v_prediction":
        x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
    else:
        raise ValueError(prediction_type)
    x0 = _clip(x0, thresholding, clip_sample, clip_sample_range)
    variance = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
    std = eta * variance ** 0.5
    if use_clipped_model_output:
        eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
    direction = (1 - a_prev - std ** 2) ** 0.5 * eps
    prev = a_prev ** 0.5 * x0 + direction
    if eta > 0:
        prev = prev + std * variance_noise
    return prev, x0


def ddpm_variance(t: int, prev_t: int, alphas_cumprod, variance_type="fixed_small"):
    """scheduler_ddpm.py:300-338 (fixed_* types)."""
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    cur_beta = 1 - a_t / a_prev
    var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
    if variance_type == "fixed_small":
        return var
    if variance_type == "fixed_small_log":
        return torch.exp(0.5 * torch.log(var))
    if variance_type == "fixed_large":
        return cur_beta
    if variance_type == "fixed_large_log":
        return torch.log(cur_beta)
    raise NotImplementedError(variance_type)


def ddpm_step(model_output, t: int, sample, alphas_cumprod, T: int, n: int, noise, prediction_type="epsilon",
              clip_sample=True, thresholding=False, clip_sample_range=1.0, variance_type="fixed_small"):
    """scheduler_ddpm.py:439-526 without the guidance hook. ``noise`` is the N(0,1) draw used when t > 0.
    Returns (prev_sample, pred_original_sample, posterior_mean)."""
    prev_t = t - T // n
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    b_t, b_prev = 1 - a_t, 1 - a_prev
    cur_alpha = a_t / a_prev
    cur_beta = 1 - cur_alpha
    if prediction_type == "epsilon":
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
    elif prediction_type == "sample":
        x0 = model_output
    elif prediction_type in ("v_prediction", "v_pred_depth"):
        x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
    else:
        raise ValueError(prediction_type)
    x0 = _clip(x0, thresholding, clip_sample, clip_sample_range)
    c0 = (a_prev ** 0.5 * cur_beta) / b_t
    ct = cur_alpha ** 0.5 * b_prev / b_t
    mean = c0 * x0 + ct * sample
    var_term = 0
    if t > 0:
        v = ddpm_variance(t, prev_t, alphas_cumprod, variance_type)
        var_term = (v if variance_type == "fixed_small_log" else v ** 0.5) * noise
    return mean + var_term, x0, mean


def add_noise(x0, noise, timesteps, alphas_cumprod):
    """scheduler_ddim.py:501-525 / scheduler_ddpm.py:528-549."""
    a = alphas_cumprod[timesteps].reshape(-1, *([1] * (x0.dim() - 1)))
    return a ** 0.5 * x0 + (1 - a) ** 0.5 * noise


def get_velocity(sample, noise, timesteps, alphas_cumprod):
    """scheduler_ddim.py:528-547 / scheduler_ddpm.py:589-609."""
    a = alphas_cumprod[timesteps].reshape(-1, *([1] * (sample.dim() - 1)))
    return a ** 0.5 * noise - (1 - a) ** 0.5 * sample

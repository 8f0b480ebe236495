"""``MyDDIMScheduler`` — same constructor, attributes, ``set_timesteps`` and ``step()`` signature as
``/root/reference/core/scheduler_ddim.py`` (class :133, ``__init__`` :189-238, ``set_timesteps`` :301-344,
``step`` :346-498), with the tensor arithmetic of ``step`` executed by one fused CUDA kernel
(``d3_sched_ddim_step``).  The scalar coefficients are computed here with the same fp32 0-dim tensor expressions
as the reference, so ``prev_sample`` / ``pred_original_sample`` are bit-identical to the reference's fp32 result.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib, ops
from ._config import BaseOutput, ConfigMixin, register_to_config

PRED_TYPES = {"epsilon": 0, "sample": 1, "v_prediction": 2}


@dataclass
class DDIMSchedulerOutput(BaseOutput):
    """Fields of the reference's output class (scheduler_ddim.py:31-50)."""

    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None
    images_perturbed_orig: torch.Tensor = None
    images_pred_prev: torch.Tensor = None
    images_purturbed_prev: torch.Tensor = None


def betas_for_alpha_bar(num_diffusion_timesteps, max_beta=0.999, alpha_transform_type="cosine"):
    """Glide cosine schedule (scheduler_ddim.py:53-94)."""
    if alpha_transform_type == "cosine":
        fn = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    elif alpha_transform_type == "exp":
        fn = lambda t: math.exp(t * -12.0)  # noqa: E731
    else:
        raise ValueError(f"Unsupported alpha_tranform_type: {alpha_transform_type}")
    n = num_diffusion_timesteps
    return torch.tensor([min(1 - fn((i + 1) / n) / fn(i / n), max_beta) for i in range(n)], dtype=torch.float32)


def rescale_zero_terminal_snr(betas):
    """scheduler_ddim.py:97-130 (arXiv 2305.08891, Alg. 1)."""
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = (alphas_bar_sqrt - aT) * (a0 / (a0 - aT))
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
    return 1 - alphas


def make_betas(beta_schedule, num_train_timesteps, beta_start, beta_end, trained_betas, cls_name, allow_sigmoid):
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        return betas_for_alpha_bar(num_train_timesteps)
    if beta_schedule == "sigmoid" and allow_sigmoid:
        betas = torch.linspace(-6, 6, num_train_timesteps)
        return torch.sigmoid(betas) * (beta_end - beta_start) + beta_start
    raise NotImplementedError(f"{beta_schedule} does is not implemented for {cls_name}")


def make_timesteps(config, num_inference_steps: int) -> np.ndarray:
    """Timestep spacing shared by both schedulers (scheduler_ddim.py:320-342, scheduler_ddpm.py:270-296)."""
    T = config.num_train_timesteps
    if config.timestep_spacing == "linspace":
        return np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
    if config.timestep_spacing == "leading":
        step_ratio = T // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        return ts + config.steps_offset
    if config.timestep_spacing == "trailing":
        step_ratio = T / num_inference_steps
        return np.round(np.arange(T, 0, -step_ratio)).astype(np.int64) - 1
    raise ValueError(f"{config.timestep_spacing} is not supported. Please make sure to choose one of 'leading' or "
                     "'trailing'.")


def clip_mode_of(config) -> int:
    # thresholding with sample_max_value == 1.0 is clamp(-1, 1) / 1; other values would need the
    # per-image quantile and are refused rather than silently approximated.
    if config.thresholding:
        if float(config.sample_max_value) != 1.0:
            raise NotImplementedError("thresholding with sample_max_value != 1.0 is outside the reference's "
                                      "configurations (create_sampler never sets it, config.py:200-203)")
        return 2
    return 1 if config.clip_sample else 0


_RANDN_POOL = None


def _randn_pool():
    global _RANDN_POOL
    if _RANDN_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _RANDN_POOL = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1), thread_name_prefix="d3-randn")
    return _RANDN_POOL


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers' ``randn_tensor`` semantics: a CPU generator draws on the CPU and the result is moved; a list of
    generators draws one sample each (batch-shard-invariant noise).  A list of CPU generators is drawn by
    a small thread pool (the generators are independent, so the values are those of the serial loop) straight into one
    pinned buffer that is then copied to the device asynchronously."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        g0 = generator[0] if isinstance(generator, (list, tuple)) else generator
        if g0.device.type == "cpu" and device.type != "cpu":
            rand_device = torch.device("cpu")
    if isinstance(generator, (list, tuple)):
        one = (1,) + tuple(shape[1:])
        if rand_device.type == "cpu" and len(generator) > 1:
            buf = torch.empty((len(generator),) + tuple(shape[1:]), dtype=dtype or torch.float32,
                              pin_memory=(device.type == "cuda"))
            list(_randn_pool().map(lambda ig: torch.randn(one, generator=ig[1], dtype=buf.dtype, out=buf[ig[0]:ig[0] + 1]),
                                   enumerate(generator)))
            return buf.to(device, non_blocking=True) if device.type == "cuda" else buf
        return torch.cat([torch.randn(one, generator=g, device=rand_device, dtype=dtype) for g in generator]).to(device)
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)


class MyDDIMScheduler(ConfigMixin):
    order = 1

    @register_to_config
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 clip_sample: bool = True, set_alpha_to_one: bool = True, steps_offset: int = 0,
                 prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False):
        self.betas = make_betas(beta_schedule, num_train_timesteps, beta_start, beta_end, trained_betas,
                                self.__class__, allow_sigmoid=False)
        if rescale_betas_zero_snr:
            self.betas = rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _get_variance(self, timestep, prev_timestep):
        ac = self.alphas_cumprod.cpu()
        alpha_prod_t = ac[timestep]
        alpha_prod_t_prev = ac[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod.cpu()
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        return (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.config.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(make_timesteps(self.config, num_inference_steps)).to(device)

    def step_coeffs(self, timestep: int, eta: float = 0.0, use_clipped_model_output: bool = False) -> "_lib.DDIMCoeffs":
        """The scalar part of ``step`` (scheduler_ddim.py:413-456), evaluated with the reference's own fp32
        0-dim tensor expressions."""
        if self.config.prediction_type not in PRED_TYPES:
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, "
                             "`sample`, or `v_prediction`")
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        ac = self.alphas_cumprod.cpu()
        alpha_prod_t = ac[timestep]
        alpha_prod_t_prev = ac[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod.cpu()
        beta_prod_t = 1 - alpha_prod_t
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * variance ** (0.5)
        c = _lib.DDIMCoeffs()
        c.prediction_type = PRED_TYPES[self.config.prediction_type]
        c.sqrt_alpha_t = float(alpha_prod_t ** (0.5))
        c.sqrt_beta_t = float(beta_prod_t ** (0.5))
        c.sqrt_alpha_prev = float(alpha_prod_t_prev ** (0.5))
        c.dir_coeff = float((1 - alpha_prod_t_prev - std_dev_t ** 2) ** (0.5))
        c.std_dev = float(std_dev_t)
        c.clip_mode = clip_mode_of(self.config)
        c.clip_range = float(self.config.clip_sample_range)
        c.use_clipped_model_output = int(bool(use_clipped_model_output))
        return c

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = True, guidance=None, decoder=None, denormer=None, raw_mask=None, left_image=None,
             right_image=None, raw_depth=None) -> Union[DDIMSchedulerOutput, Tuple]:
        if self.num_inference_steps is None:
            raise ValueError(
                "Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = int(timestep)
        c = self.step_coeffs(t, eta, use_clipped_model_output)
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError(
                    "Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                    " `variance_noise` stays `None`.")
            if variance_noise is None:
                variance_noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                              dtype=model_output.dtype)
        else:
            variance_noise = None
        guided = guidance is not None and guidance.flow_guidance_weight > 0.0
        out_dtype = sample.dtype
        m32 = model_output if model_output.dtype == torch.float32 else model_output.float()
        s32 = sample if sample.dtype == torch.float32 else sample.float()
        n32 = None if variance_noise is None else variance_noise.float()
        prev_noised, x0, prev, dirv = ops.sched_ddim_step(m32, s32, c, None if guided else n32,
                                                          want_mean=(n32 is not None and not guided),
                                                          want_dir=guided)
        if guided:
            prev = prev_noised  # no noise was added above
            if hasattr(guidance, "perturb_fused"):   # x0' and sqrt(a_prev) * x0' + dir in the guidance kernel itself
                perturbed, prev_perturbed = guidance.perturb_fused(
                    x0, timestep, decoder, denormer, c.sqrt_alpha_prev, dirv, 1.0, raw_mask=raw_mask,
                    left_image=left_image, right_image=right_image, raw_depth=raw_depth)
            else:
                perturbed = guidance.perturb(x0, timestep, decoder, denormer, raw_mask=raw_mask,
                                             left_image=left_image, right_image=right_image, raw_depth=raw_depth)
                prev_perturbed = ops.lincomb(c.sqrt_alpha_prev, perturbed, 1.0, dirv)  # scheduler_ddim.py:468
            x0 = perturbed
            prev_noised = prev_perturbed if n32 is None else ops.lincomb(1.0, prev_perturbed, c.std_dev, n32)
        else:
            perturbed = x0
            if prev is None:
                prev = prev_noised
            prev_perturbed = prev
        if out_dtype != torch.float32:
            prev_noised, x0, perturbed, prev, prev_perturbed = (v.to(out_dtype) for v in
                                                                (prev_noised, x0, perturbed, prev, prev_perturbed))
        if not return_dict:
            return (prev_noised,)
        return DDIMSchedulerOutput(prev_sample=prev_noised, pred_original_sample=x0, images_perturbed_orig=perturbed,
                                   images_pred_prev=prev, images_purturbed_prev=prev_perturbed)

    # --- training-side helpers (not on the sampling hot path; plain tensor expressions as in the reference) ---
    def add_noise(self, original_samples, noise, timesteps):
        """scheduler_ddim.py:501-525."""
        self.alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device)
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)
        a = ac[timesteps.to(original_samples.device)].flatten()
        a = a.reshape(-1, *([1] * (original_samples.dim() - 1)))
        return a ** 0.5 * original_samples + (1 - a) ** 0.5 * noise

    def get_velocity(self, sample, noise, timesteps):
        """scheduler_ddim.py:528-547."""
        self.alphas_cumprod = self.alphas_cumprod.to(device=sample.device)
        ac = self.alphas_cumprod.to(dtype=sample.dtype)
        a = ac[timesteps.to(sample.device)].flatten()
        a = a.reshape(-1, *([1] * (sample.dim() - 1)))
        return a ** 0.5 * noise - (1 - a) ** 0.5 * sample

    def __len__(self):
        return self.config.num_train_timesteps

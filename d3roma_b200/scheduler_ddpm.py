"""``MyDDPMScheduler`` — same constructor, attributes, ``set_timesteps`` and ``step()`` signature as
``/root/reference/core/scheduler_ddpm.py`` (class :101, ``__init__`` :149-206, ``set_timesteps`` :225-298,
``_get_variance`` :300-338, ``previous_timestep`` :388-401, ``step`` :403-526), with the tensor arithmetic of
``step`` executed by one fused CUDA kernel (``d3_sched_ddpm_step``).  "Dynamic thresholding" is the clamp it
arithmetically is: no per-image quantile sort.  Variance noise is drawn exactly as the reference
does (``randn_tensor(shape, generator, device, dtype)``), so seeded runs reproduce the reference's RNG stream.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib, ops
from ._config import BaseOutput, ConfigMixin, register_to_config
from .scheduler_ddim import PRED_TYPES, clip_mode_of, make_betas, make_timesteps, randn_tensor


@dataclass
class DDPMSchedulerOutput(BaseOutput):
    """Fields of the reference's output class (scheduler_ddpm.py:36-55)."""

    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None
    images_perturbed_orig: torch.Tensor = None
    images_pred_prev: torch.Tensor = None
    images_purturbed_prev: torch.Tensor = None


class MyDDPMScheduler(ConfigMixin):
    order = 1

    @register_to_config
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 variance_type: str = "fixed_small", clip_sample: bool = True, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995,
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 steps_offset: int = 0):
        self.betas = make_betas(beta_schedule, num_train_timesteps, beta_start, beta_end, trained_betas,
                                self.__class__, allow_sigmoid=True)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alphas_cumprod_prev = torch.cat([torch.tensor([1.0]), self.alphas_cumprod])[:-1]
        self.betas_tilde = (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * self.betas
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self.variance_type = variance_type
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None,
                      timesteps: Optional[List[int]] = None):
        if num_inference_steps is not None and timesteps is not None:
            raise ValueError("Can only pass one of `num_inference_steps` or `custom_timesteps`.")
        if timesteps is not None:
            for i in range(1, len(timesteps)):
                if timesteps[i] >= timesteps[i - 1]:
                    raise ValueError("`custom_timesteps` must be in descending order.")
            if timesteps[0] >= self.config.num_train_timesteps:
                raise ValueError(f"`timesteps` must start before `self.config.train_timesteps`:"
                                 f" {self.config.num_train_timesteps}.")
            timesteps = np.array(timesteps, dtype=np.int64)
            self.custom_timesteps = True
        else:
            if num_inference_steps > self.config.num_train_timesteps:
                raise ValueError(
                    f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                    f"`self.config.train_timesteps`: {self.config.num_train_timesteps} as the unet model trained "
                    f"with this scheduler can only handle maximal {self.config.num_train_timesteps} timesteps.")
            self.num_inference_steps = num_inference_steps
            self.custom_timesteps = False
            timesteps = make_timesteps(self.config, num_inference_steps)
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def previous_timestep(self, timestep):
        if self.custom_timesteps:
            index = (self.timesteps == timestep).nonzero(as_tuple=True)[0][0]
            if index == self.timesteps.shape[0] - 1:
                return torch.tensor(-1)
            return self.timesteps[index + 1]
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def _get_variance(self, t, predicted_variance=None, variance_type=None):
        prev_t = self.previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        current_beta_t = 1 - alpha_prod_t / alpha_prod_t_prev
        variance = (1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t
        variance = torch.clamp(variance, min=1e-20)
        if variance_type is None:
            variance_type = self.config.variance_type
        if variance_type == "fixed_small":
            return variance
        if variance_type == "fixed_small_log":
            return torch.exp(0.5 * torch.log(variance))
        if variance_type == "fixed_large":
            return current_beta_t
        if variance_type == "fixed_large_log":
            return torch.log(current_beta_t)
        if variance_type == "learned":
            return predicted_variance
        if variance_type == "learned_range":
            min_log, max_log = torch.log(variance), torch.log(current_beta_t)
            frac = (predicted_variance + 1) / 2
            return frac * max_log + (1 - frac) * min_log
        return variance

    def step_coeffs(self, t: int) -> "_lib.DDPMCoeffs":
        """Scalar part of ``step`` (scheduler_ddpm.py:447-504) with the reference's fp32 tensor expressions."""
        if self.config.prediction_type not in ("epsilon", "sample", "v_prediction", "v_pred_depth"):
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, "
                             "`sample` or `v_prediction`  for the DDPMScheduler.")
        if self.variance_type in ("learned", "learned_range"):
            raise NotImplementedError("learned variance types are not used by the reference's configurations")
        prev_t = self.previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        c = _lib.DDPMCoeffs()
        c.prediction_type = PRED_TYPES.get(self.config.prediction_type, 2)
        c.sqrt_alpha_t = float(alpha_prod_t ** (0.5))
        c.sqrt_beta_t = float(beta_prod_t ** (0.5))
        c.x0_coeff = float((alpha_prod_t_prev ** (0.5) * current_beta_t) / beta_prod_t)
        c.xt_coeff = float(current_alpha_t ** (0.5) * beta_prod_t_prev / beta_prod_t)
        if t > 0:
            v = self._get_variance(t)
            c.sigma = float(v if self.variance_type == "fixed_small_log" else v ** 0.5)
        else:
            c.sigma = 0.0
        c.clip_mode = clip_mode_of(self.config)
        c.clip_range = float(self.config.clip_sample_range)
        return c

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, generator=None,
             return_dict: bool = True, guidance=None, decoder=None, raw_depth=None, raw_mask=None, left_image=None,
             right_image=None) -> Union[DDPMSchedulerOutput, Tuple]:
        t = int(timestep)
        c = self.step_coeffs(t)
        noise = None
        if t > 0:
            noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                 dtype=model_output.dtype)
        guided = guidance is not None and guidance.flow_guidance_weight > 0.0
        out_dtype = sample.dtype
        m32 = model_output if model_output.dtype == torch.float32 else model_output.float()
        s32 = sample if sample.dtype == torch.float32 else sample.float()
        n32 = None if noise is None else noise.float()
        prev_noised, x0, mean = ops.sched_ddpm_step(m32, s32, c, None if guided else n32,
                                                    want_mean=(n32 is not None and not guided))
        if guided:
            # the reference's call omits the positional `denormer` (scheduler_ddpm.py:507-511, a TypeError there);
            # the hook is invoked with denormer=None here and the guidance object falls back to its own `.denormer`
            # (set by the pipeline from its `denorm=` argument).
            mean = prev_noised
            if hasattr(guidance, "perturb_fused"):
                perturbed, mean_perturbed = guidance.perturb_fused(
                    x0, t, decoder, None, c.x0_coeff, s32, c.xt_coeff, raw_depth=raw_depth, raw_mask=raw_mask,
                    left_image=left_image, right_image=right_image)
            else:
                perturbed = guidance.perturb(x0, t, decoder, None, raw_depth=raw_depth, raw_mask=raw_mask,
                                             left_image=left_image, right_image=right_image)
                mean_perturbed = ops.lincomb(c.x0_coeff, perturbed, c.xt_coeff, s32)  # scheduler_ddpm.py:512
            prev_noised = mean_perturbed if n32 is None else ops.lincomb(1.0, mean_perturbed, c.sigma, n32)
        else:
            perturbed = x0
            if mean is None:
                mean = prev_noised
            mean_perturbed = mean
        if out_dtype != torch.float32:
            prev_noised, x0, perturbed, mean, mean_perturbed = (v.to(out_dtype) for v in
                                                                (prev_noised, x0, perturbed, mean, mean_perturbed))
        if not return_dict:
            return (prev_noised,)
        return DDPMSchedulerOutput(prev_sample=prev_noised, pred_original_sample=x0, images_perturbed_orig=perturbed,
                                   images_pred_prev=mean, images_purturbed_prev=mean_perturbed)

    # --- training-side helpers (not on the sampling hot path; plain tensor expressions as in the reference) ---
    def add_noise(self, original_samples, noise, timesteps):
        """scheduler_ddpm.py:528-549."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        a = ac[timesteps.to(original_samples.device)].flatten().reshape(-1, *([1] * (original_samples.dim() - 1)))
        return a ** 0.5 * original_samples + (1 - a) ** 0.5 * noise

    def posterior_mean_variance(self, clean_images, noisy_images, timesteps):
        """scheduler_ddpm.py:559-587."""
        dev, dt = clean_images.device, clean_images.dtype
        t = timesteps.to(dev)
        a_t = self.alphas_cumprod.to(device=dev, dtype=dt)[t]
        a_prev = self.alphas_cumprod_prev.to(device=dev, dtype=dt)[t]
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        c0 = (a_prev ** 0.5 * (1 - cur_alpha)) / b_t
        ct = cur_alpha ** 0.5 * b_prev / b_t
        shape = (-1, *([1] * (clean_images.dim() - 1)))
        return c0.reshape(shape) * clean_images + ct.reshape(shape) * noisy_images, None

    def get_velocity(self, sample, noise, timesteps):
        """scheduler_ddpm.py:589-609."""
        ac = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        a = ac[timesteps.to(sample.device)].flatten().reshape(-1, *([1] * (sample.dim() - 1)))
        return a ** 0.5 * noise - (1 - a) ** 0.5 * sample

    def __len__(self):
        return self.config.num_train_timesteps

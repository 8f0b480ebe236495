"""ORACLE (test infrastructure, not product code): CPU restatement of the pixel-space denoising loop
``GuidedDiffusionPipeline.__call__`` (/root/reference/core/custom_pipelines.py:97-205) on top of
``oracle/unet_ref.py`` and ``oracle/schedulers_ref.py``.

Pinned against ``tests/golden/pipeline.npz`` (the reference's own pipeline class driving the restated UNet;
oracle/gen_golden.py) in tests/test_oracle_pipeline.py.  Also used, timed, as the CPU baseline of bench.py.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import schedulers_ref as S

COND_ORDER = {
    # custom_pipelines.py:158-171
    "rgb": ("rgb",),
    "rgb+raw": ("rgb", "raw"),
    "rgb+right": ("left", "right"),
    "left+right+raw": ("left", "right", "raw"),
    "rgb+left+right": ("rgb", "left", "right"),
    "rgb+left+right+raw": ("rgb", "left", "right", "raw"),
}


@torch.no_grad()
def run_pipeline(unet, x_T: torch.Tensor, conds: dict, cond_channels: str, *, sampler: str = "ddim",
                 num_inference_steps: int = 10, num_train_timesteps: int = 128,
                 beta_schedule: str = "squaredcos_cap_v2", prediction_type: str = "sample",
                 steps_offset: int = 1, ddpm_noise: Optional[Callable[[int, torch.Size], torch.Tensor]] = None,
                 record: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """Returns ``images`` = final sample clamped to [-1, 1] (custom_pipelines.py:196). ``x_T`` is the CPU-drawn
    start noise (:117-123). DDIM: set_alpha_to_one=False, clip_sample=True (config.py:204-208). DDPM:
    thresholding (== clamp), fixed_small variance, noise supplied by ``ddpm_noise(step_index, shape)``."""
    T, n = num_train_timesteps, num_inference_steps
    ac = torch.cumprod(1.0 - S.make_betas(beta_schedule, T), dim=0)
    off = steps_offset if sampler == "ddim" else 0
    timesteps = S.make_timesteps(T, n, "leading", off)
    cond = torch.cat([conds[k] for k in COND_ORDER[cond_channels]], dim=1)
    image = x_T
    for i, t in enumerate(timesteps):
        m = unet(torch.cat([image, cond], dim=1), torch.tensor(int(t))).sample
        if record is not None:
            record.append(m.clone())
        if sampler == "ddim":
            image, _ = S.ddim_step(m, int(t), image, ac, ac[0], T, n, prediction_type=prediction_type,
                                   clip_sample=True)
        else:
            noise = ddpm_noise(i, m.shape) if int(t) > 0 else None
            image, _, _ = S.ddpm_step(m, int(t), image, ac, T, n, noise, prediction_type=prediction_type,
                                      thresholding=True)
    return image.clamp(-1, 1)

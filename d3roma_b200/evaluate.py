"""``eval_batch`` — the batched caller of the denoising loop, mirroring ``/root/reference/evaluate.py:66-148`` for the
part that is on the hot path: replicate-pad to a multiple of ``divis_by`` (:73-74) -> ``num_inference_rounds`` calls of
the pipeline (:90-104) -> unpad (:103) -> denormalise (:122) -> mean over rounds (:123).  Metric computation
(``compute_errors``, numpy, :134-138) and plotting are evaluation-side post-processing outside this path.
"""
from __future__ import annotations

from typing import Optional

import torch

from .prepost import InputPadder, Normalizer


@torch.no_grad()
def denoise_batch(pipeline, normalizer: Normalizer, *, rgb=None, left=None, right=None, sim_disp=None,
                  raw_depth=None, mask=None, cond_channels: str = "left+right+raw", depth_channels: int = 1,
                  num_inference_steps: int = 10, num_intermediate_images: int = 1, num_inference_rounds: int = 1,
                  divis_by: int = 32, generator=None, to_host: bool = False) -> torch.Tensor:
    """Inputs are [B,C,h,w] fp32 tensors in [-1, 1] (``sim_disp`` = normalised raw disparity), on the host (pinned
    or pageable) or on the device.  Returns the predicted disparity in pixels, [B,1,h,w] (mean over rounds)."""
    dev = pipeline.device
    some = next(t for t in (rgb, left, right, sim_disp) if t is not None)
    padder = InputPadder(some.shape, divis_by=divis_by)

    def up(t):
        return None if t is None else t.to(dev, torch.float32, non_blocking=True)

    rgb, left, right, sim_disp, raw_depth = padder.pad(up(rgb), up(left), up(right), up(sim_disp), up(raw_depth))
    preds = []
    for _ in range(num_inference_rounds):
        out = pipeline(rgb, left, right, sim_disp, raw_depth, mask, num_inference_steps=num_inference_steps,
                       num_intermediate_images=num_intermediate_images, depth_channels=depth_channels,
                       cond_channels=cond_channels, generator=generator)
        preds.append(padder.unpad(out.images))
    pred = torch.cat(preds, dim=1)                       # [B, R*dc, h, w]  (evaluate.py:113)
    disp = normalizer.denormalize(pred)                  # [B, R, h, w]
    disp = disp.mean(dim=1, keepdim=True) if disp.shape[1] > 1 else disp
    if to_host:
        host = torch.empty(disp.shape, dtype=disp.dtype, pin_memory=True)
        host.copy_(disp, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host
    return disp

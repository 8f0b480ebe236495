"""ORACLE (test infrastructure, not product code): CPU restatement of the stereo photometric guidance
gradient of the reference — ``FlowGuidance.grad`` stereo branch (/root/reference/core/guidance.py:293-331),
``SSIM.forward`` (:28-42) and ``camera.unnormalize_disp`` (utils_d3roma/camera.py:599-602, 778-780) —
with torch autograd on the CPU, exactly as the reference computes it.

Pinned against ``tests/golden/guidance.npz`` (made by replaying those reference lines with the reference's
own ``SSIM`` module in oracle/gen_golden.py) in tests/test_oracle_guidance.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def ssim_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """guidance.py:28-42: 3x3 mean windows on reflect-padded images, C1=1e-4, C2=9e-4, clamp((1-n/d)/2,0,1)."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x + sigma_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def unnormalize_disp(d, min_disp: float, max_disp: float, shift: float = 0.0):
    """camera.py:778-780 followed by the clip of camera.py:599-602."""
    return (min_disp + (d - shift) * (max_disp - min_disp)).clip(min_disp, max_disp)


def guidance_grad(disp_norm: torch.Tensor, left: torch.Tensor, right: torch.Tensor, min_disp: float,
                  max_disp: float, shift: float = 0.0, w: float = 0.95):
    """d(loss)/d(disp_norm) with loss = w * sum(SSIM_loss(left, warp)) + (1-w) * sum|left - warp| where
    warp = right sampled at x - disparity (bilinear, border padding, align_corners) — guidance.py:298-323.
    disp_norm [B,1,H,W], left/right [B,3,H,W(+margin)] fp32 CPU. Returns (grad [B,1,H,W], warp [B,3,H,W])."""
    d = disp_norm.detach().squeeze(1).clone().requires_grad_(True)
    B, H, W = d.shape
    with torch.enable_grad():
        disp = unnormalize_disp(d, min_disp, max_disp, shift)
        margin = right.shape[-1] - W
        xx, yy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
        xx = xx[None].repeat(B, 1, 1)
        yy = yy[None].repeat(B, 1, 1)
        gx = (xx - disp + margin) / ((W + margin - 1) / 2.0) - 1
        gy = yy / ((H - 1) / 2.0) - 1
        warp = F.grid_sample(right, torch.stack([gx, gy], dim=-1), align_corners=True, mode="bilinear",
                             padding_mode="border")
        lcrop = left[..., margin:]
        loss = F.l1_loss(lcrop, warp, reduction="sum")
        loss = w * ssim_loss(lcrop, warp).sum() + (1 - w) * loss
        (g,) = torch.autograd.grad(loss, d)
    return g.unsqueeze(1), warp.detach()

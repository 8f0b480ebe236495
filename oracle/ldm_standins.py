"""ORACLE tooling (test infrastructure only): tiny deterministic stand-ins for the three third-party networks of the
latent family (``UNet2DConditionModel``, ``AutoencoderKL``, text encoder output) with the attribute surface the
reference's ``GuidedLatentDiffusionPipeline`` touches (``vae.encoder / quant_conv / post_quant_conv / decoder``,
``unet(x, t, ctx).sample``; /root/reference/core/custom_pipelines.py:35-59, 255-272, 386).  They let the reference's
own pipeline class and ``d3roma_b200.GuidedLatentDiffusionPipeline`` be driven with identical modules so that the LOOP
logic can be compared; they say nothing about the real networks."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn


class TinyVAE(nn.Module):
    def __init__(self, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.encoder = nn.Sequential(nn.Conv2d(3, 8, 3, stride=2, padding=1), nn.SiLU(),
                                     nn.Conv2d(8, 8, 3, stride=2, padding=1), nn.SiLU(),
                                     nn.Conv2d(8, 8, 3, stride=2, padding=1))
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.decoder = nn.Sequential(nn.Conv2d(4, 16, 3, padding=1), nn.SiLU(), nn.Upsample(scale_factor=8, mode="nearest"),
                                     nn.Conv2d(16, 3, 3, padding=1))
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.05))
            self.decoder[-1].weight.mul_(0.04)    # keep the decoded maps inside (-1, 1): the pipeline clamps them
            self.decoder[0].weight.mul_(0.3)


class TinyCondUNet(nn.Module):
    def __init__(self, in_channels: int, seed: int = 1):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = nn.Conv2d(in_channels, 16, 3, padding=1)
        self.c2 = nn.Conv2d(16, 4, 3, padding=1)
        self.ctx = nn.Linear(8, 16)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.2 if p.dim() > 1 else 0.05))

    def forward(self, x, t, encoder_hidden_states):
        tt = torch.as_tensor(t, dtype=torch.float32, device=x.device).reshape(-1)[:1]
        h = self.c1(x) + torch.sin(tt / 100.0).view(1, 1, 1, 1)
        h = h + self.ctx(encoder_hidden_states.mean(dim=1)).view(x.shape[0], 16, 1, 1)
        return SimpleNamespace(sample=self.c2(torch.nn.functional.silu(h)))


def empty_text_embed(seed: int = 2) -> torch.Tensor:
    """Stand-in for the [1, 2, 1024] empty-prompt embedding (here [1, 2, 8])."""
    return torch.randn(1, 2, 8, generator=torch.Generator().manual_seed(seed))

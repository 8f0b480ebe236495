"""Seeded synthetic stereo batches with the reference's dataset batch keys
(``ActiveStereoDataset.__getitem__``, /root/reference/data/stereo_datasets.py:520-544) — there are no dataset files in
this environment, so benchmarks and tests feed the pipeline with synthetic tensors of the same names, shapes, dtypes and
value ranges: images in [-1, 1] (``(uint8 / 255 - 0.5) * 2``, inference.py:193), ground-truth depth
uniform in [0.3, 2.0] m, a "raw" sensor depth = ground truth + noise with 20 % holes, disparities = fxb / depth.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from .prepost import SIM_CAMERA, Camera, Normalizer


def make_batch(indices: Sequence[int], h: int = 360, w: int = 640, camera: Camera = SIM_CAMERA,
               normalizer: Normalizer = None, seed: int = 1234, device="cpu") -> Dict[str, torch.Tensor]:
    """One batch (host tensors unless ``device`` is given).  Every sample is generated from ``seed + global index``."""
    if normalizer is None:
        normalizer = Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    bound, t, s = float(normalizer.ch_bounds[0]), float(normalizer.t), float(normalizer.s)
    out = {k: [] for k in ("raw_disp", "normalized_disp", "sim_disp", "sim_mask", "sim_disp_unnorm", "normalized_rgb",
                           "left_image", "right_image", "mask", "depth", "raw_depth")}
    for i in indices:
        g = torch.Generator().manual_seed(seed + int(i))
        img = lambda: (torch.randint(0, 256, (3, h, w), generator=g).float() / 255.0 - 0.5) * 2  # noqa: E731
        rgb, left, right = img(), img(), img()
        depth = torch.rand((1, h, w), generator=g) * 1.7 + 0.3
        raw_depth = depth * (1.0 + 0.02 * torch.randn((1, h, w), generator=g))
        raw_depth[torch.rand((1, h, w), generator=g) < 0.2] = 0.0
        disp = camera.fxb / depth
        sim = torch.where(raw_depth > 0, camera.fxb / raw_depth.clamp_min(1e-6), torch.zeros_like(raw_depth))
        norm = lambda d: ((d.clamp(max=bound) / bound) - t) * s  # noqa: E731  (Normalizer "average", gamma 1)
        out["raw_disp"].append(disp)
        out["normalized_disp"].append(norm(disp))
        out["sim_disp"].append(norm(sim))
        out["sim_mask"].append((sim > 0).float())
        out["sim_disp_unnorm"].append(sim)
        out["normalized_rgb"].append(rgb)
        out["left_image"].append(left)
        out["right_image"].append(right)
        out["mask"].append(torch.ones_like(disp))
        out["depth"].append(depth)
        out["raw_depth"].append(raw_depth)
    batch = {k: torch.stack(v).to(device) for k, v in out.items()}
    batch["fxb"] = torch.full((len(indices),), float(camera.fxb), dtype=torch.float32, device=device)
    batch["index"] = torch.tensor(list(indices))
    return batch

"""``eval_batch`` — the batched caller of the denoising loop, mirroring ``/root/reference/evaluate.py:66-148`` for the
part that is on the hot path: replicate-pad to a multiple of ``divis_by`` (:73-74) -> ``num_inference_rounds`` calls of
the pipeline (:90-104) -> unpad (:103) -> denormalise (:122) -> mean over rounds (:123).  Metric computation
(``compute_errors``, numpy, :134-138) and plotting are evaluation-side post-processing outside this path.
"""
from __future__ import annotations

import functools
from typing import Optional

import torch

from types import SimpleNamespace

import numpy as np

from . import metrics as M
from .prepost import InputPadder, Normalizer


def default_config(**overrides) -> SimpleNamespace:
    """The fields of the reference's ``TrainingConfig`` (config.py:36-136) that ``eval_batch`` reads, with the
    pixel-space SceneFlow defaults (conf/task/train_sceneflow.yaml, eval_sceneflow.yaml)."""
    cfg = dict(divis_by=32, num_inference_rounds=1, num_inference_timesteps=10, num_intermediate_images=1,
               noise_rgb=False, depth_channels=1, cond_channels="left+right+raw", ssi=False, normalize_mode="average",
               num_chs=1, ch_bounds=[128.], ch_gammas=[1.], norm_t=0.5, norm_s=2.0, safe_ssi=False,
               ransac_error_threshold=0.6, ensemble=False, eval_dataset=["synthetic"], camera_resolution="640x360",
               prediction_space="disp")
    cfg.update(overrides)
    return SimpleNamespace(**cfg)


def denormalize(config, pred_disps, raw_disp=None, mask=None):
    """inference.py:8-35 / evaluate.py:44-64: SSI scale-shift alignment or Normalizer.denormalize."""
    return Normalizer.from_config(config).denormalize(pred_disps, raw_disp, mask)


@torch.no_grad()
def eval_batch(config, pipeline, disable_bar=True, fxb=None, normalized_rgb=None, raw_disp=None, mask=None,
               left_image=None, right_image=None, sim_disp=None, raw_depth=None, generator=None, as_numpy=True,
               **kwargs):
    """``eval_batch`` of /root/reference/evaluate.py:66-148 with the same signature and return tuple
    ``(pred_disps_unnormalized, metrics, uncertainties, error_map, out)``: pad -> R rounds of the pipeline -> unpad ->
    denormalise -> mean over rounds -> metrics.  Everything runs on the device; with ``as_numpy=False`` the maps stay
    CUDA tensors and only the [B, 11] metric rows are computed (no D2H sync of full-resolution maps)."""
    dev = pipeline.device
    up = lambda t: None if t is None else t.to(dev, torch.float32)  # noqa: E731
    normalized_rgb, left_image, right_image, sim_disp, raw_depth, raw_disp, mask = map(
        up, (normalized_rgb, left_image, right_image, sim_disp, raw_depth, raw_disp, mask))
    image_shape = normalized_rgb.shape if normalized_rgb is not None else left_image.shape
    padder = InputPadder(image_shape, divis_by=config.divis_by)
    p_rgb, p_left, p_right, p_sim, p_raw = padder.pad(normalized_rgb, left_image, right_image, sim_disp, raw_depth)
    assert config.num_inference_rounds > 0, "num_inference_rounds should be greater than 0"
    preds, out = [], None
    denorm = functools.partial(denormalize, config)                            # evaluate.py:76
    for r in range(config.num_inference_rounds):
        pipeline.set_progress_bar_config(desc=f"Denoising, round={r}", disable=disable_bar)
        out = pipeline(p_rgb, p_left, p_right, p_sim, p_raw, mask, num_inference_steps=config.num_inference_timesteps,
                       num_intermediate_images=config.num_intermediate_images, add_noise_rgb=config.noise_rgb,
                       depth_channels=config.depth_channels, cond_channels=config.cond_channels, denorm=denorm,
                       generator=generator)
        out.images = padder.unpad(out.images)
        preds.append(out.images)
    # only the last round's intermediates are kept, unpadded and denormalised like the reference's (evaluate.py:105-110);
    # the pipeline aliases one stack five times, so it is converted once and every field gets its own handle on the result
    if out.images_pred_orig is not None:
        inter = denorm(padder.unpad(out.images_pred_orig), raw_disp, mask)
        out.images_pred_orig = inter
        out.images_perturbed_orig = inter.clone()
        out.images_pred_prev = inter.clone()
        out.images_sampled_prev = inter.clone()
        out.images_purturbed_pred_prev = inter.clone()
    pred_disps = torch.cat(preds, dim=1)                                       # [B, R*dc, H, W]
    if getattr(config, "ensemble", False):                                     # evaluate.py:117-120
        pred_disps = torch.stack([M.ensemble_depths(pred_disps[i])[0] for i in range(pred_disps.shape[0])]).unsqueeze(1)
    pred_un = denormalize(config, pred_disps, raw_disp, mask).mean(dim=1)      # [B, H, W]
    gt = raw_disp.squeeze(1)
    if config.eval_dataset[0] == "SceneFlow":                                  # evaluate.py:126-130
        scale = 960. / float(config.camera_resolution.split("x")[0])
        lo, hi = 0.5 / scale, 192.0 / scale
        m = (gt >= lo) & (gt < hi)
        pred_un = pred_un.clamp(lo, hi)
    else:
        m = mask.squeeze(1) > 0
    rows = M.compute_errors_device(gt, pred_un, m, fxb)                        # [B, 11] fp64 on the device
    mf = m.to(pred_un.dtype)
    n_rounds = pred_disps.shape[1] // config.depth_channels
    unc = pred_disps.std(dim=1, unbiased=False) * mf if n_rounds > 1 else None  # np.std over rounds (:140-142)
    err = (gt - pred_un).abs() * mf
    if as_numpy:
        rows_h = rows.cpu().numpy()
        return (pred_un.cpu().numpy(), (rows_h[:, :5], rows_h[:, 5:]), None if unc is None else unc.cpu().numpy(),
                err.cpu().numpy(), out)
    return pred_un, rows, unc, err, out


@torch.no_grad()
def evaluate_intermediate_metrics(config, fxb, intermediates, gt_disps, gt_masks, gt_indexes=None):
    """``evaluate_intermediate_metrics`` of /root/reference/evaluate.py:28-57: the metric rows of every plotted step's
    predicted x0 (``intermediates.images_pred_orig`` [B, N, H, W], already denormalised by ``eval_batch``) against the
    ground truth -> ``(disp_metrics [B, N, 5], depth_metrics [B, N, 6])`` numpy.  One reduction kernel over the B*N maps;
    nothing but the rows leaves the device.  Reference quirk kept: ``fxb.view(B, 1).repeat(B, N, 1)`` flattens to
    ``fxb[i % B]`` for map i = b*N + n (not ``fxb[b]``) - identical whenever the batch shares one camera."""
    dev = intermediates.images_pred_orig.device
    pred = intermediates.images_pred_orig.to(torch.float32)
    B, N, H, W = pred.shape
    gt = gt_disps.to(dev, torch.float32).repeat(1, N, 1, 1)
    m = gt_masks.to(dev).repeat(1, N, 1, 1).bool()
    if config.eval_dataset[0] == "SceneFlow":
        scale = 960. / float(config.camera_resolution.split("x")[0])
        lo, hi = 0.5 / scale, 192.0 / scale
        m = (pred >= lo) & (pred < hi) & m
        pred = pred.clamp(lo, hi)
    else:
        pred = pred.clamp(min=0.25)                                            # prevent 0 disparities (:47)
    fx = None
    if fxb is not None:
        fx = torch.as_tensor(fxb).reshape(-1).to(dev, torch.float64)[torch.arange(B * N, device=dev) % B]
    rows = M.compute_errors_device(gt.reshape(-1, H, W).contiguous(), pred.reshape(-1, H, W).contiguous(),
                                   m.reshape(-1, H, W).contiguous(), fx).cpu().numpy()
    return rows[:, :5].reshape(B, N, 5), rows[:, 5:].reshape(B, N, 6)


SPATIAL_KEYS = ("raw_disp", "normalized_disp", "sim_disp", "left_image", "right_image", "mask")


@torch.no_grad()
def eval_batch_coarse_to_fine(config, pipeline, batch: dict, disable_bar=True, generator=None):
    """The ``config.coarse_to_fine`` branch of the evaluation loop (evaluate.py:362-398): the six spatial entries of the
    batch are cut into four quadrants (the reference hard-codes 270 / 480, the half-way lines of its 540 x 960 frames;
    here ``H // 2`` / ``W // 2``), every quadrant goes through ``eval_batch`` on its own, the predictions are stitched
    back and the metric rows averaged over the quadrants.  Returns ``(pred_disps [B, H, W] numpy, disp_metrics [B, 5])``."""
    import numpy as np
    H, W = batch["raw_disp"].shape[2:]
    hy, hx = H // 2, W // 2
    cuts = ((slice(0, hy), slice(0, hx)), (slice(hy, H), slice(0, hx)), (slice(0, hy), slice(hx, W)),
            (slice(hy, H), slice(hx, W)))                                          # j = 0..3 as in the reference
    pred = np.empty((batch["raw_disp"].shape[0], H, W), dtype=np.float32)
    rows = []
    for ys, xs in cuts:
        # the reference slices SPATIAL_KEYS only (its SceneFlow batches carry no full-size rgb / raw_depth the pipeline
        # reads); every other [B, C, H, W] entry of the batch is cut the same way here
        mini = {k: (v[:, :, ys, xs] if (k in SPATIAL_KEYS or (torch.is_tensor(v) and v.dim() == 4
                                                             and tuple(v.shape[2:]) == (H, W))) and v is not None else v)
                for k, v in batch.items()}
        p, metrics, _, _, _ = eval_batch(config, pipeline, disable_bar, mini.get("fxb"), mini.get("normalized_rgb"),
                                         mini.get("raw_disp"), mini.get("mask"), mini.get("left_image"),
                                         mini.get("right_image"), mini.get("sim_disp"), mini.get("raw_depth"),
                                         generator=generator)
        pred[:, ys, xs] = p
        rows.append(metrics[0])
    return pred, np.mean(rows, axis=0)


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
                       cond_channels=cond_channels, denorm=normalizer.denormalize, generator=generator)
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

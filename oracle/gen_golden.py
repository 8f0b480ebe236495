"""ORACLE tooling: generate the golden vectors under ``tests/golden/`` by running the REFERENCE's own
Python (``/root/reference/core/*.py`` and ``utils_d3roma/*.py``, imported unmodified behind
``oracle/diffusers_stub.py``) in this container.  The reference tree does not exist on the GPU box,
so the vectors are committed as small ``.npz`` fixtures; this script is what made them.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

What is pinned:
  * ddim.npz / ddpm.npz   - ``MyDDIMScheduler`` / ``MyDDPMScheduler`` tables, timesteps and full multi-step
                            trajectories of ``step()`` for every prediction type / spacing / variance type used
                            (core/scheduler_ddim.py:346-498, core/scheduler_ddpm.py:403-526), plus ``add_noise`` /
                            ``get_velocity``.
  * prepost.npz           - ``Normalizer.normalize/denormalize`` (utils_d3roma/utils.py:171-273),
                            ``InputPadder.pad/unpad`` (:297-325), ``normalize_disp/unnormalize_disp`` known
                            answers (utils_d3roma/camera.py:1027-1044) and the sim camera scalars
                            (``Realsense.default_sim``: fxb, min_disp, max_disp).
  * guidance.npz          - the stereo photometric-guidance gradient: the arithmetic of
                            ``FlowGuidance.grad`` (core/guidance.py:298-323) evaluated with the reference's own
                            ``SSIM`` module, ``camera.unnormalize_disp`` and autograd (the function as written
                            cannot return - SURVEY §0.4 - so its body is replayed line by line here).
  * postproc.npz          - ``compute_errors`` (utils_d3roma/utils.py:554-653), ``compute_scale_and_shift`` (:488-528)
                            and ``Normalizer(ssi=True).normalize/denormalize`` (:171-182, 231-251) outputs.
  * pipeline.npz          - ``GuidedDiffusionPipeline.__call__`` (core/custom_pipelines.py:97-205) driving the
                            restated UNet (oracle/unet_ref.py, seed 0) on a small seeded input: per-step model
                            outputs and the returned ``images`` for DDIM-10 and DDPM-8.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
from oracle import diffusers_stub  # noqa: E402

diffusers_stub.install()
sys.path.insert(0, str(REF))

from core.scheduler_ddim import MyDDIMScheduler  # noqa: E402
from core.scheduler_ddpm import MyDDPMScheduler  # noqa: E402
from core.custom_pipelines import GuidedDiffusionPipeline  # noqa: E402
from core.guidance import SSIM, FlowGuidance  # noqa: E402
from utils_d3roma.utils import InputPadder, Normalizer, compute_errors, compute_scale_and_shift  # noqa: E402
from utils_d3roma import camera as ref_camera  # noqa: E402

OUT = ROOT / "tests" / "golden"

DDIM_CONFIGS = {
    # the BASELINE / SURVEY C.1 configuration (conf/task/train_sceneflow.yaml + config.py:204-208)
    "cos128_sample": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample",
                          clip_sample=True, set_alpha_to_one=False, steps_offset=1),
    "cos128_epsilon": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon",
                           clip_sample=True, set_alpha_to_one=False, steps_offset=1),
    "cos128_v": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="v_prediction",
                     clip_sample=True, set_alpha_to_one=False, steps_offset=1),
    "cos128_sample_thresh": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2",
                                 prediction_type="sample", thresholding=True, set_alpha_to_one=False,
                                 steps_offset=1),
    # LDM family constants
    "sl1000_epsilon": dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                           beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=False,
                           set_alpha_to_one=False, steps_offset=1),
    "lin1000_v_trailing": dict(num_train_timesteps=1000, beta_schedule="linear", prediction_type="v_prediction",
                               clip_sample=True, set_alpha_to_one=True, timestep_spacing="trailing"),
    "lin1000_eps_linspace": dict(num_train_timesteps=1000, beta_schedule="linear", prediction_type="epsilon",
                                 clip_sample=True, clip_sample_range=0.8, timestep_spacing="linspace"),
}

DDPM_CONFIGS = {
    "cos128_sample_thresh": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2",
                                 prediction_type="sample", thresholding=True, clip_sample=True),
    "cos128_epsilon": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon",
                           clip_sample=True),
    "cos128_v_log": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="v_prediction",
                         clip_sample=True, variance_type="fixed_small_log"),
    "lin1000_eps_large": dict(num_train_timesteps=1000, beta_schedule="linear", prediction_type="epsilon",
                              clip_sample=False, variance_type="fixed_large"),
    "sig200_sample": dict(num_train_timesteps=200, beta_schedule="sigmoid", prediction_type="sample",
                          clip_sample=True),
}


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def gen_ddim(store):
    shape = (2, 1, 4, 6)
    for name, cfg in DDIM_CONFIGS.items():
        for n_steps, eta in ((10, 0.0), (8, 0.5)):
            sch = MyDDIMScheduler.from_config(cfg)
            sch.set_timesteps(n_steps)
            g = torch.Generator().manual_seed(7)
            x = torch.randn(shape, generator=g) * 1.3
            key = f"{name}/n{n_steps}_eta{eta}"
            store[f"{key}/timesteps"] = _np(sch.timesteps)
            store[f"{key}/alphas_cumprod"] = _np(sch.alphas_cumprod)
            store[f"{key}/final_alpha_cumprod"] = _np(sch.final_alpha_cumprod)
            store[f"{key}/x0"] = _np(x)
            for i, t in enumerate(sch.timesteps):
                m = torch.randn(shape, generator=g) * 0.9
                noise = torch.randn(shape, generator=g) if eta > 0 else None
                out = sch.step(m, t, x, eta=eta, variance_noise=noise)
                store[f"{key}/m{i}"] = _np(m)
                if noise is not None:
                    store[f"{key}/noise{i}"] = _np(noise)
                store[f"{key}/prev{i}"] = _np(out.prev_sample)
                store[f"{key}/orig{i}"] = _np(out.pred_original_sample)
                x = out.prev_sample
        # add_noise / get_velocity
        sch = MyDDIMScheduler.from_config(cfg)
        g = torch.Generator().manual_seed(11)
        a, b = torch.randn(3, 1, 4, 6, generator=g), torch.randn(3, 1, 4, 6, generator=g)
        ts = torch.tensor([0, sch.config.num_train_timesteps // 2, sch.config.num_train_timesteps - 1])
        store[f"{name}/an_a"], store[f"{name}/an_b"], store[f"{name}/an_t"] = _np(a), _np(b), _np(ts)
        store[f"{name}/add_noise"] = _np(sch.add_noise(a, b, ts))
        store[f"{name}/velocity"] = _np(sch.get_velocity(a, b, ts))


def gen_ddpm(store):
    shape = (2, 1, 4, 6)
    for name, cfg in DDPM_CONFIGS.items():
        for n_steps in (8, 100):
            sch = MyDDPMScheduler.from_config(cfg)
            sch.set_timesteps(n_steps)
            g = torch.Generator().manual_seed(5)
            x = torch.randn(shape, generator=g) * 1.2
            key = f"{name}/n{n_steps}"
            store[f"{key}/timesteps"] = _np(sch.timesteps)
            store[f"{key}/x0"] = _np(x)
            steps = list(enumerate(sch.timesteps))
            if n_steps == 100:
                steps = steps[:4] + steps[-3:]
            store[f"{key}/step_ids"] = np.array([i for i, _ in steps])
            for i, t in steps:
                m = torch.randn(shape, generator=g) * 0.8
                out = sch.step(m, t, x, generator=torch.Generator().manual_seed(1000 + i))
                store[f"{key}/m{i}"] = _np(m)
                store[f"{key}/x{i}"] = _np(x)
                store[f"{key}/prev{i}"] = _np(out.prev_sample)
                store[f"{key}/orig{i}"] = _np(out.pred_original_sample)
                store[f"{key}/mean{i}"] = _np(out.images_pred_prev)
                x = out.prev_sample
        sch = MyDDPMScheduler.from_config(cfg)
        store[f"{name}/betas"] = _np(sch.betas)
        store[f"{name}/alphas_cumprod"] = _np(sch.alphas_cumprod)
        store[f"{name}/posterior_variance"] = _np(sch.posterior_variance)
        g = torch.Generator().manual_seed(13)
        a, b = torch.randn(3, 1, 4, 6, generator=g), torch.randn(3, 1, 4, 6, generator=g)
        ts = torch.tensor([0, sch.config.num_train_timesteps // 2, sch.config.num_train_timesteps - 1])
        store[f"{name}/an_a"], store[f"{name}/an_b"], store[f"{name}/an_t"] = _np(a), _np(b), _np(ts)
        store[f"{name}/add_noise"] = _np(sch.add_noise(a, b, ts))
        store[f"{name}/velocity"] = _np(sch.get_velocity(a, b, ts))
        pm, _ = sch.posterior_mean_variance(a, b, ts)
        store[f"{name}/posterior_mean"] = _np(pm)


def gen_prepost(store):
    g = torch.Generator().manual_seed(3)
    disp = torch.rand(1, 20, 36, generator=g) * 150.0  # some values above the 128 bound -> clamp path
    disp[0, :3, :5] = 0.0
    for tag, kw in (("avg1", dict(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.)),
                    ("avg3", dict(ssi=False, mode="average", num_chs=3, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.)),
                    ("avg1g", dict(ssi=False, mode="average", num_chs=1, ch_bounds=[64.], ch_gammas=[0.5], t=0.5, s=2.))):
        n = Normalizer(**kw)
        y = n.normalize(disp.clone())[0]
        store[f"norm/{tag}/in"] = _np(disp)
        store[f"norm/{tag}/out"] = _np(y)
        pred = (torch.rand(2, 2 * kw["num_chs"], 20, 36, generator=g) * 2 - 1)
        store[f"norm/{tag}/den_in"] = _np(pred)
        store[f"norm/{tag}/den_out"] = _np(n.denormalize(pred.clone()))
    for hw in ((360, 640), (540, 960), (37, 50), (64, 96)):
        x = torch.randn(2, 3, *hw, generator=g)
        p = InputPadder(x.shape, divis_by=32)
        y = p.pad(x)[0]
        store[f"pad/{hw[0]}x{hw[1]}/pad"] = np.array(p._pad)
        if hw[0] < 100:
            store[f"pad/{hw[0]}x{hw[1]}/in"] = _np(x)
            store[f"pad/{hw[0]}x{hw[1]}/out"] = _np(y)
        assert torch.equal(p.unpad(y), x)
        store[f"pad/{hw[0]}x{hw[1]}/shape"] = np.array(y.shape)
    # known answers asserted by the reference's own __main__ (utils_d3roma/camera.py:1027-1044)
    ka_in = [(20, 20, 120, 0), (120, 20, 120, 0), (30, 20, 120, 0), (70, 20, 120, 0), (0, 20, 120, 0),
             (20, 20, 120, 0.2), (0, 20, 120, 0.2), (120, 20, 120, 0.2)]
    store["camera/normalize_in"] = np.array(ka_in, dtype=np.float64)
    store["camera/normalize_out"] = np.array([ref_camera.normalize_disp(d, a, b, s) for d, a, b, s in ka_in])
    ku_in = [(0, 20, 120, 0), (1, 20, 120, 0), (0.5, 20, 120, 0), (0.2, 20, 120, 0.2), (1.2, 20, 120, 0.2),
             (0.7, 20, 120, 0.2)]
    store["camera/unnormalize_in"] = np.array(ku_in, dtype=np.float64)
    store["camera/unnormalize_out"] = np.array([ref_camera.unnormalize_disp(d, a, b, s) for d, a, b, s in ku_in])
    cam = ref_camera.Realsense.default_sim()
    cam.change_resolution("640x360") if hasattr(cam, "change_resolution") else None
    store["camera/sim"] = np.array([cam.fxb_depth, cam.min_disp, cam.max_disp, cam.config["shift"]], dtype=np.float64)
    # depth <-> disparity (inference.py:128-131, 246-249)
    depth = torch.rand(1, 20, 36, generator=g) * 2.0
    depth[depth < 0.3] = 0.0
    d2 = torch.zeros_like(depth)
    v = depth > 0
    d2[v] = float(cam.fxb_depth) / depth[v]
    store["camera/depth_in"] = _np(depth)
    store["camera/disp_out"] = _np(d2)


def _guidance_grad_reference(fg, ssim, cam, disp_raw, left_images, right_images, use_ssim=True):
    """core/guidance.py:298-323 replayed with the reference's own modules (the function cannot return)."""
    disp_raw = disp_raw.squeeze(1).clone()
    B, H, W = disp_raw.shape
    device = disp_raw.device
    with torch.enable_grad():
        w = 0.95
        disp_raw.requires_grad_(True)
        disp = cam.unnormalize_disp(disp_raw)
        margin_left = right_images.shape[-1] - W
        xx, yy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing='xy')
        xx = xx[None, ...].repeat(B, 1, 1).to(device)
        yy = yy[None, ...].repeat(B, 1, 1).to(device)
        xx = (xx - disp + margin_left) / ((W + margin_left - 1) / 2.) - 1
        yy = yy / ((H - 1) / 2.) - 1
        grid = torch.stack([xx, yy], dim=-1)
        warp_left_image = F.grid_sample(right_images, grid, align_corners=True, mode="bilinear",
                                        padding_mode="border")
        loss = F.l1_loss(left_images[..., margin_left:], warp_left_image, reduction='sum')
        if use_ssim:
            ssim_loss = ssim(left_images[..., margin_left:], warp_left_image).sum()
            loss = w * ssim_loss + (1 - w) * loss
        grads = torch.autograd.grad(loss, disp_raw)[0]
    return grads.unsqueeze(1), warp_left_image.detach()


def gen_guidance(store):
    cam = ref_camera.Realsense.default_sim()
    ssim = SSIM()
    fg = FlowGuidance(flow_guidance_weight=1.0)
    g = torch.Generator().manual_seed(21)
    for tag, (B, H, W) in (("a", (2, 16, 72)), ("b", (1, 20, 96))):
        left = torch.rand(B, 3, H, W, generator=g) * 2 - 1
        right = torch.rand(B, 3, H, W, generator=g) * 2 - 1
        # smooth-ish images so that SSIM is not saturated everywhere
        left = F.avg_pool2d(F.pad(left, (1, 1, 1, 1), mode="replicate"), 3, 1)
        right = F.avg_pool2d(F.pad(right, (1, 1, 1, 1), mode="replicate"), 3, 1)
        dn = torch.rand(B, 1, H, W, generator=g) * 0.2 - 0.03  # normalised disparity, some below 0 (clip)
        grad, warp = _guidance_grad_reference(fg, ssim, cam, dn, left, right)
        store[f"{tag}/left"], store[f"{tag}/right"], store[f"{tag}/disp"] = _np(left), _np(right), _np(dn)
        store[f"{tag}/grad"], store[f"{tag}/warp"] = _np(grad), _np(warp)
    store["cam"] = np.array([cam.min_disp, cam.max_disp, cam.config["shift"]], dtype=np.float64)


def gen_pipeline(store):
    from oracle.unet_ref import make_unet
    H, W = 32, 64
    g = torch.Generator().manual_seed(1234)
    left = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    right = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    raw = torch.rand(1, 1, H, W, generator=g) * 2 - 1
    store["left"], store["right"], store["raw"] = _np(left), _np(right), _np(raw)
    unet = make_unet(8, 1, seed=0, sample_size=(H, W))

    class Rec(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m, self.outs, self.sample_size = m, [], m.sample_size

        def forward(self, x, t):
            o = self.m(x, t)
            self.outs.append(o.sample.clone())
            return o

    for tag, sch, n in (
            ("ddim10", MyDDIMScheduler.from_config(DDIM_CONFIGS["cos128_sample"]), 10),
            ("ddpm8", MyDDPMScheduler.from_config(DDPM_CONFIGS["cos128_sample_thresh"]), 8)):
        rec = Rec(unet)
        pipe = GuidedDiffusionPipeline(rec, sch, FlowGuidance(flow_guidance_weight=0.0))
        torch.manual_seed(99)  # DDPM variance noise comes from the default generator
        out = pipe(None, left, right, raw, None, None, num_inference_steps=n, num_intermediate_images=2,
                   depth_channels=1, cond_channels="left+right+raw",
                   generator=torch.Generator().manual_seed(42))
        store[f"{tag}/images"] = _np(out.images)
        store[f"{tag}/intermediate"] = _np(out.images_pred_orig)
        store[f"{tag}/model_outputs"] = np.stack([_np(o) for o in rec.outs])


def make_stereo_pair(B, H, W, gen, max_shift=9.0):
    """Synthetic rectified pair with a known disparity: smooth random left image, smooth ground-truth disparity in
    [min_d, min_d + max_shift] px, right = left sampled at x + disp (so that warping right by the true disparity
    reproduces left up to interpolation)."""
    cam = ref_camera.Realsense.default_sim()
    left = torch.rand(B, 3, H + 8, W + 8, generator=gen)
    for _ in range(3):
        left = F.avg_pool2d(left, 3, 1)
    left = (left[..., 1:H + 1, 1:W + 1] - left.mean()) / left.std() * 0.35
    left = left.clamp(-1, 1)
    d = torch.rand(B, 1, H // 8 + 2, W // 8 + 2, generator=gen)
    d = F.interpolate(d, size=(H, W), mode="bilinear", align_corners=True)
    disp = float(cam.min_disp) + 0.5 + d[:, 0] * max_shift                     # [B,H,W] pixels
    # right(x') with x' = x - disp(x): build it by sampling left at x' + disp (first-order inverse warp)
    xx, yy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    gx = (xx[None] + disp) / ((W - 1) / 2.0) - 1
    gy = (yy[None] / ((H - 1) / 2.0) - 1).expand_as(gx)
    right = F.grid_sample(left, torch.stack([gx, gy], dim=-1), align_corners=True, mode="bilinear",
                          padding_mode="border")
    return left, right, disp


def gen_guided(store):
    """The guided DDIM loop: the reference's MyDDIMScheduler.step with its guidance hook (scheduler_ddim.py:461-472)
    driven by a FlowGuidance whose perturb() is the chain rule over the reference's own pieces - Normalizer.denormalize
    (the `denormer` hook argument), camera.normalize_disp / unnormalize_disp, the SSIM module and the lines of
    FlowGuidance.grad (guidance.py:298-323) - since the reference leaves that hook unconnected."""
    from oracle.unet_ref import make_unet
    cam = ref_camera.Realsense.default_sim()
    ssim = SSIM()
    H, W, B = 32, 64, 2
    gen = torch.Generator().manual_seed(777)
    left, right, disp_gt = make_stereo_pair(B, H, W, gen)
    normer = Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    raw = torch.stack([normer.normalize((disp_gt[b:b + 1] + torch.randn(1, H, W, generator=gen) * 2.0).clamp(min=0))[0]
                       for b in range(B)])
    x_T = torch.randn(B, 1, H, W, generator=gen)
    weight = 2e-3

    class StereoGuidance(FlowGuidance):
        def __init__(self, w):
            super().__init__(flow_guidance_weight=w, flow_guidance_mode="stereo")
            self.camera = cam
            self.log = []

        def loss_of(self, x0, denormer, left_image, right_image):
            px = denormer(x0)                                   # [B,1,H,W] pixels (utils.py:260-273)
            d_hat = cam.normalize_disp(px)                      # camera.py:771-776
            g, _ = _guidance_grad_reference(self, ssim, cam, d_hat, left_image, right_image)
            return g

        def perturb(self, pred_original_sample, t, decoder, denormer, raw_mask=None, left_image=None,
                    right_image=None, raw_depth=None):
            x = pred_original_sample.detach().clone().requires_grad_(True)
            with torch.enable_grad():
                px = denormer(x)
                d_hat = cam.normalize_disp(px).squeeze(1)
                disp = cam.unnormalize_disp(d_hat)
                Bn, Hh, Ww = disp.shape
                margin_left = right_image.shape[-1] - Ww
                xx, yy = torch.meshgrid(torch.arange(Ww), torch.arange(Hh), indexing='xy')
                xx = xx[None, ...].repeat(Bn, 1, 1)
                yy = yy[None, ...].repeat(Bn, 1, 1)
                xx = (xx - disp + margin_left) / ((Ww + margin_left - 1) / 2.) - 1
                yy = yy / ((Hh - 1) / 2.) - 1
                grid = torch.stack([xx, yy], dim=-1)
                warp = F.grid_sample(right_image, grid, align_corners=True, mode="bilinear", padding_mode="border")
                loss = F.l1_loss(left_image[..., margin_left:], warp, reduction='sum')
                loss = 0.95 * ssim(left_image[..., margin_left:], warp).sum() + 0.05 * loss
                g = torch.autograd.grad(loss, x)[0]
            out = (pred_original_sample - self.flow_guidance_weight * g).detach()
            self.log.append(dict(x0=pred_original_sample.clone(), g=g.clone(), x0p=out.clone(), loss=float(loss)))
            return out

    unet = make_unet(8, 1, seed=0, sample_size=(H, W))
    for tag, w in (("guided", weight), ("plain", 0.0)):
        sch = MyDDIMScheduler.from_config(DDIM_CONFIGS["cos128_sample"])
        sch.set_timesteps(10)
        guid = StereoGuidance(w)
        image = x_T.clone()
        cond = torch.cat([left, right, raw], dim=1)
        for i, t in enumerate(sch.timesteps):
            with torch.no_grad():
                m = unet(torch.cat([image, cond], dim=1), t).sample
            out = sch.step(m, t, image, guidance=guid, denormer=normer.denormalize, left_image=left, right_image=right)
            if tag == "guided":
                store[f"guided/m{i}"], store[f"guided/x{i}"] = _np(m), _np(image)
                store[f"guided/x0{i}"] = _np(guid.log[-1]["x0"])
                store[f"guided/g{i}"] = _np(guid.log[-1]["g"])
                store[f"guided/x0p{i}"] = _np(out.images_perturbed_orig)
                store[f"guided/prev{i}"] = _np(out.prev_sample)
                store[f"guided/loss{i}"] = np.float64(guid.log[-1]["loss"])
            image = out.prev_sample
        store[f"{tag}/images"] = _np(image.clamp(-1, 1))
    store["left"], store["right"], store["raw"], store["x_T"] = _np(left), _np(right), _np(raw), _np(x_T)
    store["disp_gt"] = _np(disp_gt)
    store["weight"] = np.float64(weight)
    store["cam"] = np.array([cam.min_disp, cam.max_disp, cam.config["shift"]], dtype=np.float64)
    # photometric loss of the final disparities (reference pieces): guidance must not increase it
    with torch.no_grad():
        for tag in ("guided", "plain"):
            px = normer.denormalize(torch.from_numpy(store[f"{tag}/images"]))
            d_hat = cam.normalize_disp(px)
    for tag in ("guided", "plain"):
        x = torch.from_numpy(store[f"{tag}/images"])
        g = StereoGuidance(1.0)
        g.perturb(x, 0, None, normer.denormalize, left_image=left, right_image=right)
        store[f"{tag}/final_loss"] = np.float64(g.log[-1]["loss"])


def gen_unet_oracle(store):
    """Regression fixture of the UNet restatement itself (oracle/unet_ref.py).  diffusers is absent, so this does NOT pin
    the restatement to the reference's arithmetic - it freezes the restatement's own outputs (seeded weights, seeded
    inputs) so that it cannot drift unnoticed while kernels are compared against it."""
    from oracle.unet_ref import make_unet
    torch.set_num_threads(8)
    for tag, (cin, dc, hw, t) in {"c8": (8, 1, (32, 64), 49), "c5": (5, 1, (32, 32), 109), "c12": (12, 3, (32, 32), 1)}.items():
        net = make_unet(cin, dc, seed=0, sample_size=hw)
        x = torch.randn(1, cin, *hw, generator=torch.Generator().manual_seed(123))
        store[f"{tag}/x"] = _np(x)
        store[f"{tag}/y"] = _np(net(x, torch.tensor(t)).sample)
        store[f"{tag}/meta"] = np.array([cin, dc, hw[0], hw[1], t])
        store[f"{tag}/n_params"] = np.array(sum(p.numel() for p in net.parameters()))


def gen_latent(store):
    """The reference's GuidedLatentDiffusionPipeline (core/custom_pipelines.py:208-460) driven with tiny stand-in
    networks (oracle/ldm_standins.py): pins the LOOP logic of the latent family - latent encode / decode conventions,
    conditioning order per cond_channels key, scheduler steps, returned image_final and intermediates."""
    import contextlib
    import io
    from core.custom_pipelines import GuidedLatentDiffusionPipeline
    from oracle.ldm_standins import TinyCondUNet, TinyVAE, empty_text_embed
    g = torch.Generator().manual_seed(2025)
    B, H, W = 2, 32, 64
    rgb, left, right = (torch.rand(B, 3, H, W, generator=g) * 2 - 1 for _ in range(3))
    sim = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    raw_depth = torch.rand(B, 1, H, W, generator=g) * 2.0
    raw_depth[raw_depth < 0.4] = 0.0
    store["rgb"], store["left"], store["right"], store["sim"], store["raw_depth"] = (_np(v) for v in
                                                                                       (rgb, left, right, sim, raw_depth))
    vae, ete = TinyVAE(), empty_text_embed()
    cases = {
        # clipped-x0 scheduler configs keep the random stand-in UNet's trajectories inside the decoder's linear range
        "ddim_lrr": ("left+right+raw", 16, MyDDIMScheduler.from_config(DDIM_CONFIGS["cos128_v"]), 10, 5, 0.0),
        "ddim_rgbraw": ("rgb+raw", 12, MyDDIMScheduler.from_config(DDIM_CONFIGS["lin1000_v_trailing"]), 8, 2, 0.0),
        "ddpm_rgb": ("rgb", 8, MyDDPMScheduler.from_config(DDPM_CONFIGS["cos128_sample_thresh"]), 8, 4, 0.0),
        "ddim_grad": ("rgb+left+right+raw", 20, MyDDIMScheduler.from_config(DDIM_CONFIGS["cos128_epsilon"]), 4, 2, 1.0),
    }
    for tag, (cc, cin, sch, n, plots, w) in cases.items():
        unet = TinyCondUNet(cin)
        guid = FlowGuidance(flow_guidance_weight=w, flow_guidance_mode="gradient")
        pipe = GuidedLatentDiffusionPipeline(unet, vae, None, None, sch, guid)
        torch.manual_seed(321)     # DDPM variance noise: default generator
        with contextlib.redirect_stdout(io.StringIO()):
            out = pipe(rgb, left, right, sim, raw_depth, None, empty_text_embed=ete, num_inference_steps=n,
                       num_intermediate_images=plots, cond_channels=cc, denorm=lambda d: (d / 2.0 + 0.5) * 2.0,
                       generator=torch.Generator().manual_seed(7))
        store[f"{tag}/images"] = _np(out.images)
        store[f"{tag}/intermediate"] = _np(out.images_pred_orig)


def gen_latent_nets(store):
    """The reference's GuidedLatentDiffusionPipeline driven with the ORACLE restatements of its two networks
    (oracle/ldm_ref.py, reduced widths, seeded default init): the end-to-end target of the CUDA UNet2DConditionModel +
    AutoencoderKL inside d3roma_b200.GuidedLatentDiffusionPipeline (tests/test_latent_nets_gpu.py)."""
    import contextlib
    import io
    from core.custom_pipelines import GuidedLatentDiffusionPipeline
    from oracle.ldm_ref import make_cond_unet, make_vae
    g = torch.Generator().manual_seed(77)
    B, H, W = 2, 64, 96
    rgb, left, right = (torch.rand(B, 3, H, W, generator=g) * 2 - 1 for _ in range(3))
    sim = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    store["rgb"], store["left"], store["right"], store["sim"] = (_np(v) for v in (rgb, left, right, sim))
    ete = torch.randn(1, 2, 64, generator=g)
    store["empty_text_embed"] = _np(ete)
    vae = make_vae(seed=11, block_out_channels=(128, 128, 128, 128))
    cases = {
        "ddim_rgbraw": ("rgb+raw", 12, MyDDIMScheduler.from_config(DDIM_CONFIGS["sl1000_epsilon"]), 6, 3),
        "ddim_lrr": ("left+right+raw", 16, MyDDIMScheduler.from_config(DDIM_CONFIGS["lin1000_v_trailing"]), 4, 2),
    }
    for tag, (cc, cin, sch, n, plots) in cases.items():
        unet = make_cond_unet(in_channels=cin, seed=5, block_out_channels=(64, 128, 128), attention_head_dim=(1, 2, 2),
                              cross_attention_dim=64)
        pipe = GuidedLatentDiffusionPipeline(unet, vae, None, None, sch,
                                             FlowGuidance(flow_guidance_weight=0.0, flow_guidance_mode="gradient"))
        with contextlib.redirect_stdout(io.StringIO()):
            out = pipe(rgb, left, right, sim, None, None, empty_text_embed=ete, num_inference_steps=n,
                       num_intermediate_images=plots, cond_channels=cc, generator=torch.Generator().manual_seed(9))
        store[f"{tag}/images"] = _np(out.images)
        store[f"{tag}/intermediate"] = _np(out.images_pred_orig)


def gen_piecewise(store):
    """Normalizer(mode="piecewise") of the reference (utils_d3roma/utils.py:198-229, 275-293): 3 channels with bounds
    [64, 32, 32] (the class defaults), 2 channels with a gamma, and the 1-channel degenerate case."""
    g = torch.Generator().manual_seed(17)
    disp = torch.rand(1, 20, 36, generator=g) * 150.0     # beyond the 128 total bound -> clamp path
    disp[0, :3, :5] = 0.0
    disp[0, 5, :4] = torch.tensor([64.0, 96.0, 128.0, 63.999])   # exactly on the channel boundaries
    for tag, kw in (("pw3", dict(num_chs=3, ch_bounds=[64., 32., 32.], ch_gammas=[1., 1., 1.])),
                    ("pw2g", dict(num_chs=2, ch_bounds=[64., 64.], ch_gammas=[0.5, 1.])),
                    ("pw1", dict(num_chs=1, ch_bounds=[128.], ch_gammas=[1.]))):
        n = Normalizer(ssi=False, mode="piecewise", t=0.5, s=2.0, **kw)
        store[f"{tag}/in"] = _np(disp)
        store[f"{tag}/out"] = _np(n.normalize(disp.clone())[0])
        pred = torch.rand(2, 2 * kw["num_chs"], 20, 36, generator=g) * 2 - 1
        store[f"{tag}/den_in"] = _np(pred)
        store[f"{tag}/den_out"] = _np(n.denormalize(pred.clone()))
        # round trip of a normalised map
        store[f"{tag}/roundtrip"] = _np(n.denormalize(n.normalize(disp.clone())[0][None]))


def gen_reproject(store):
    """Depth -> colour-frame reprojection (camera.py:610-701) run by the reference's own real-camera object
    (Realsense.default_real("wsl"): non-identity colour<-depth extrinsics) on small synthetic depth maps with holes:
    the aggressive (4-corner) fill, the single rounded fill, and the cropped variant."""
    cam = ref_camera.Realsense.default_real("wsl")
    rng = np.random.default_rng(5)
    H, W = cam.H, cam.W
    store["hw"] = np.array([H, W])
    store["K_depth"], store["K_color"] = cam.K_depth.vec, cam.K_color.vec
    store["color_to_depth"], store["T_fc"], store["T_cl"] = cam.extrinsics["color_to_depth"], cam._T_fc, cam.T_cl
    ys, xs = np.mgrid[0:H, 0:W]
    # sensor-like depth: millimetre-quantised uint16 (stored as such: the fixture stays small), metres = mm / 1000
    mm = np.round((0.45 + 0.35 * np.sin(xs / 37.0) * np.cos(ys / 23.0)) * 1000.0 + rng.integers(0, 3, (H, W)))
    mm[(xs * 7 + ys * 13) % 11 == 0] = 0                 # holes
    mm[100:140, 200:260] = 310                           # a near box: occlusion / collisions at its border
    mm = mm.astype(np.uint16)
    depth = mm.astype(np.float32) / np.float32(1000.0)
    store["depth_mm"] = mm
    to_mm = lambda o: np.round(o.astype(np.float64) * 1000.0).astype(np.uint16)   # noqa: E731  (outputs are copies of inputs)
    cam.config["aggressive_fill"] = True
    store["out_aggressive_mm"] = to_mm(cam.transform_depth_to_rgb_frame(depth.copy()))
    cam.config["aggressive_fill"] = False
    store["out_round_mm"] = to_mm(cam.transform_depth_to_rgb_frame(depth.copy()))
    # cropped variant: a 200 x 320 window of the sensor image with shifted principal point
    oy, ox, ch, cw = 60, 110, 200, 320
    cropped_K = cam.K_depth.arr.copy()
    cropped_K[0, 2] -= ox
    cropped_K[1, 2] -= oy
    store["cropped_K"], store["crop"] = cropped_K, np.array([oy, ox, ch, cw])
    store["out_cropped_mm"] = to_mm(cam.transform_cropped_depth_to_rgb_frame(depth[oy:oy + ch, ox:ox + cw].copy(),
                                                                             cropped_K))


def gen_postproc(store):
    """compute_errors / compute_scale_and_shift / Normalizer(ssi=True) of the reference (utils_d3roma/utils.py)."""
    g = torch.Generator().manual_seed(31)
    B, H, W = 3, 24, 40
    gt = torch.rand(B, H, W, generator=g) * 60 + 5
    gt[torch.rand(B, H, W, generator=g) < 0.25] = 0.0           # invalid pixels
    pred = gt + torch.randn(B, H, W, generator=g) * 1.5
    pred[1, :2, :3] = -0.5                                       # non-positive predictions -> clip branch (sample 1)
    mask = gt > 0
    md, mz = compute_errors(gt.numpy(), pred.numpy(), "disp", mask.numpy(), [24.5471958] * B)
    store["err/gt"], store["err/pred"], store["err/mask"] = _np(gt), _np(pred), _np(mask)
    store["err/disp"], store["err/depth"] = md, mz
    md2, mz2 = compute_errors(gt.numpy(), np.where(mask.numpy(), np.abs(pred.numpy()) + 0.1, 0).astype(np.float32),
                              "disp", None, None)
    store["err/disp_nofxb"], store["err/depth_nofxb"] = md2, mz2
    p2 = torch.rand(4, H * W, generator=g) * 2 - 1
    t2 = p2 * 37.5 + 20.0 + torch.randn(4, H * W, generator=g) * 0.3
    m2 = torch.rand(4, H * W, generator=g) > 0.3
    store["ssi/pred"], store["ssi/target"], store["ssi/mask"] = _np(p2), _np(t2), _np(m2)
    store["ssi/st"] = _np(compute_scale_and_shift(p2, t2, m2.float()))
    n = Normalizer(ssi=True, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0, safe_ssi=False)
    x = gt[:1].clone()
    y, low, up = n.normalize(x, mask[:1])
    store["ssi/norm_in"], store["ssi/norm_mask"], store["ssi/norm_out"] = _np(x), _np(mask[:1]), _np(y)
    store["ssi/low_up"] = np.array([float(low), float(up)], dtype=np.float32)
    yb = (torch.rand(2, 2, H, W, generator=g) * 2 - 1)
    raw = torch.rand(2, 1, H, W, generator=g) * 60 + 5
    mk = torch.rand(2, 1, H, W, generator=g) > 0.2
    store["ssi/den_in"], store["ssi/den_raw"], store["ssi/den_mask"] = _np(yb), _np(raw), _np(mk)
    store["ssi/den_out"] = _np(n.denormalize(yb, raw, mk.float()))


def gen_frontend(store):
    """Frame preprocessing of inference.py:108-213 pinned to cv2 itself: small uint8 frames / depth maps and what
    ``cv2.resize`` + the reference's torch arithmetic produce for them (restated in oracle/frontend_ref.py because the
    reference class needs hydra + a checkpoint to construct)."""
    import cv2
    rng = np.random.default_rng(2024)
    store["cv2_version"] = np.array([int(v) for v in cv2.__version__.split(".")[:3]])
    cases = [("down", (45, 80), (36, 64), 3), ("up", (20, 30), (36, 64), 3), ("same", (36, 64), (36, 64), 3),
             ("odd", (53, 97), (36, 64), 4), ("gray", (48, 72), (36, 64), 1)]
    for name, (hs, ws), (h, w), c in cases:
        img = rng.integers(0, 256, (hs, ws) if c == 1 else (hs, ws, c), dtype=np.uint8)
        depth = (rng.random((hs, ws)) * 1.7 + 0.3).astype(np.float32)
        depth[rng.random((hs, ws)) < 0.2] = 0.0
        src = img if c == 1 else img[..., :3]
        if c == 1:
            src = np.tile(src[..., None], (1, 1, 3))
        res = cv2.resize(np.ascontiguousarray(src), (w, h), interpolation=cv2.INTER_LINEAR)
        t = torch.from_numpy(res).permute(2, 0, 1).float()
        d = cv2.resize(depth, dsize=(w, h), interpolation=cv2.INTER_NEAREST)
        dt = torch.from_numpy(d[..., None]).permute(2, 0, 1).float()
        raw_disp = torch.zeros_like(dt)
        valid = dt > 0
        raw_disp[valid] = 24.547 / dt[valid]
        store[f"{name}/img"], store[f"{name}/depth"] = img, depth
        store[f"{name}/size"] = np.array([h, w])
        store[f"{name}/resized_u8"] = res
        store[f"{name}/normalized"] = _np((t / 255. - 0.5) * 2)
        store[f"{name}/raw_disp"] = _np(raw_disp)
        store[f"{name}/norm_disp"] = _np(((torch.clamp(raw_disp, max=128.0) / 128.0) - 0.5) * 2.0)


def gen_ransac(store):
    """utils_d3roma/ransac.py RANSAC.fit run unmodified on the CPU with the permutations it drew recorded, plus the
    safe_ssi branch of Normalizer.denormalize (utils.py:240-251)."""
    import io
    import contextlib
    from utils_d3roma.ransac import RANSAC
    g = torch.Generator().manual_seed(77)
    B, H, W = 3, 24, 40
    HW = H * W
    gt = torch.rand(B, HW, generator=g) * 60 + 5
    a = torch.tensor([[50.], [64.], [33.]])                 # true scale / shift per row: gt = a * pred + b
    b = torch.tensor([[3.], [-2.], [1.]])
    pred = (gt - b) / a + torch.randn(B, HW, generator=g) * 0.004
    outl = torch.rand(B, HW, generator=g) < 0.02           # 2 % outliers
    pred = torch.where(outl, pred + (torch.rand(B, HW, generator=g) - 0.5) * 0.2, pred)
    mask = (torch.rand(B, HW, generator=g) > 0.2).float()
    perms = []
    real_randperm = torch.randperm

    def rec(n, **kw):
        p = real_randperm(n, **kw)
        perms.append(p.clone())
        return p
    torch.manual_seed(5)
    torch.randperm = rec
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            reg = RANSAC(n=0.1, k=10, d=0.2, t=0.6).fit(pred, gt, mask)
            perms_fit = [p.clone() for p in perms]
            perms.clear()
            n = Normalizer(ssi=True, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0,
                           safe_ssi=True, ransac_error_threshold=0.6)
            y = pred.view(B, 1, H, W)
            den = n.denormalize(y, gt.view(B, 1, H, W), mask.view(B, 1, H, W))
            perms_den = [p.clone() for p in perms]
    finally:
        torch.randperm = real_randperm
    store["pred"], store["gt"], store["mask"] = _np(pred), _np(gt), _np(mask)
    store["perms_fit"] = np.stack([_np(p) for p in perms_fit]).astype(np.int32)
    store["best_fit"], store["best_error"] = _np(reg.best_fit), _np(reg.best_error)
    store["best_num_inlier"] = _np(reg.best_num_inlier)
    store["perms_den"] = np.stack([_np(p) for p in perms_den]).astype(np.int32)
    store["den_out"] = _np(den)


def gen_ensemble(store):
    """ensemble_depths of the reference (utils_d3roma/utils.py:28-120) run unmodified on seeded stacks of maps
    (oracle/ensemble_ref.make_inputs): odd / even counts (torch.median's lower median), identical maps (ties), both
    reductions, a single pair."""
    import warnings
    from utils_d3roma.utils import ensemble_depths
    from oracle.ensemble_ref import make_inputs
    for tag, n, seed, ties, red in (("med5", 5, 0, False, "median"), ("med4", 4, 1, False, "median"),
                                    ("med6t", 6, 2, True, "median"), ("mean3", 3, 3, False, "mean"),
                                    ("mean5t", 5, 4, True, "mean"), ("med2", 2, 5, False, "median")):
        x = make_inputs(n, seed=seed, ties=ties)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            aligned, unc = ensemble_depths(x.clone(), reduction=red)
        store[f"{tag}/aligned"], store[f"{tag}/uncertainty"] = _np(aligned), _np(unc)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for name, fn in (("ddim", gen_ddim), ("ddpm", gen_ddpm), ("prepost", gen_prepost), ("guidance", gen_guidance),
                     ("pipeline", gen_pipeline), ("guided", gen_guided), ("reproject", gen_reproject), ("piecewise", gen_piecewise), ("latent", gen_latent), ("latent_nets", gen_latent_nets), ("unet_oracle", gen_unet_oracle), ("postproc", gen_postproc), ("frontend", gen_frontend), ("ransac", gen_ransac), ("ensemble", gen_ensemble)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:     # python oracle/gen_golden.py ensemble  -> only that fixture
            continue
        store = {}
        with torch.no_grad():
            fn(store)
        np.savez_compressed(OUT / f"{name}.npz", **store)
        print(f"{name}.npz: {len(store)} arrays, {(OUT / (name + '.npz')).stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()

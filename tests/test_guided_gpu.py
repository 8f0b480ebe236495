"""Guided sampling on the GPU: the hook chain x0 -> pixels -> camera-normalised disparity -> photometric gradient,
the fused update and the scheduler recombination (d3_guidance_step) against the golden recorded from the reference's
scheduler / SSIM / Normalizer / camera code (tests/golden/guided.npz) and against the CPU oracle."""
import functools
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
DDIM_CFG = dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample",
                clip_sample=True, set_alpha_to_one=False, steps_offset=1)


def _z():
    z = np.load(G / "guided.npz")
    return z, (lambda k: torch.from_numpy(z[k]).cuda())


def _normer():
    import d3roma_b200 as d
    return d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)


def _frac_bad(got, ref, scale, rtol=2e-3, atol_rel=2e-4):
    return ((got - ref).abs() > rtol * ref.abs() + atol_rel * scale).float().mean().item()


def test_x0_gradient_vs_reference_golden():
    import d3roma_b200 as d
    z, t = _z()
    fg = d.FlowGuidance(flow_guidance_weight=float(z["weight"]), flow_guidance_mode="stereo")
    assert abs(fg.camera.min_disp - z["cam"][0]) < 1e-5 and abs(fg.camera.max_disp - z["cam"][1]) < 1e-4
    for i in range(10):
        g = fg.loss_grad_x0(t(f"guided/x0{i}"), _normer().denormalize, t("left"), t("right"))
        ref = t(f"guided/g{i}")
        assert _frac_bad(g, ref, ref.abs().max()) < 2e-3, i   # isolated pixels sit on a clamp / floor discontinuity


def test_guided_ddim_steps_teacher_forced_vs_reference_golden():
    """Every step of the reference's guided DDIM trajectory: same model output and sample in, x0' and prev_sample out."""
    import d3roma_b200 as d
    z, t = _z()
    w = float(z["weight"])
    fg = d.FlowGuidance(flow_guidance_weight=w, flow_guidance_mode="stereo")
    sch = d.MyDDIMScheduler.from_config(DDIM_CFG)
    sch.set_timesteps(10)
    for i, ts in enumerate(sch.timesteps):
        out = sch.step(t(f"guided/m{i}"), ts, t(f"guided/x{i}"), guidance=fg, denormer=_normer().denormalize,
                       left_image=t("left"), right_image=t("right"))
        gmax = float(np.abs(z[f"guided/g{i}"]).max())
        for got, key in ((out.images_perturbed_orig, f"guided/x0p{i}"), (out.prev_sample, f"guided/prev{i}")):
            bad = ((got - t(key)).abs() > w * (2e-3 * t(f"guided/g{i}").abs() + 2e-4 * gmax) + 2e-6).float().mean().item()
            assert bad < 2e-3, (i, key, bad)
        # the unfused route (a guidance object without perturb_fused) gives the same tensors bit for bit
        class Plain:
            flow_guidance_weight = w
            perturb = staticmethod(functools.partial(fg.perturb))
        out2 = sch.step(t(f"guided/m{i}"), ts, t(f"guided/x{i}"), guidance=Plain(), denormer=_normer().denormalize,
                        left_image=t("left"), right_image=t("right"))
        assert torch.equal(out2.prev_sample, out.prev_sample) and torch.equal(out2.images_perturbed_orig,
                                                                             out.images_perturbed_orig)


def test_guided_pipeline_fp32_vs_reference_golden():
    """The whole guided loop (restated UNet in the fp32 path) through GuidedDiffusionPipeline.__call__ with denorm=."""
    import d3roma_b200 as d
    from oracle.unet_ref import make_unet
    z, t = _z()
    ref_net = make_unet(8, 1, seed=0, sample_size=(32, 64))
    net = d.UNet2DModel(sample_size=(32, 64), in_channels=8, out_channels=1, compute_dtype="fp32").to("cuda")
    net.load_state_dict(ref_net.state_dict())
    outs = {}
    for tag, w in (("guided", float(z["weight"])), ("plain", 0.0)):
        fg = d.FlowGuidance(flow_guidance_weight=w, flow_guidance_mode="stereo")
        pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(DDIM_CFG), fg)

        class FixedNoise(torch.Generator):
            pass
        # x_T of the golden is fed through sample_from (the pipeline's device-resident entry) to share the start noise
        cond = torch.cat([t("left"), t("right"), t("raw")], 1).contiguous()
        outs[tag] = pipe.sample_from(t("x_T"), cond, 10, left_images=t("left"), right_images=t("right"),
                                     denorm=_normer().denormalize)
        err = (outs[tag] - t(f"{tag}/images")).abs()
        assert err.mean().item() < 1e-3 and err.quantile(0.99).item() < 1e-2, (tag, err.mean().item(), err.max().item())
    assert not torch.equal(outs["guided"], outs["plain"])


def test_denormer_forms_and_errors():
    import d3roma_b200 as d
    from d3roma_b200.evaluate import default_config, denormalize
    z, t = _z()
    fg = d.FlowGuidance(flow_guidance_weight=1e-3, flow_guidance_mode="stereo")
    x0, l, r = t("guided/x00"), t("left"), t("right")
    n = _normer()
    ref = fg.perturb(x0, 5, None, n.denormalize, left_image=l, right_image=r)
    for den in (n, functools.partial(denormalize, default_config()), lambda y: (y / 2.0 + 0.5) * 128.0):
        assert torch.equal(fg.perturb(x0, 5, None, den, left_image=l, right_image=r), ref)
    fg.denormer = n                                              # the DDPM hook passes denormer=None
    assert torch.equal(fg.perturb(x0, 5, None, None, left_image=l, right_image=r), ref)
    fg.denormer = None
    with pytest.raises(ValueError):
        fg.perturb(x0, 5, None, None, left_image=l, right_image=r)
    with pytest.raises(NotImplementedError):
        fg.perturb(x0, 5, None, d.Normalizer(ssi=True, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.]),
                   left_image=l, right_image=r)
    with pytest.raises(ValueError):
        fg.perturb(x0, 5, None, lambda y: y ** 2, left_image=l, right_image=r)
    with pytest.raises(NotImplementedError):
        d.FlowGuidance().perturb(x0, 5, None, n)
    # perturb_start_ratio gates the early (noisy) part of the trajectory
    fg2 = d.FlowGuidance(flow_guidance_weight=1e-3, perturb_start_ratio=0.5, flow_guidance_mode="stereo")
    fg2.num_train_timesteps = 128
    assert torch.equal(fg2.perturb(x0, 100, None, n, left_image=l, right_image=r), x0)
    assert not torch.equal(fg2.perturb(x0, 20, None, n, left_image=l, right_image=r), x0)


def test_guided_descent_lowers_photometric_loss_on_a_true_stereo_pair():
    """Synthetic pair where right = left warped by a known disparity: starting 2 px off, repeated guidance updates must
    decrease the reference's photometric loss (evaluated by the CPU oracle) and move the disparity towards the truth."""
    import d3roma_b200 as d
    from oracle import guided_pipeline_ref as GP
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(11)
    B, H, W = 2, 96, 160
    cam = d.SIM_CAMERA
    left = torch.rand(B, 3, H + 8, W + 8, generator=gen)
    for _ in range(3):
        left = F.avg_pool2d(left, 3, 1)
    left = ((left[..., 1:H + 1, 1:W + 1] - left.mean()) / left.std() * 0.35).clamp(-1, 1)
    dsm = F.interpolate(torch.rand(B, 1, H // 8 + 2, W // 8 + 2, generator=gen), size=(H, W), mode="bilinear",
                        align_corners=True)[:, 0]
    disp = float(cam.min_disp) + 2.5 + dsm * 6.0
    xx, yy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    grid = torch.stack([(xx[None] + disp) / ((W - 1) / 2.0) - 1, (yy[None] / ((H - 1) / 2.0) - 1).expand_as(disp)], -1)
    right = F.grid_sample(left, grid, align_corners=True, mode="bilinear", padding_mode="border")
    n = _normer()
    x0 = (((disp + 2.0) / 128.0) - 0.5) * 2.0                                   # 2 px too large everywhere
    x0 = x0[:, None].contiguous()
    fg = d.FlowGuidance(flow_guidance_weight=2e-4, flow_guidance_mode="stereo")
    x = x0.cuda()
    losses, errs = [], []
    for it in range(13):
        if it % 4 == 0:
            l, _ = GP.x0_loss_and_grad(x.cpu(), left, right, 128.0, float(cam.min_disp), float(cam.max_disp))
            losses.append(float(l))
            errs.append(float((n.denormalize(x)[:, 0].cpu() - disp)[..., 12:-12].abs().mean()))
        x = fg.perturb(x, 1, None, n, left_image=left.cuda(), right_image=right.cuda())
    assert all(b < a for a, b in zip(losses, losses[1:])), losses
    assert errs[-1] < errs[0], errs

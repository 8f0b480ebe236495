"""Loop logic of ``GuidedLatentDiffusionPipeline`` against golden outputs of the REFERENCE's own pipeline
class (core/custom_pipelines.py:208-460) driven with the same tiny stand-in networks (oracle/ldm_standins.py): latent
encode / decode conventions, conditioning order of every ``cond_channels`` key, scheduler steps on the CUDA kernels,
returned ``image_final`` and intermediates.  The stand-ins are test infrastructure; the real networks are covered by
tests/test_ldm_unet_gpu.py / test_ldm_vae_gpu.py."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"

# scheduler configs of the golden run (oracle/gen_golden.py DDIM_CONFIGS / DDPM_CONFIGS, same names)
DDIM = {
    "cos128_v": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="v_prediction",
                     clip_sample=True, set_alpha_to_one=False, steps_offset=1),
    "lin1000_v_trailing": dict(num_train_timesteps=1000, beta_schedule="linear", prediction_type="v_prediction",
                               clip_sample=True, set_alpha_to_one=True, timestep_spacing="trailing"),
    "cos128_epsilon": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon",
                           clip_sample=True, set_alpha_to_one=False, steps_offset=1),
}
DDPM = {"cos128_sample_thresh": dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2",
                                      prediction_type="sample", thresholding=True, clip_sample=True)}


def _configs():
    return DDIM, DDPM


CASES = {
    "ddim_lrr": ("left+right+raw", 16, "ddim", "cos128_v", 10, 5, 0.0),
    "ddim_rgbraw": ("rgb+raw", 12, "ddim", "lin1000_v_trailing", 8, 2, 0.0),
    "ddpm_rgb": ("rgb", 8, "ddpm", "cos128_sample_thresh", 8, 4, 0.0),
    "ddim_grad": ("rgb+left+right+raw", 20, "ddim", "cos128_epsilon", 4, 2, 1.0),
}


@pytest.mark.parametrize("tag", list(CASES))
def test_latent_loop_vs_reference_golden(tag):
    import d3roma_b200 as d
    import d3roma_b200.scheduler_ddpm as sd
    from oracle.ldm_standins import TinyCondUNet, TinyVAE, empty_text_embed
    z = np.load(G / "latent.npz")
    ddim_cfgs, ddpm_cfgs = _configs()
    cc, cin, kind, cfg, n, plots, w = CASES[tag]
    sch = (d.MyDDIMScheduler.from_config(ddim_cfgs[cfg]) if kind == "ddim"
           else d.MyDDPMScheduler.from_config(ddpm_cfgs[cfg]))
    unet, vae = TinyCondUNet(cin).cuda(), TinyVAE().cuda()
    pipe = d.GuidedLatentDiffusionPipeline(unet, vae, None, None, sch,
                                           d.FlowGuidance(flow_guidance_weight=w, flow_guidance_mode="gradient"))
    rgb, left, right, sim, raw_depth = (torch.from_numpy(z[k]) for k in ("rgb", "left", "right", "sim", "raw_depth"))
    torch.manual_seed(321)
    orig = sd.randn_tensor   # the reference draws the DDPM variance noise from the default generator of the device
    sd.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(tuple(shape), dtype=dtype).to(device)
    try:
        out = pipe(rgb, left, right, sim, raw_depth, None, empty_text_embed=empty_text_embed(), num_inference_steps=n,
                   num_intermediate_images=plots, cond_channels=cc, denorm=lambda x: (x / 2.0 + 0.5) * 2.0,
                   generator=torch.Generator().manual_seed(7))
    finally:
        sd.randn_tensor = orig
    img = out.images.cpu().numpy()
    tol = 5e-3 if w > 0 else 2e-4   # the guided case runs ten Adam steps per scheduler step (sign-like updates)
    err = np.abs(img - z[f"{tag}/images"]).max()
    erri = np.abs(out.images_pred_orig.cpu().numpy() - z[f"{tag}/intermediate"]).max()
    print(f"{tag}: final {err:.2e} intermediates {erri:.2e}")
    assert out.images.shape == (2, 1, 32, 64) and out.images_pred_orig.shape == (2, plots, 32, 64)
    assert err <= tol and erri <= tol
    assert out.images_sampled_prev is out.images_pred_orig


def test_latent_pipeline_contract():
    import d3roma_b200 as d
    from oracle.ldm_standins import TinyCondUNet, TinyVAE, empty_text_embed
    ddim_cfgs, _ = _configs()
    pipe = d.GuidedLatentDiffusionPipeline(TinyCondUNet(8).cuda(), TinyVAE().cuda(), None, None,
                                           d.MyDDIMScheduler.from_config(ddim_cfgs["cos128_v"]),
                                           d.FlowGuidance(flow_guidance_weight=0.0, flow_guidance_mode="gradient"))
    imp = d.GuidedLatentDiffusionPipeline(pipe.unet, pipe.vae, None, None, pipe.scheduler,
                                          d.FlowGuidance(flow_guidance_weight=0.0))   # default mode: "imputation"
    with pytest.raises(AssertionError):      # custom_pipelines.py:392
        imp(torch.rand(1, 3, 32, 64), empty_text_embed=empty_text_embed(), num_inference_steps=4,
            num_intermediate_images=1, cond_channels="rgb")
    rgb = torch.rand(1, 3, 32, 64) * 2 - 1
    with pytest.raises(AssertionError):
        pipe(rgb, empty_text_embed=empty_text_embed(), num_inference_steps=10, num_intermediate_images=4)
    with pytest.raises(ValueError):
        pipe(rgb, empty_text_embed=empty_text_embed(), num_inference_steps=4, num_intermediate_images=4,
             cond_channels="bogus")
    out = pipe(rgb, empty_text_embed=empty_text_embed(), num_inference_steps=4, num_intermediate_images=1,
               cond_channels="rgb", generator=torch.Generator().manual_seed(0))
    assert out.images.shape == (1, 1, 32, 64) and out.images.is_cuda and float(out.images.abs().max()) <= 1.0

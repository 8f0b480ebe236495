"""Host-side logic of the drop-in scheduler classes (no GPU): tables, timesteps, coefficients, error behaviour —
checked against the golden vectors made by the reference's own classes and against the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

from d3roma_b200.scheduler_ddim import MyDDIMScheduler
from d3roma_b200.scheduler_ddpm import MyDDPMScheduler
from oracle import schedulers_ref as S

G = Path(__file__).parent / "golden"
from tests.test_oracle_schedulers import DDIM, DDPM  # noqa: E402
from tests._golden import SOURCES, load, same  # noqa: E402


def _ddim(c):
    kw = dict(num_train_timesteps=c["T"], beta_schedule=c["sched"], prediction_type=c["pt"], clip_sample=c["clip"],
              thresholding=c["thr"], set_alpha_to_one=c["a1"], steps_offset=c["off"],
              timestep_spacing=c.get("spacing", "leading"), clip_sample_range=c.get("crange", 1.0))
    if "bs" in c:
        kw.update(beta_start=c["bs"], beta_end=c["be"])
    return MyDDIMScheduler.from_config(kw)


@pytest.mark.parametrize("name", list(DDIM))
def test_ddim_tables_and_timesteps(name):
    z = np.load(G / "ddim.npz")
    for n, eta in ((10, 0.0), (8, 0.5)):
        s = _ddim(DDIM[name])
        s.set_timesteps(n)
        key = f"{name}/n{n}_eta{eta}"
        assert np.array_equal(s.timesteps.numpy(), z[f"{key}/timesteps"])
        assert np.array_equal(s.alphas_cumprod.numpy(), z[f"{key}/alphas_cumprod"])
        assert np.array_equal(s.final_alpha_cumprod.numpy(), z[f"{key}/final_alpha_cumprod"])
        # coefficients agree with the oracle's fp32 scalars
        ac = s.alphas_cumprod
        for t in s.timesteps.tolist():
            c = s.step_coeffs(t, eta)
            prev_t = t - DDIM[name]["T"] // n
            a_prev = ac[prev_t] if prev_t >= 0 else s.final_alpha_cumprod
            assert c.sqrt_alpha_t == float(ac[t] ** 0.5) and c.sqrt_alpha_prev == float(a_prev ** 0.5)


@pytest.mark.parametrize("source", SOURCES)
@pytest.mark.parametrize("name", list(DDPM))
def test_ddpm_tables_and_timesteps(name, source):
    z = load("ddpm", source)
    c = DDPM[name]
    s = MyDDPMScheduler.from_config(dict(num_train_timesteps=c["T"], beta_schedule=c["sched"],
                                         prediction_type=c["pt"], clip_sample=c["clip"], thresholding=c["thr"],
                                         variance_type=c["vt"]))
    assert np.array_equal(s.betas.numpy(), z[f"{name}/betas"])
    assert np.array_equal(s.alphas_cumprod.numpy(), z[f"{name}/alphas_cumprod"])
    assert np.array_equal(np.asarray(s.posterior_variance), z[f"{name}/posterior_variance"])
    for n in (8, 100):
        s.set_timesteps(n)
        assert np.array_equal(s.timesteps.numpy(), z[f"{name}/n{n}/timesteps"])
    a, b, t = (torch.from_numpy(z[f"{name}/an_{k}"]) for k in "abt")
    assert same(s.add_noise(a, b, t).numpy(), z[f"{name}/add_noise"], source)
    assert same(s.get_velocity(a, b, t).numpy(), z[f"{name}/velocity"], source)
    assert same(s.posterior_mean_variance(a, b, t)[0].numpy(), z[f"{name}/posterior_mean"], source)


def test_survey_appendix_c4_variances():
    s = MyDDPMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2"))
    s.set_timesteps(8)
    assert s.timesteps.tolist() == [112, 96, 80, 64, 48, 32, 16, 0]
    got = [float(s._get_variance(t)) for t in (112, 96, 80)]
    np.testing.assert_allclose(got, [0.67662752, 0.43904093, 0.28689373], rtol=1e-6)
    s.set_timesteps(100)
    assert s.timesteps.tolist() == list(range(99, -1, -1))


def test_reference_error_behaviour():
    s = MyDDIMScheduler(num_train_timesteps=128)
    with pytest.raises(ValueError):
        s.set_timesteps(129)
    with pytest.raises(ValueError):  # set_timesteps not called (scheduler_ddim.py:396-399)
        s.step(torch.zeros(1), 3, torch.zeros(1))
    s = MyDDIMScheduler(num_train_timesteps=128, prediction_type="bogus")
    s.set_timesteps(8)
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), 16, torch.zeros(1))
    s = MyDDIMScheduler(num_train_timesteps=128)
    s.set_timesteps(8)
    with pytest.raises(ValueError):  # both generator and variance_noise (scheduler_ddim.py:475-479)
        s.step(torch.zeros(1), 16, torch.zeros(1), eta=0.5, generator=torch.Generator(), variance_noise=torch.zeros(1))
    with pytest.raises(NotImplementedError):
        MyDDIMScheduler(beta_schedule="sigmoid")
    d = MyDDPMScheduler(num_train_timesteps=128)
    with pytest.raises(ValueError):
        d.set_timesteps(8, timesteps=[5, 3])
    with pytest.raises(ValueError):
        d.set_timesteps(timesteps=[3, 5])
    with pytest.raises(ValueError):
        d.set_timesteps(timesteps=[128, 5])
    assert len(d) == 128 and d.config.variance_type == "fixed_small"
    # unknown config keys are ignored (config.py:206 passes skip_prk_steps)
    assert MyDDIMScheduler.from_config(dict(num_train_timesteps=50, skip_prk_steps=True)).config.num_train_timesteps == 50


def test_pipeline_argument_contract():
    from d3roma_b200.pipelines import COND_ORDER, GuidedDiffusionPipeline, GuidedPipelineOutput
    import inspect
    sig = inspect.signature(GuidedDiffusionPipeline.__call__)
    assert list(sig.parameters)[1:7] == ["rgb_images", "left_images", "right_images", "sim_disp", "raw_depth", "mask"]
    assert sig.parameters["num_inference_steps"].default == 128 and sig.parameters["cond_channels"].default == "rgb"
    assert set(COND_ORDER) == {"rgb", "rgb+raw", "rgb+right", "left+right+raw", "rgb+left+right", "rgb+left+right+raw"}
    out = GuidedPipelineOutput(images=torch.zeros(1))
    out.images = torch.ones(1)  # callers reassign fields (evaluate.py:103-111)
    assert out.images.item() == 1 and out["images"].item() == 1


def test_randn_tensor_threaded_list_matches_serial_draws():
    """A list of CPU generators is drawn by a thread pool into one buffer: same values as diffusers' serial loop
    (one ``torch.randn((1, ...), generator=g)`` per sample, concatenated)."""
    import torch
    from d3roma_b200.scheduler_ddim import randn_tensor
    shape = (5, 1, 24, 40)
    a = randn_tensor(shape, generator=[torch.Generator().manual_seed(100 + i) for i in range(5)])
    b = torch.cat([torch.randn((1,) + shape[1:], generator=torch.Generator().manual_seed(100 + i)) for i in range(5)])
    assert torch.equal(a, b)
    g = torch.Generator().manual_seed(3)
    c = randn_tensor(shape, generator=g)
    assert torch.equal(c, torch.randn(shape, generator=torch.Generator().manual_seed(3)))

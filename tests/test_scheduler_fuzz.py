"""Seeded fuzz of the drop-in schedulers' host logic against the reference's own classes imported live from
/root/reference (skipped where the tree is absent): random constructor configurations - beta schedules, spacings,
offsets, variance types, zero-SNR rescaling - must give the same tables, timesteps, previous timesteps and variances, and
the same exceptions for configurations the reference refuses."""
import sys

import numpy as np
import pytest
import torch

from tests._golden import HAVE_REF, REF

pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference tree absent")


@pytest.fixture(scope="module")
def ref():
    from oracle import diffusers_stub
    diffusers_stub.install()
    sys.path.insert(0, str(REF))
    try:
        from core import scheduler_ddim, scheduler_ddpm
        yield scheduler_ddim.MyDDIMScheduler, scheduler_ddpm.MyDDPMScheduler
    finally:
        sys.path.remove(str(REF))


def _build(cls, cfg):
    try:
        return cls.from_config(cfg), None
    except Exception as e:          # noqa: BLE001 - the exception type is what is compared
        return None, type(e)


def _call(fn, *args):
    """(value, None) or (None, exception type): out-of-table timesteps must fail the same way on both sides."""
    try:
        return fn(*args), None
    except Exception as e:          # noqa: BLE001
        return None, type(e)


def _random_cfg(rng, ddpm: bool):
    T = int(rng.choice([50, 128, 200, 1000]))
    cfg = dict(num_train_timesteps=T,
               beta_schedule=str(rng.choice(["linear", "scaled_linear", "squaredcos_cap_v2"])),
               beta_start=float(rng.choice([1e-4, 8.5e-4])), beta_end=float(rng.choice([0.012, 0.02])),
               prediction_type=str(rng.choice(["epsilon", "sample", "v_prediction"])),
               timestep_spacing=str(rng.choice(["leading", "trailing", "linspace"])),
               steps_offset=int(rng.choice([0, 1])), clip_sample=bool(rng.integers(2)))
    if ddpm:
        cfg["variance_type"] = str(rng.choice(["fixed_small", "fixed_small_log", "fixed_large", "fixed_large_log"]))
        if rng.random() < 0.15:
            cfg["beta_schedule"] = "sigmoid"
    else:
        cfg["set_alpha_to_one"] = bool(rng.integers(2))
        cfg["rescale_betas_zero_snr"] = bool(rng.random() < 0.25)
    return cfg, T


@pytest.mark.parametrize("seed", range(6))
def test_ddim_host_logic_matches_reference(ref, seed):
    from d3roma_b200 import MyDDIMScheduler
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        cfg, T = _random_cfg(rng, ddpm=False)
        ours, e1 = _build(MyDDIMScheduler, cfg)
        theirs, e2 = _build(ref[0], cfg)
        assert e1 == e2, (cfg, e1, e2)
        if ours is None:
            continue
        assert torch.equal(ours.alphas_cumprod, theirs.alphas_cumprod), cfg
        assert torch.equal(torch.as_tensor(ours.final_alpha_cumprod), torch.as_tensor(theirs.final_alpha_cumprod)), cfg
        assert float(ours.init_noise_sigma) == float(theirs.init_noise_sigma)
        n = int(rng.choice([1, 5, 10, 25, T, T + 1]))
        r1 = r2 = None
        try:
            ours.set_timesteps(n)
        except Exception as e:      # noqa: BLE001
            r1 = type(e)
        try:
            theirs.set_timesteps(n)
        except Exception as e:      # noqa: BLE001
            r2 = type(e)
        assert r1 == r2, (cfg, n, r1, r2)
        if r1 is not None:
            continue
        assert np.array_equal(np.asarray(ours.timesteps), np.asarray(theirs.timesteps)), (cfg, n)
        for t in np.asarray(ours.timesteps).tolist():
            prev = t - T // n
            (v1, x1), (v2, x2) = _call(ours._get_variance, t, prev), _call(theirs._get_variance, t, prev)
            assert x1 == x2, (cfg, n, t, x1, x2)
            if x1 is None:
                assert float(v1) == float(v2), (cfg, n, t)


@pytest.mark.parametrize("seed", range(6))
def test_ddpm_host_logic_matches_reference(ref, seed):
    from d3roma_b200 import MyDDPMScheduler
    rng = np.random.default_rng(2000 + seed)
    for _ in range(12):
        cfg, T = _random_cfg(rng, ddpm=True)
        ours, e1 = _build(MyDDPMScheduler, cfg)
        theirs, e2 = _build(ref[1], cfg)
        assert e1 == e2, (cfg, e1, e2)
        if ours is None:
            continue
        for name in ("betas", "alphas", "alphas_cumprod"):
            assert torch.equal(torch.as_tensor(getattr(ours, name)), torch.as_tensor(getattr(theirs, name))), (cfg, name)
        n = int(rng.choice([1, 4, 10, 50, T, T + 1]))
        r1 = r2 = None
        try:
            ours.set_timesteps(n)
        except Exception as e:      # noqa: BLE001
            r1 = type(e)
        try:
            theirs.set_timesteps(n)
        except Exception as e:      # noqa: BLE001
            r2 = type(e)
        assert r1 == r2, (cfg, n, r1, r2)
        if r1 is not None:
            continue
        assert np.array_equal(np.asarray(ours.timesteps), np.asarray(theirs.timesteps)), (cfg, n)
        for t in np.asarray(ours.timesteps).tolist():
            assert int(ours.previous_timestep(t)) == int(theirs.previous_timestep(t)), (cfg, n, t)
            (v1, x1), (v2, x2) = _call(ours._get_variance, t), _call(theirs._get_variance, t)
            assert x1 == x2, (cfg, n, t, x1, x2)
            if x1 is None:
                assert float(v1) == pytest.approx(float(v2), rel=1e-6, abs=1e-12), (cfg, n, t)


# ---------------------------------------------------------------------------------------------- the step itself
def _split(c, m, s):
    """sched_math.cuh::split_x0_eps + clip_x0 in torch fp32 (one rounding per op, the kernel's order)."""
    sa, sb = torch.tensor(c.sqrt_alpha_t), torch.tensor(c.sqrt_beta_t)
    if c.prediction_type == 0:
        x0, eps = (s - sb * m) / sa, m
    elif c.prediction_type == 1:
        x0, eps = m, (s - sa * m) / sb
    else:
        x0, eps = sa * s - sb * m, sa * m + sb * s
    if c.clip_mode == 1:
        x0 = x0.clamp(-c.clip_range, c.clip_range)
    elif c.clip_mode == 2:
        x0 = x0.clamp(-1.0, 1.0)
    return x0, eps, sa, sb


def _ddim_kernel(c, m, s, noise):
    """ddim_step_kernel (csrc/scheduler.cu) restated on the coefficient struct the host logic produced."""
    x0, eps, sa, sb = _split(c, m, s)
    if c.use_clipped_model_output:
        eps = (s - sa * x0) / sb
    p = torch.tensor(c.sqrt_alpha_prev) * x0 + torch.tensor(c.dir_coeff) * eps
    return (p if noise is None else p + torch.tensor(c.std_dev) * noise), x0


def _ddpm_kernel(c, m, s, noise):
    x0, _, _, _ = _split(c, m, s)
    mean = torch.tensor(c.x0_coeff) * x0 + torch.tensor(c.xt_coeff) * s
    return (mean if noise is None else mean + torch.tensor(c.sigma) * noise), x0


@pytest.mark.parametrize("seed", range(4))
def test_ddim_step_coefficients_reproduce_reference_step(ref, seed):
    """step_coeffs (the host half of the drop-in ``step``) fed through the kernel's arithmetic gives the reference's
    ``step`` on random tensors, for random configurations, eta and use_clipped_model_output."""
    from d3roma_b200 import MyDDIMScheduler
    rng = np.random.default_rng(3000 + seed)
    g = torch.Generator().manual_seed(seed)
    checked = 0
    for _ in range(12):
        cfg, T = _random_cfg(rng, ddpm=False)
        ours, e1 = _build(MyDDIMScheduler, cfg)
        theirs, _ = _build(ref[0], cfg)
        if ours is None:
            continue
        n = int(rng.choice([5, 10, 25]))
        ours.set_timesteps(n)
        theirs.set_timesteps(n)
        eta = float(rng.choice([0.0, 0.0, 0.5]))
        clipped = bool(rng.integers(2))
        x = torch.randn(2, 1, 6, 8, generator=g)
        for t in np.asarray(ours.timesteps).tolist():
            if t >= T:
                break                                   # leading spacing + offset at n = T: out of the table on both sides
            m = torch.randn(2, 1, 6, 8, generator=g) * 0.7
            noise = torch.randn(2, 1, 6, 8, generator=g) if eta > 0 else None
            want = theirs.step(m, t, x, eta=eta, use_clipped_model_output=clipped, variance_noise=noise)
            c = ours.step_coeffs(t, eta, clipped)
            prev, x0 = _ddim_kernel(c, m, x, noise)
            assert torch.allclose(x0, want.pred_original_sample, rtol=1e-5, atol=1e-6, equal_nan=True), (cfg, t)
            assert torch.allclose(prev, want.prev_sample, rtol=1e-5, atol=2e-6, equal_nan=True), (cfg, t, eta, clipped)
            x = want.prev_sample
            checked += 1
    assert checked > 50


@pytest.mark.parametrize("seed", range(4))
def test_ddpm_step_coefficients_reproduce_reference_step(ref, seed):
    from d3roma_b200 import MyDDPMScheduler
    from d3roma_b200.scheduler_ddim import randn_tensor
    rng = np.random.default_rng(4000 + seed)
    g = torch.Generator().manual_seed(seed)
    checked = 0
    for _ in range(12):
        cfg, T = _random_cfg(rng, ddpm=True)
        cfg["thresholding"] = bool(rng.integers(2))
        ours, e1 = _build(MyDDPMScheduler, cfg)
        theirs, e2 = _build(ref[1], cfg)
        assert e1 == e2
        if ours is None:
            continue
        n = int(rng.choice([4, 10, 50]))
        ours.set_timesteps(n)
        theirs.set_timesteps(n)
        x = torch.randn(2, 1, 6, 8, generator=g)
        for t in np.asarray(ours.timesteps).tolist():
            if t >= T:
                break
            m = torch.randn(2, 1, 6, 8, generator=g) * 0.7
            want = theirs.step(m, t, x, generator=torch.Generator().manual_seed(77 + t))
            noise = randn_tensor(m.shape, generator=torch.Generator().manual_seed(77 + t), device=m.device,
                                 dtype=m.dtype) if t > 0 else None
            c = ours.step_coeffs(t)
            prev, x0 = _ddpm_kernel(c, m, x, noise)
            assert torch.allclose(x0, want.pred_original_sample, rtol=1e-5, atol=1e-6, equal_nan=True), (cfg, t)
            assert torch.allclose(prev, want.prev_sample, rtol=1e-5, atol=2e-6, equal_nan=True), (cfg, t)
            x = want.prev_sample
            checked += 1
    assert checked > 50

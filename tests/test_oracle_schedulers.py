"""Pin oracle/schedulers_ref.py against golden trajectories produced by the reference's own
MyDDIMScheduler / MyDDPMScheduler (tests/golden/ddim.npz, ddpm.npz; generator: oracle/gen_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import schedulers_ref as S
from tests._golden import SOURCES, load, same

G = Path(__file__).parent / "golden"

DDIM = {
    "cos128_sample": dict(T=128, sched="squaredcos_cap_v2", pt="sample", clip=True, thr=False, a1=False, off=1),
    "cos128_epsilon": dict(T=128, sched="squaredcos_cap_v2", pt="epsilon", clip=True, thr=False, a1=False, off=1),
    "cos128_v": dict(T=128, sched="squaredcos_cap_v2", pt="v_prediction", clip=True, thr=False, a1=False, off=1),
    "cos128_sample_thresh": dict(T=128, sched="squaredcos_cap_v2", pt="sample", clip=True, thr=True, a1=False, off=1),
    "sl1000_epsilon": dict(T=1000, sched="scaled_linear", bs=0.00085, be=0.012, pt="epsilon", clip=False, thr=False,
                           a1=False, off=1),
    "lin1000_v_trailing": dict(T=1000, sched="linear", pt="v_prediction", clip=True, thr=False, a1=True, off=0,
                               spacing="trailing"),
    "lin1000_eps_linspace": dict(T=1000, sched="linear", pt="epsilon", clip=True, thr=False, a1=True, off=0,
                                 spacing="linspace", crange=0.8),
}
DDPM = {
    "cos128_sample_thresh": dict(T=128, sched="squaredcos_cap_v2", pt="sample", clip=True, thr=True, vt="fixed_small"),
    "cos128_epsilon": dict(T=128, sched="squaredcos_cap_v2", pt="epsilon", clip=True, thr=False, vt="fixed_small"),
    "cos128_v_log": dict(T=128, sched="squaredcos_cap_v2", pt="v_prediction", clip=True, thr=False,
                         vt="fixed_small_log"),
    "lin1000_eps_large": dict(T=1000, sched="linear", pt="epsilon", clip=False, thr=False, vt="fixed_large"),
    "sig200_sample": dict(T=200, sched="sigmoid", pt="sample", clip=True, thr=False, vt="fixed_small"),
}


def test_appendix_c_tables():
    """SURVEY Appendix C.1 coefficients (produced by the reference during the survey)."""
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)
    assert list(S.make_timesteps(128, 10, "leading", 1)) == [109, 97, 85, 73, 61, 49, 37, 25, 13, 1]
    np.testing.assert_allclose(ac[[109, 97, 73, 1, 0]].numpy(),
                               [0.04726592, 0.12758839, 0.37347299, 0.99880046, 0.9995482564], rtol=2e-6)
    # Appendix C.3 one-step vectors, t = 73
    x = torch.tensor([0.04613046, 0.40240282, -1.01152909, 0.21673824, -0.61227310, 0.50361073]).view(1, 1, 2, 3)
    m = torch.tensor([0.23096681, 0.69307733, -0.26686057, 2.17846584, 0.10213730, -0.25900969]).view(1, 1, 2, 3)
    exp = {
        "epsilon": [-0.00068770, 0.30887908, -0.90508997, 0.79226935, -0.64895976, 0.54007089],
        "sample": [0.08293737, 0.48036534, -0.93613076, -0.25755358, -0.51813614, 0.39399129],
        "v_prediction": [0.01201137, 0.29720649, -0.96189702, 0.32332170, -0.62061954, 0.53595483],
    }
    for pt, e in exp.items():
        prev, _ = S.ddim_step(m, 73, x, ac, ac[0], 128, 10, prediction_type=pt)
        np.testing.assert_allclose(prev.flatten().numpy(), e, atol=2e-6)


@pytest.mark.parametrize("source", SOURCES)
@pytest.mark.parametrize("name", list(DDIM))
def test_ddim_trajectories_bit_exact(name, source):
    """Bit-exact against the reference run on this host ("live"); the committed fixtures come from another host's SIMD
    path and are matched to a few ulp where they are not identical (tests/_golden.py)."""
    z = load("ddim", source)
    c = DDIM[name]
    betas = S.make_betas(c["sched"], c["T"], c.get("bs", 1e-4), c.get("be", 0.02))
    ac = torch.cumprod(1 - betas, 0)
    final = torch.tensor(1.0) if c["a1"] else ac[0]
    for n, eta in ((10, 0.0), (8, 0.5)):
        key = f"{name}/n{n}_eta{eta}"
        ts = S.make_timesteps(c["T"], n, c.get("spacing", "leading"), c["off"])
        assert np.array_equal(ts, z[f"{key}/timesteps"])
        assert np.array_equal(ac.numpy(), z[f"{key}/alphas_cumprod"])
        x = torch.from_numpy(z[f"{key}/x0"])
        for i, t in enumerate(ts):
            m = torch.from_numpy(z[f"{key}/m{i}"])
            noise = torch.from_numpy(z[f"{key}/noise{i}"]) if eta > 0 else None
            prev, x0 = S.ddim_step(m, int(t), x, ac, final, c["T"], n, prediction_type=c["pt"],
                                   clip_sample=c["clip"], thresholding=c["thr"],
                                   clip_sample_range=c.get("crange", 1.0), eta=eta, variance_noise=noise)
            assert same(prev.numpy(), z[f"{key}/prev{i}"], source), (key, i)
            assert same(x0.numpy(), z[f"{key}/orig{i}"], source), (key, i)
            x = torch.from_numpy(np.asarray(z[f"{key}/prev{i}"]))
    a, b, t = (torch.from_numpy(z[f"{name}/an_{k}"]) for k in "abt")
    assert same(S.add_noise(a, b, t, ac).numpy(), z[f"{name}/add_noise"], source)
    assert same(S.get_velocity(a, b, t, ac).numpy(), z[f"{name}/velocity"], source)


@pytest.mark.parametrize("source", SOURCES)
@pytest.mark.parametrize("name", list(DDPM))
def test_ddpm_trajectories_bit_exact(name, source):
    z = load("ddpm", source)
    c = DDPM[name]
    betas = S.make_betas(c["sched"], c["T"])
    ac = torch.cumprod(1 - betas, 0)
    assert np.array_equal(betas.numpy(), z[f"{name}/betas"])
    assert np.array_equal(ac.numpy(), z[f"{name}/alphas_cumprod"])
    for n in (8, 100):
        key = f"{name}/n{n}"
        ts = S.make_timesteps(c["T"], n, "leading", 0)
        assert np.array_equal(ts, z[f"{key}/timesteps"])
        for i in z[f"{key}/step_ids"]:
            t = int(ts[i])
            x = torch.from_numpy(z[f"{key}/x{i}"])
            m = torch.from_numpy(z[f"{key}/m{i}"])
            noise = torch.randn(m.shape, generator=torch.Generator().manual_seed(1000 + int(i)))
            prev, x0, mean = S.ddpm_step(m, t, x, ac, c["T"], n, noise, prediction_type=c["pt"],
                                         clip_sample=c["clip"], thresholding=c["thr"], variance_type=c["vt"])
            assert same(x0.numpy(), z[f"{key}/orig{i}"], source), (key, i)
            assert same(mean.numpy(), z[f"{key}/mean{i}"], source), (key, i)
            assert same(prev.numpy(), z[f"{key}/prev{i}"], source), (key, i)

"""Timestep samplers (d3roma_b200/resample.py) against the reference's own module (core/resample.py, imported live when
/root/reference is present), stand-alone properties, and the cross-rank loss exchange over gloo (world size 2)."""
import sys

import numpy as np
import pytest
import torch

from d3roma_b200 import resample as R
from tests._golden import HAVE_REF, REF


def _ref():
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from core import resample
    return resample


@pytest.mark.skipif(not HAVE_REF, reason="reference tree absent")
def test_samplers_match_reference_draw_for_draw():
    ref = _ref()
    rng = np.random.default_rng(0)
    snr = rng.random(128) + 0.1
    for name, arg in (("uniform", 128), ("snr", snr), ("loss-second-moment", 128)):
        a, b = R.create_named_schedule_sampler(name, arg), ref.create_named_schedule_sampler(name, arg)
        if name == "loss-second-moment":            # 25 losses per timestep on average, uneven: some rows wrap, in order
            ts = rng.integers(0, 128, 3200).tolist() + list(range(128)) * 10
            losses = rng.random(len(ts)).tolist()
            a.update_with_all_losses(ts, losses)
            b.update_with_all_losses(ts, losses)
            assert a._warmed_up() and b._warmed_up()
            for t in (0, 17, 127):
                np.testing.assert_array_equal(a.history(t), b._loss_history[t])
            np.testing.assert_allclose(a.weights(), b.weights(), rtol=1e-14)
        np.random.seed(5)
        ta, wa = a.sample(64, "cpu")
        np.random.seed(5)
        tb, wb = b.sample(64, "cpu")
        assert torch.equal(ta, tb) and ta.dtype == torch.int64
        assert torch.allclose(wa, wb, rtol=1e-6) and wa.dtype == torch.float32
    with pytest.raises(NotImplementedError):
        R.create_named_schedule_sampler("nope", 10)


def test_sampler_properties():
    u = R.create_named_schedule_sampler("uniform", 50)
    np.random.seed(1)
    t, w = u.sample(1000, "cpu")
    assert int(t.min()) >= 0 and int(t.max()) < 50 and torch.equal(w, torch.ones(1000))
    s = R.LossSecondMomentResampler(8, history_per_term=3, uniform_prob=0.01)
    assert np.array_equal(s.weights(), np.ones(8))                         # not warmed up: uniform
    s.update_with_all_losses([0, 0, 0, 0], [1.0, 2.0, 3.0, 4.0])           # wraps: keeps 2, 3, 4
    np.testing.assert_array_equal(s.history(0), [2.0, 3.0, 4.0])
    for t in range(1, 8):
        s.update_with_all_losses([t] * 3, [0.5] * 3)
    w = s.weights()
    assert abs(w.sum() - 1.0) < 1e-12 and w[0] == w.max() and (w[1:] == w[1]).all()
    # importance weights keep the expectation: E_p[w_t f(t)] = mean_t f(t)
    np.random.seed(2)
    t, iw = s.sample(200000, "cpu")
    f = torch.arange(8, dtype=torch.float32) ** 2
    assert abs(float((iw * f[t]).mean()) - float(f.mean())) < 0.05 * float(f.mean())


def _worker(rank, world, port, q):
    import os
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from d3roma_b200 import resample
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = resample.LossSecondMomentResampler(16, history_per_term=2)
    g = torch.Generator().manual_seed(rank)
    for step in range(3):
        n = 5 + 3 * rank + step                      # ranks contribute different batch sizes
        ts = torch.randint(0, 16, (n,), generator=g)
        s.update_with_local_losses(ts, torch.rand(n, generator=g))
    q.put((rank, s._loss_counts.copy(), [np.sort(s.history(t)) for t in range(16)]))
    dist.barrier()
    dist.destroy_process_group()


def test_update_with_local_losses_over_gloo():
    from tests._util import run_ranks
    got = dict((r, (c, h)) for r, c, h in run_ranks(_worker, 2))
    # what one process would hold after the same updates in rank order
    want = R.LossSecondMomentResampler(16, history_per_term=2)
    gens = [torch.Generator().manual_seed(r) for r in range(2)]
    for step in range(3):
        ts, ls = [], []
        for r in range(2):
            n = 5 + 3 * r + step
            ts += torch.randint(0, 16, (n,), generator=gens[r]).tolist()
            ls += torch.rand(n, generator=gens[r]).tolist()
        want.update_with_all_losses(ts, ls)
    for r in range(2):
        np.testing.assert_array_equal(got[r][0], want._loss_counts)
        for t in range(16):
            if want._loss_counts[t]:
                np.testing.assert_allclose(got[r][1][t], np.sort(want.history(t)), rtol=1e-7)

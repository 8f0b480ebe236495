"""Host-side logic of the training step (no GPU): ``d3roma_b200.training.UNetTrainer`` driven over torch emulations of
the C-ABI ops (tests/_emulated_train_ops.py) must reproduce autograd on the oracle UNet (oracle/train_ref.py) - tape
order, gradient accumulation over skip / residual / shortcut branches, transposed and sub-pixel-phase weight packs of the
data gradients, the parameter layout, gradient clipping and AdamW."""
import copy

import numpy as np
import pytest
import torch

from oracle import schedulers_ref as S, train_ref
from oracle.unet_ref import make_unet
from tests import _emulated_train_ops as E


@pytest.fixture()
def trainer_mod(monkeypatch):
    from d3roma_b200 import training
    for name in dir(E):
        if not name.startswith("_") and callable(getattr(E, name)) and name not in ("math",):
            monkeypatch.setattr(training.ops, name, getattr(E, name), raising=False)
    return training


def _inputs(B, H, W, dc, cc, seed=0):
    g = torch.Generator().manual_seed(seed)
    clean = torch.rand(B, dc, H, W, generator=g) * 2 - 1
    cond = torch.rand(B, cc, H, W, generator=g) * 2 - 1
    noise = torch.randn(B, dc, H, W, generator=g)
    t = torch.randint(0, 128, (B,), generator=g)
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.2).float()
    w = torch.rand(B, generator=g) + 0.5
    return clean, cond, noise, t, mask, w


@pytest.mark.parametrize("dc,cin,loss_type", [(1, 8, "mse"), (3, 12, "l1")])
def test_backward_matches_autograd(trainer_mod, dc, cin, loss_type):
    torch.manual_seed(0)
    B, H, W = 2, 32, 64
    ref = make_unet(cin, dc, seed=0, sample_size=(H, W))
    clean, cond, noise, t, mask, w = _inputs(B, H, W, dc, cin - dc)
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)
    noisy = S.add_noise(clean, noise, t, ac)
    loss_ref, out_ref, grads_ref = train_ref.loss_and_grads(ref, noisy, cond, t, clean, mask, w, loss_type)

    tr = trainer_mod.UNetTrainer(in_channels=cin, out_channels=dc, loss_type=loss_type, device="cpu")
    tr._adt = torch.float32          # the emulated ops compute in the dtype of their inputs
    tr.load_state_dict(ref.state_dict())
    # the layout round-trips
    sd = tr.state_dict()
    for k, v in ref.state_dict().items():
        assert torch.equal(sd[k], v), k
    tr.zero_grad()
    pred = tr.forward(noisy, cond, t)
    assert float((pred - out_ref).abs().max()) < 2e-4
    dy, loss_sum = E.masked_loss_grad(pred, clean, mask, w, l1=loss_type == "l1", grad_scale=1.0 / pred.numel(),
                                      dbias=tr.G("out.bias"))
    assert abs(float((loss_sum * w.double()).sum() / pred.numel()) - float(loss_ref)) < 1e-5
    tr.backward(dy, bias_grad_done=True)
    got = tr.grad_dict()
    assert set(got) == set(grads_ref)
    worst = ("", 0.0)
    gmax = max(float(g.abs().max()) for g in grads_ref.values())
    for k, g in grads_ref.items():
        # per-tensor relative error; tensors whose gradient is analytically zero (to_k.bias: softmax is invariant to a
        # constant added to every score of a row) are measured against the global gradient scale instead of noise
        scale = max(float(g.abs().max()), 1e-6 * gmax)
        err = float((got[k] - g).abs().max()) / scale
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < 2e-4, worst


def test_train_steps_match_adamw(trainer_mod):
    torch.manual_seed(0)
    B, H, W, dc, cin = 2, 32, 32, 1, 8
    ref = make_unet(cin, dc, seed=1, sample_size=(H, W))
    tr = trainer_mod.UNetTrainer(in_channels=cin, out_channels=dc, lr=1e-3, max_grad_norm=0.5, device="cpu")
    tr._adt = torch.float32
    tr.load_state_dict(ref.state_dict())
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)

    class Sched:     # the two calls train_step makes on the noise scheduler
        @staticmethod
        def add_noise(a, b, t):
            return S.add_noise(a, b, t, ac)

        @staticmethod
        def get_velocity(a, b, t):
            return S.get_velocity(a, b, t, ac)

    for step, pt in enumerate(("sample", "epsilon", "v_prediction")):
        clean, cond, noise, t, mask, w = _inputs(B, H, W, dc, cin - dc, seed=10 + step)
        l_ref, gn_ref = train_ref.train_step(ref, opt, clean, cond, noise, t, ac, mask, w, pt, "mse", 0.5)
        l = tr.train_step(clean, cond, noise, t, Sched, mask=mask, weights=w, prediction_type=pt)
        assert abs(float(l) - float(l_ref)) < 2e-4 * max(1.0, abs(float(l_ref))), (step, float(l), float(l_ref))
        assert abs(tr.grad_norm - float(gn_ref)) < 1e-2 * float(gn_ref)
    sd = tr.state_dict()
    worst = 0.0
    for k, v in ref.state_dict().items():
        worst = max(worst, float((sd[k] - v).abs().max()))
    # three AdamW steps of size lr = 1e-3 each: parameters must agree far below one step
    assert worst < 2e-4, worst

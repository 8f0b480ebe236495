"""Host-side logic of the training step (no GPU): ``d3roma_b200.training.UNetTrainer`` driven over torch emulations of
the C-ABI ops (tests/_emulated_train_ops.py) must reproduce autograd on the oracle UNet (oracle/train_ref.py) - tape
order, gradient accumulation over skip / residual / shortcut branches, transposed and sub-pixel-phase weight packs of the
data gradients, the parameter layout, gradient clipping and AdamW."""
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


# ---------------------------------------------------------------------------------------------- data parallel (gloo)
def _ddp_worker(rank, world, port, q):
    import os
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from d3roma_b200 import training
    for name in dir(E):
        if not name.startswith("_") and callable(getattr(E, name)) and name != "math":
            setattr(training.ops, name, getattr(E, name))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = _run_steps(training, lambda x: x[rank::world], world)
    q.put({k: v.numpy() for k, v in out.items()})      # every rank reports (ranks hold identical parameters)
    dist.barrier()
    dist.destroy_process_group()


def _run_steps(training, shard, world):
    B, H, W, dc, cin = 2, 32, 32, 1, 8
    ref = make_unet(cin, dc, seed=2, sample_size=(H, W))
    tr = training.UNetTrainer(in_channels=cin, out_channels=dc, lr=1e-3, max_grad_norm=0.3, device="cpu")
    tr._adt = torch.float32
    tr.load_state_dict(ref.state_dict())
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)

    class Sched:
        add_noise = staticmethod(lambda a, b, t: S.add_noise(a, b, t, ac))
        get_velocity = staticmethod(lambda a, b, t: S.get_velocity(a, b, t, ac))

    for step in range(2):
        clean, cond, noise, t, mask, w = (shard(x) for x in _inputs(B, H, W, dc, cin - dc, seed=20 + step))
        tr.train_step(clean, cond, noise, t, Sched, mask=mask, weights=w, prediction_type="epsilon", world_size=world)
    # with more than one rank the up path's gradients are all-reduced while the down path's backward still runs
    assert tr.overlapped_reductions == (2 if world > 1 else 0)
    sd = tr.state_dict()
    keys = ["conv_in.weight", "mid_block.attentions.0.to_q.weight", "up_blocks.5.resnets.2.conv2.weight", "up_blocks.0.resnets.0.conv1.weight",
            "up_blocks.1.attentions.1.to_out.0.weight", "up_blocks.3.resnets.1.conv_shortcut.weight",
            "down_blocks.2.resnets.0.conv_shortcut.weight", "time_embedding.linear_2.bias", "conv_out.bias"]
    return {k: sd[k] for k in keys}


def test_data_parallel_matches_global_batch(trainer_mod):
    """world_size 2 over gloo: each rank takes every second sample, gradients are all-reduced and averaged before
    clipping / AdamW - the parameters after two steps equal those of one process training on the whole batch."""
    from tests._util import run_ranks
    want = _run_steps(trainer_mod, lambda x: x, 1)
    got = run_ranks(_ddp_worker, 2, timeout=600)[0]
    for k, v in want.items():
        assert float(np.abs(got[k] - v.numpy()).max()) < 2e-5, k


# ---------------------------------------------------------------------------------- train_step with the reference's signature
class _Sampler:
    """Stand-in for core/resample.py's samplers: a fixed draw with non-uniform importance weights."""

    def sample(self, batch_size, device):
        g = torch.Generator().manual_seed(77)
        return torch.randint(0, 128, (batch_size,), generator=g).to(device), (torch.rand(batch_size, generator=g) + 0.5).to(device)


def test_train_step_reference_signature(trainer_mod):
    """``training.train_step(accelerator, config, model, optimizer, lr_scheduler, noise_scheduler, t_sampler, step_counter,
    normalized_disp=..., ...)`` against the restated front end + step of train.py:157-538: out-of-range targets (clamp +
    mask update), a 40 x 48 crop padded to 64 x 64 (replicate / zero), cond_channels order, weighted loss, the returned
    per-sample losses."""
    import inspect
    from types import SimpleNamespace
    sig = list(inspect.signature(trainer_mod.train_step).parameters)
    assert sig[:8] == ["accelerator", "config", "model", "optimizer", "lr_scheduler", "noise_scheduler", "t_sampler",
                       "step_counter"]                                                         # train.py:157
    assert sig[8:14] == ["normalized_disp", "normalized_rgb", "mask", "left_image", "right_image", "sim_disp"]
    B, H, W = 2, 40, 48
    for cond_channels, cin, pt, lt in (("left+right+raw", 8, "sample", "mse"), ("rgb+raw", 5, "epsilon", "l1")):
        cfg = SimpleNamespace(ldm=False, loss_type=lt, noise_strategy="randn", cond_channels=cond_channels,
                              prediction_type=pt, noise_rgb=cin == 5, depth_channels=1, clip_sample=True,
                              clip_sample_range=1.0, divis_by=32, clip_grad_norm=0.7)
        ref = make_unet(cin, 1, seed=4, sample_size=(64, 64))
        tr = trainer_mod.UNetTrainer(in_channels=cin, out_channels=1, lr=1e-3, loss_type=lt, device="cpu")
        tr._adt = torch.float32
        tr.load_state_dict(ref.state_dict())
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-3)
        ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)

        class Sched:
            add_noise = staticmethod(lambda a, b, t: S.add_noise(a, b, t, ac))
            get_velocity = staticmethod(lambda a, b, t: S.get_velocity(a, b, t, ac))

        g = torch.Generator().manual_seed(31)
        disp = torch.rand(B, 1, H, W, generator=g) * 2.4 - 1.2          # partly outside [-1, 1]
        rgb, left, right = (torch.rand(B, 3, H, W, generator=g) * 2 - 1 for _ in range(3))
        raw = torch.rand(B, 1, H, W, generator=g) * 2 - 1
        mask = (torch.rand(B, 1, H, W, generator=g) > 0.2).float()
        kw = dict(normalized_disp=disp, normalized_rgb=rgb, mask=mask, left_image=left, right_image=right, sim_disp=raw)
        torch.manual_seed(123)
        l_ref, per_ref, gn_ref, t_ref = train_ref.train_step_full(cfg, ref, opt, ac, _Sampler(), **{k: v.clone() for k, v in kw.items()})
        torch.manual_seed(123)
        loss, info = trainer_mod.train_step(None, cfg, tr, None, None, Sched, _Sampler(), None, **kw)
        assert abs(loss - float(l_ref)) < 2e-4 * max(1.0, abs(float(l_ref))), (loss, float(l_ref))
        assert torch.equal(info["t"], t_ref) and not info["abnormal_loss_detected"]
        assert float((info["losses"]["loss"] - per_ref).abs().max()) < 1e-4
        assert abs(info["gradn"] - float(gn_ref)) < 1e-2 * float(gn_ref)
        sd = tr.state_dict()
        # one AdamW step moves every parameter by ~lr = 1e-3 (first step: lr * sign(g)); agreement far below that
        assert max(float((sd[k] - v).abs().max()) for k, v in ref.state_dict().items()) < 3e-4
    with pytest.raises(NotImplementedError):
        trainer_mod.train_step(None, SimpleNamespace(ldm=True), tr, None, None, Sched, _Sampler(), None, **kw)


def test_pyramid_noise_matches_reference(trainer_mod):
    """``pyramid_noise_like`` (utils_d3roma/utils.py:122-133): the same random draws in the same order as the reference's
    own function (imported from /root/reference when present; otherwise its restatement below)."""
    import random
    import torch.nn.functional as F
    from tests._golden import HAVE_REF

    def restated(x, discount=0.9):
        b, c, w, h = x.shape
        noise = torch.randn_like(x)
        for i in range(10):
            r = random.random() * 2 + 2
            w, h = max(1, int(w / (r ** i))), max(1, int(h / (r ** i)))
            noise += F.interpolate(torch.randn(b, c, w, h).to(x), size=x.shape[-2:], mode="bilinear") * discount ** i
            if w == 1 or h == 1:
                break
        return noise / noise.std()

    fn = restated
    if HAVE_REF:
        import sys
        from oracle import diffusers_stub
        diffusers_stub.install()
        sys.path.insert(0, "/root/reference")
        try:
            from utils_d3roma.utils import pyramid_noise_like as fn   # the reference's own
        finally:
            sys.path.remove("/root/reference")
    x = torch.zeros(2, 1, 64, 96)
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        random.seed(seed)
        want = fn(x)
        torch.manual_seed(seed)
        random.seed(seed)
        got = trainer_mod.pyramid_noise_like(x)
        assert float((got - want).abs().max()) < 1e-5
        assert abs(float(got.std()) - 1.0) < 1e-5


def test_gradient_accumulation_over_micro_batches(trainer_mod):
    """``accelerator.accumulate`` (train.py:281): two forward / backward passes without ``zero_grad`` accumulate into the
    flat gradient buffer - equal to autograd accumulating the same two losses."""
    B, H, W, dc, cin = 1, 32, 32, 1, 8
    ref = make_unet(cin, dc, seed=6, sample_size=(H, W))
    tr = trainer_mod.UNetTrainer(in_channels=cin, out_channels=dc, device="cpu")
    tr._adt = torch.float32
    tr.load_state_dict(ref.state_dict())
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)
    ref.zero_grad(set_to_none=True)
    tr.zero_grad()
    for seed in (50, 51):
        clean, cond, noise, t, mask, w = _inputs(B, H, W, dc, cin - dc, seed=seed)
        noisy = S.add_noise(clean, noise, t, ac)
        out = ref(torch.cat([noisy, cond], 1), t)
        out = out.sample if hasattr(out, "sample") else out
        ((out * mask - clean * mask) ** 2).mean().backward()           # grads accumulate in .grad
        pred = tr.forward(noisy, cond, t)
        dy, _ = E.masked_loss_grad(pred, clean, mask, None, grad_scale=1.0 / pred.numel(), dbias=tr.G("out.bias"))
        tr.backward(dy, bias_grad_done=True)
    got = tr.grad_dict()
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters())
    for k, p in ref.named_parameters():
        scale = max(float(p.grad.abs().max()), 1e-6 * gmax)
        assert float((got[k] - p.grad).abs().max()) / scale < 2e-4, k

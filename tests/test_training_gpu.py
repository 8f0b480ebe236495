"""GPU parity of the training-step kernels: every backward / optimizer kernel against the torch restatement
of the same op (tests/_emulated_train_ops.py, autograd where a derivative is involved), then the whole backward pass and
whole optimisation steps of ``UNetTrainer`` against autograd / AdamW on the oracle UNet (oracle/train_ref.py)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16]


def _ops():
    from d3roma_b200 import ops
    return ops


def _rel(got, want):
    return float((got.float() - want.float()).abs().max()) / (float(want.float().abs().max()) + 1e-20)


def _rand(shape, dtype, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(dtype)


def _quad_stats(x):
    """fp64 [B, C/4, 2] sum / sum of squares per 4-channel quad, as the conv epilogues produce them."""
    B, H, W, C = x.shape
    v = x.double().view(B, H * W, C // 4, 4)
    return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()


WGRAD_CASES = [
    # B, H, W, Cin, Cout, ksize, stride
    (2, 16, 24, 128, 128, 3, 1),
    (1, 8, 8, 64, 128, 1, 1),
    (2, 16, 16, 256, 128, 3, 1),
    (2, 16, 16, 128, 256, 3, 2),
    (1, 24, 40, 384, 256, 1, 1),
    (2, 12, 20, 512, 512, 3, 1),
    (3, 32, 32, 128, 128, 3, 1),
    (2, 6, 10, 256, 256, 3, 1),
    (1, 16, 16, 64, 128, 3, 1),      # halo form with an odd number of (kw, channel block) groups: last work item half full
    (2, 8, 24, 192, 128, 3, 1),
    (1, 16, 16, 64, 256, 3, 2),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", WGRAD_CASES)
def test_conv_wgrad(case, dtype):
    from tests import _emulated_train_ops as E
    ops = _ops()
    B, H, W, Cin, Cout, k, s = case
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _rand((B, H, W, Cin), dtype, g)
    dy = _rand((B, H // s, W // s, Cout), dtype, g)
    K = k * k * Cin
    dw = torch.zeros((Cout, K + 64), dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, dw, ksize=k, stride=s, col0=64)
    want = torch.zeros_like(dw)
    E.conv2d_wgrad(x, dy, want, ksize=k, stride=s, col0=64)
    assert float(dw[:, :64].abs().max()) == 0.0
    err = _rel(dw, want)
    assert err < 2e-3, (case, err)
    # accumulation: a second call adds
    ops.conv2d_wgrad(x, dy, dw, ksize=k, stride=s, col0=64)
    assert _rel(dw, 2 * want) < 2e-3


def test_conv_wgrad_pitched_operands():
    """x and dy as channel slices of wider buffers (the concat halves of the up blocks)."""
    from tests import _emulated_train_ops as E
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    xb = _rand((2, 16, 16, 384), torch.bfloat16, g)
    dyb = _rand((2, 16, 16, 256), torch.bfloat16, g)
    x, dy = xb[..., 128:384], dyb[..., :128]
    dw = torch.zeros((128, 256), dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, dw, ksize=1)
    want = torch.zeros_like(dw)
    E.conv2d_wgrad(x.contiguous(), dy.contiguous(), want, ksize=1)
    assert _rel(dw, want) < 2e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C1,C2,silu", [(128, 0, True), (256, 128, True), (512, 0, False), (512, 512, True), (256, 0, True)])
def test_groupnorm_silu_bwd(C1, C2, silu, dtype):
    from tests import _emulated_train_ops as E
    ops = _ops()
    B, H, W = 2, 12, 20
    g = torch.Generator(device="cuda").manual_seed(3)
    x1 = _rand((B, H, W, C1), dtype, g) + 0.3
    x2 = (_rand((B, H, W, C2), dtype, g, 2.0) - 0.5) if C2 else None
    C = C1 + C2
    gamma = torch.rand(C, generator=g, device="cuda") + 0.5
    beta = torch.randn(C, generator=g, device="cuda") * 0.3
    dy = _rand((B, H, W, C), dtype, g)
    addends = [_rand((B, H, W, C), dtype, g) for _ in range(3 if C1 == 128 else (1 if C2 == 0 else 0))]
    st1, st2 = _quad_stats(x1), (_quad_stats(x2) if C2 else None)
    dga, dbe = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    Cs = C1                      # channel sums over the first source only (the concat case of the up blocks)
    cs = torch.zeros((B, Cs + 8), device="cuda")
    dx = ops.groupnorm_silu_bwd(x1, st1, gamma, beta, dy, x2=x2, stats2=st2, silu=silu, addends=addends, dgamma=dga,
                                dbeta=dbe, colsum=cs[:, 8:])
    wga, wbe, wcs = torch.zeros_like(dga), torch.zeros_like(dbe), torch.zeros((B, Cs), device="cuda")
    want = E.groupnorm_silu_bwd(x1.float(), None, gamma, beta, dy.float(), x2=None if x2 is None else x2.float(), silu=silu,
                                addends=[t.float() for t in addends], dgamma=wga, dbeta=wbe, colsum=wcs)
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3
    assert _rel(dx, want) < tol, _rel(dx, want)
    assert _rel(dga, wga) < 5e-3 and _rel(dbe, wbe) < 5e-3, (_rel(dga, wga), _rel(dbe, wbe))
    assert _rel(cs[:, 8:], wcs) < 2e-2
    assert float(cs[:, :8].abs().max()) == 0.0


def test_colsum_add_sumpool():
    from tests import _emulated_train_ops as E
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(4)
    for dtype in DTYPES:
        x = _rand((3, 10, 12, 1536), dtype, g)
        out2, out1 = torch.zeros((3, 1536), device="cuda"), torch.zeros(1536, device="cuda")
        ops.colsum_nhwc16(x, out2)
        ops.colsum_nhwc16(x, out1)
        assert _rel(out2, x.float().sum((1, 2))) < 1e-4 and _rel(out1, x.float().sum((0, 1, 2))) < 1e-4
        a, b = _rand((2, 8, 8, 384), dtype, g), _rand((2, 8, 8, 384), dtype, g)
        assert torch.equal(ops.add_nhwc16(a[..., 128:], b[..., :256]), E.add_nhwc16(a[..., 128:], b[..., :256]))
        u = _rand((2, 16, 24, 128), dtype, g)
        ad = _rand((2, 8, 12, 128), dtype, g)
        assert _rel(ops.sumpool2x_nhwc16(u, addend=ad), E.sumpool2x_nhwc16(u, addend=ad)) < 1e-2
        assert torch.equal(ops.sumpool2x_nhwc16(ops.upsample_nearest2x(ad)).float(), (4 * ad.float()).to(dtype).float())


@pytest.mark.parametrize("dc", [1, 3])
def test_conv_out_bwd(dc):
    from tests import _emulated_train_ops as E
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W = 2, 20, 28
    h = _rand((B, H, W, 128), torch.bfloat16, g) + 0.2
    gamma = torch.rand(128, generator=g, device="cuda") + 0.5
    beta = torch.randn(128, generator=g, device="cuda") * 0.3
    w = torch.randn((dc, 9, 128), generator=g, device="cuda") * 0.1
    dy = torch.randn((B, dc, H, W), generator=g, device="cuda")
    dw = torch.zeros_like(w)
    da = ops.conv_out_bwd(h, _quad_stats(h), gamma, beta, w, dy, dw)
    wdw = torch.zeros_like(w)
    wda = E.conv_out_bwd(h.float(), None, gamma, beta, w, dy, wdw)
    assert _rel(da, wda) < 1e-2, _rel(da, wda)
    assert _rel(dw, wdw) < 2e-3, _rel(dw, wdw)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw", [(12, 20), (6, 10), (17, 30)])
def test_attention_hd8_bwd(hw, dtype):
    from tests import _emulated_train_ops as E
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(6)
    B, C = 2, 512
    qkv = _rand((B, *hw, 3 * C), dtype, g)
    d_o = _rand((B, *hw, C), dtype, g)
    o = ops.attention_hd8(qkv, C)
    dqkv = ops.attention_hd8_bwd(qkv, o, d_o)
    want = E.attention_hd8_bwd(qkv.float(), None, d_o.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 5e-3
    assert _rel(dqkv, want) < tol, _rel(dqkv, want)


def test_loss_linear_optimizer_kernels():
    from tests import _emulated_train_ops as E
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    B, dc, H, W = 3, 3, 20, 24
    pred, tgt = torch.randn((B, dc, H, W), generator=g, device="cuda"), torch.randn((B, dc, H, W), generator=g, device="cuda")
    mask = (torch.rand((B, 1, H, W), generator=g, device="cuda") > 0.3).float()
    wts = torch.rand(B, generator=g, device="cuda") + 0.5
    for l1 in (False, True):
        db, wdb = torch.zeros(dc, device="cuda"), torch.zeros(dc, device="cuda")
        dy, ls = ops.masked_loss_grad(pred, tgt, mask, wts, l1=l1, grad_scale=0.37, dbias=db)
        wdy, wls = E.masked_loss_grad(pred, tgt, mask, wts, l1=l1, grad_scale=0.37, dbias=wdb)
        assert _rel(dy, wdy) < 1e-6 and _rel(ls, wls) < 1e-6 and _rel(db, wdb) < 1e-4
    dy, ls = ops.masked_loss_grad(pred, tgt, None, None)
    assert _rel(dy, 2 * (pred - tgt)) < 1e-6
    # small Linear backward
    x, w = torch.randn((5, 512), generator=g, device="cuda"), torch.randn((640, 512), generator=g, device="cuda") * 0.05
    dyl = torch.randn((5, 640), generator=g, device="cuda")
    for act in (0, 1):
        dw, db = torch.zeros_like(w), torch.zeros(640, device="cuda")
        wdw, wdb = torch.zeros_like(w), torch.zeros(640, device="cuda")
        dx = ops.linear_small_bwd(x, w, dyl, act_in=act, dw=dw, db=db)
        wdx = E.linear_small_bwd(x, w, dyl, act_in=act, dw=wdw, db=wdb)
        assert _rel(dx, wdx) < 1e-4 and _rel(dw, wdw) < 1e-4 and _rel(db, wdb) < 1e-5
    # AdamW + clipping + 16-bit mirror
    n = 100_000
    p, gr = torch.randn(n, generator=g, device="cuda"), torch.randn(n, generator=g, device="cuda") * 3
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    p16, p16b = torch.zeros(4096, dtype=torch.bfloat16, device="cuda"), torch.zeros(4096, dtype=torch.bfloat16, device="cuda")
    for step in (1, 2, 3):
        ss, norm = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(1, device="cuda")
        ops.sumsq(gr, ss)
        kw = dict(lr=1e-2, step=step, sumsq_buf=ss, inv_loss_scale=0.5, max_norm=1.0, n16=4096)
        ops.adamw_step(p, gr, m, v, p16=p16, norm_out=norm, **kw)
        E.adamw_step(p2, gr, m2, v2, p16=p16b, **kw)
        assert abs(float(norm) - float(gr.double().norm()) * 0.5) < 1e-3 * float(norm)
    assert _rel(p, p2) < 1e-5 and _rel(v, v2) < 1e-4 and torch.equal(p16, p[:4096].to(torch.bfloat16))
    # transposed weight packs
    wm = torch.randn((256, 9 * 128 + 64), generator=g, device="cuda")
    for tm in ([8 - j for j in range(9)], [-1, 4, 5, -1], [0]):
        out = torch.zeros((128, len(tm) * 256), dtype=torch.float16, device="cuda")
        want = torch.zeros_like(out)
        ops.pack_taps16(wm, 64, 128, tm, out)
        E.pack_taps16(wm, 64, 128, tm, want)
        assert torch.equal(out, want)


def _setup(dc, cin, B, H, W, dtype, seed=0, loss_scale=1.0, **kw):
    import d3roma_b200.training as T
    from oracle.unet_ref import make_unet
    ref = make_unet(cin, dc, seed=seed, sample_size=(H, W)).cuda()
    tr = T.UNetTrainer(in_channels=cin, out_channels=dc, compute_dtype=dtype, loss_scale=loss_scale, device="cuda", **kw)
    tr.load_state_dict(ref.state_dict())
    return ref, tr


@pytest.mark.parametrize("dtype,loss_scale,dc,cin,loss_type", [("bf16", 1.0, 1, 8, "mse"), ("fp16", 4096.0, 1, 8, "mse"),
                                                                ("bf16", 1.0, 3, 12, "l1"), ("bf16", 1.0, 1, 5, "mse")])
def test_backward_matches_autograd(dtype, loss_scale, dc, cin, loss_type):
    """left+right+raw (8 -> 1), the SceneFlow variant (12 -> 3, L1 loss) and rgb+raw (5 -> 1)."""
    from oracle import schedulers_ref as S, train_ref
    from tests.test_training_host import _inputs
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, H, W = 2, 64, 96
    ref, tr = _setup(dc, cin, B, H, W, dtype, loss_scale=loss_scale, loss_type=loss_type)
    clean, cond, noise, t, mask, w = (x.cuda() for x in _inputs(B, H, W, dc, cin - dc))
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0).cuda()
    noisy = S.add_noise(clean, noise, t, ac)
    loss_ref, out_ref, grads_ref = train_ref.loss_and_grads(ref, noisy, cond, t, clean, mask, w, loss_type)
    ops = _ops()
    tr.zero_grad()
    pred = tr.forward(noisy, cond, t)
    assert float((pred - out_ref).abs().max()) < (2e-2 if dtype == "bf16" else 1e-2)
    dy, loss_sum = ops.masked_loss_grad(pred, clean, mask, w, l1=loss_type == "l1", grad_scale=loss_scale / pred.numel(),
                                        dbias=tr.G("out.bias"))
    tr.backward(dy, bias_grad_done=True)
    got = tr.grad_dict()
    gmax = max(float(g.abs().max()) for g in grads_ref.values())
    num = den = 0.0
    worst = ("", 0.0)
    for k, g in grads_ref.items():
        a = got[k].cuda().double().flatten()
        b = g.double().flatten()
        num += float(((a - b) ** 2).sum())
        den += float((b ** 2).sum())
        if float(b.abs().max()) > 1e-4 * gmax:
            cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
            if 1 - cos > worst[1]:
                worst = (k, 1 - cos)
    rel_l2 = math.sqrt(num / den)
    print(f"[{dtype} {cin}->{dc} {loss_type}] whole-gradient relative L2 error {rel_l2:.3e}; worst per-tensor 1-cos {worst}")
    # L1: the gradient of |e| is a sign - a 16-bit forward flips it where |e| is at the rounding level
    lim = (6e-2 if dtype == "bf16" else 2e-2) * (3.0 if loss_type == "l1" else 1.0)
    assert rel_l2 < lim, rel_l2
    assert worst[1] < lim / 2, worst


def test_backward_full_resolution():
    """One sample at the bench resolution (384 x 640: 48 x 80 weight-gradient tiles per sample, every stride-2 level,
    2400-token attention at level 4 is NOT part of this net - its attention runs at 24 x 40) against fp32 autograd."""
    _full_resolution_case(384, 640)


def test_backward_sceneflow_resolution():
    """544 x 960 (BASELINE configs[3] shape): 17 x 30 maps at the deepest level (partial 8 x 8 weight-gradient tiles),
    2040-token attention backward."""
    _full_resolution_case(544, 960)


def _full_resolution_case(H, W):
    from oracle import schedulers_ref as S, train_ref
    from tests.test_training_host import _inputs
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, dc, cin = 1, 1, 8
    ref, tr = _setup(dc, cin, B, H, W, "bf16")
    clean, cond, noise, t, mask, w = (x.cuda() for x in _inputs(B, H, W, dc, cin - dc, seed=9))
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0).cuda()
    noisy = S.add_noise(clean, noise, t, ac)
    _, out_ref, grads_ref = train_ref.loss_and_grads(ref, noisy, cond, t, clean, mask, w, "mse")
    tr.zero_grad()
    pred = tr.forward(noisy, cond, t)
    dy, _ = _ops().masked_loss_grad(pred, clean, mask, w, grad_scale=1.0 / pred.numel(), dbias=tr.G("out.bias"))
    tr.backward(dy, bias_grad_done=True)
    got = tr.grad_dict()
    num = sum(float(((got[k].cuda().double() - g.double()) ** 2).sum()) for k, g in grads_ref.items())
    den = sum(float((g.double() ** 2).sum()) for g in grads_ref.values())
    rel = math.sqrt(num / den)
    print(f"[bf16, 1 x {H} x {W}] whole-gradient relative L2 error {rel:.3e}")
    assert rel < 6e-2, rel


def test_train_steps_track_adamw_oracle():
    """Five optimisation steps on a fixed batch: the loss falls and follows the fp32 AdamW oracle."""
    from oracle import schedulers_ref as S, train_ref
    from tests.test_training_host import _inputs
    B, H, W, dc, cin = 2, 32, 64, 1, 8
    ref, tr = _setup(dc, cin, B, H, W, "bf16", lr=2e-4, max_grad_norm=1.0)
    opt = torch.optim.AdamW(ref.parameters(), lr=2e-4)
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0).cuda()

    class Sched:
        add_noise = staticmethod(lambda a, b, t: S.add_noise(a, b, t, ac))
        get_velocity = staticmethod(lambda a, b, t: S.get_velocity(a, b, t, ac))

    clean, cond, noise, t, mask, w = (x.cuda() for x in _inputs(B, H, W, dc, cin - dc, seed=3))
    ours, theirs = [], []
    for _ in range(5):
        theirs.append(float(train_ref.train_step(ref, opt, clean, cond, noise, t, ac, mask, w, "sample", "mse", 1.0)[0]))
        ours.append(float(tr.train_step(clean, cond, noise, t, Sched, mask=mask, weights=w, prediction_type="sample")))
    print("loss ours", ours, "oracle", theirs)
    assert ours[-1] < 0.7 * ours[0]
    for a, b in zip(ours, theirs):
        assert abs(a - b) < 0.1 * abs(b) + 1e-3, (ours, theirs)


def test_nonfinite_gradients_skip_the_step_and_weights_hand_over_to_inference():
    """(1) An overflowed gradient buffer must not touch parameters or moments (GradScaler semantics).  (2) After real
    steps, ``trainer.state_dict()`` loads into the inference ``UNet2DModel`` and both forwards agree (train.py:640 runs
    the validation pipeline on the weights being trained)."""
    import d3roma_b200 as d
    from oracle import schedulers_ref as S
    from tests.test_training_host import _inputs
    B, H, W, dc, cin = 2, 32, 64, 1, 8
    ref, tr = _setup(dc, cin, B, H, W, "bf16", lr=1e-3)
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0).cuda()

    class Sched:
        add_noise = staticmethod(lambda a, b, t: S.add_noise(a, b, t, ac))
        get_velocity = staticmethod(lambda a, b, t: S.get_velocity(a, b, t, ac))

    clean, cond, noise, t, mask, w = (x.cuda() for x in _inputs(B, H, W, dc, cin - dc, seed=5))
    tr.train_step(clean, cond, noise, t, Sched, mask=mask, weights=w)
    p0, m0 = tr._p.clone(), tr._m.clone()
    tr._g.fill_(float("inf"))
    tr.optimizer_step()
    assert torch.equal(tr._p, p0) and torch.equal(tr._m, m0) and not math.isfinite(tr.grad_norm)
    tr.train_step(clean, cond, noise, t, Sched, mask=mask, weights=w)
    assert math.isfinite(tr.grad_norm) and not torch.equal(tr._p, p0)
    # hand-over: the trained weights through the inference model
    net = d.UNet2DModel(sample_size=(H, W), in_channels=cin, out_channels=dc, compute_dtype="bf16").to("cuda")
    net.load_state_dict(tr.state_dict())
    x = torch.cat([clean, cond], 1)
    want = tr.forward(clean, cond, torch.full((B,), 37, device="cuda"))
    got = net(x, 37).sample
    # two bf16 implementations of the same network (the inference forward fuses differently): each is within ~1e-2 of
    # the fp32 result
    assert float((got - want).abs().max()) < 6e-2, float((got - want).abs().max())
    assert float((got - want).abs().mean()) < 6e-3, float((got - want).abs().mean())


def test_train_step_reference_signature_on_gpu():
    """The drop-in ``training.train_step`` (reference signature) on the device: zero / replicate padding and the
    out-of-range clamp kernels against torch, then a few steps on one batch (loss falls; returned fields)."""
    import torch.nn.functional as F
    from types import SimpleNamespace
    import d3roma_b200 as d
    from d3roma_b200 import training
    from tests.test_training_host import _Sampler
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((2, 3, 37, 50), generator=g, device="cuda")
    assert torch.equal(ops.pad_zero(x, [7, 7, 13, 14]), F.pad(x, (7, 7, 13, 14)))
    disp = torch.rand((2, 3, 20, 24), generator=g, device="cuda") * 2.4 - 1.2
    m = (torch.rand((2, 1, 20, 24), generator=g, device="cuda") > 0.3).float()
    want_x = disp.clamp(-1, 1)
    want_m = m * (want_x.abs() < 1).all(dim=1, keepdim=True).float()
    got_x, got_m = ops.clamp_target_(disp.clone(), m.clone(), 1.0)
    assert torch.equal(got_x, want_x) and torch.equal(got_m, want_m)

    B, H, W = 2, 40, 48
    cfg = SimpleNamespace(ldm=False, loss_type="mse", noise_strategy="randn", cond_channels="left+right+raw",
                          prediction_type="sample", noise_rgb=False, depth_channels=1, clip_sample=True,
                          clip_sample_range=1.0, divis_by=32, clip_grad_norm=1.0)
    ref, tr = _setup(1, 8, B, 64, 64, "bf16", lr=3e-4)
    sched = d.MyDDPMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2",
                                               prediction_type="sample", clip_sample=True))
    gc = torch.Generator().manual_seed(3)
    disp = (torch.rand(B, 1, H, W, generator=gc) * 2.2 - 1.1).cuda()
    left, right = ((torch.rand(B, 3, H, W, generator=gc) * 2 - 1).cuda() for _ in range(2))
    raw = (torch.rand(B, 1, H, W, generator=gc) * 2 - 1).cuda()
    mask = (torch.rand(B, 1, H, W, generator=gc) > 0.2).float().cuda()
    losses = []
    for _ in range(4):
        torch.manual_seed(5)        # the same noise draw every step: the loss of THIS batch must fall
        loss, info = training.train_step(None, cfg, tr, None, None, sched, _Sampler(), None, normalized_disp=disp,
                                         mask=mask, left_image=left, right_image=right, sim_disp=raw)
        losses.append(loss)
        assert info["t"].shape == (B,) and info["losses"]["loss"].shape == (B,) and math.isfinite(info["gradn"])
    assert losses[-1] < 0.8 * losses[0], losses


def test_pyramid_noise_kernels():
    """Bilinear up-sample + accumulate and the division by the standard deviation against torch on the same draws."""
    import random
    import torch.nn.functional as F
    from d3roma_b200 import training
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    for (h, w), (H, W) in (((7, 11), (64, 96)), ((1, 3), (384, 640)), ((96, 160), (384, 640)), ((5, 1), (33, 47))):
        src = torch.randn((2, 3, h, w), generator=g, device="cuda")
        dst = torch.randn((2, 3, H, W), generator=g, device="cuda")
        want = dst + 0.81 * F.interpolate(src, size=(H, W), mode="bilinear")
        ops.bilinear_upsample_acc_(dst, src, 0.81)
        assert float((dst - want).abs().max()) < 2e-6
    x = torch.randn((4, 1, 384, 640), generator=g, device="cuda") * 3 + 0.2
    want = x / x.std()
    assert float((ops.div_by_std_(x.clone()) - want).abs().max()) < 1e-5
    z = torch.zeros(2, 1, 128, 160, device="cuda")
    torch.manual_seed(4)
    random.seed(4)
    a = training.pyramid_noise_like(z)
    assert a.shape == z.shape and abs(float(a.std()) - 1.0) < 1e-4 and bool(torch.isfinite(a).all())

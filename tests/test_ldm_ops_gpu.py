"""Kernels of the latent family (csrc/ldm_ops.cu + the clipped-output / asymmetric-padding modes of the igemm conv)
against plain PyTorch fp32 references of the same ops on the same (16-bit rounded) inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DT = [torch.float16, torch.bfloat16]


def _tol(dt):
    return 2e-3 if dt == torch.float16 else 1.6e-2


def _rnd(shape, dt, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dt)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C1,C2,hw,silu,eps", [(320, 0, (24, 40), True, 1e-5), (640, 320, (12, 20), True, 1e-5),
                                               (1280, 1280, (6, 10), True, 1e-5), (1280, 640, (7, 10), False, 1e-6),
                                               (64, 0, (9, 5), True, 1e-6), (128, 64, (16, 24), True, 1e-5)])
def test_groupnorm_generic(dt, C1, C2, hw, silu, eps):
    from d3roma_b200 import ops
    B, (H, W) = 3, hw
    x1 = _rnd((B, H, W, C1), dt, 1, 2.0) + 0.5
    x2 = _rnd((B, H, W, C2), dt, 2) if C2 else None
    C = C1 + C2
    g = torch.randn(C, device="cuda") * 0.5 + 1.0
    b = torch.randn(C, device="cuda") * 0.2
    chan = torch.zeros(B * C * 2, dtype=torch.float64, device="cuda")
    ops.gn_stats(x1, chan, x2=x2)
    cat = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref_s = cat.double().sum(dim=(1, 2))
    ref_q = (cat.double() ** 2).sum(dim=(1, 2))
    got = chan.view(B, C, 2)
    assert torch.allclose(got[..., 0], ref_s, rtol=1e-6, atol=1e-3) and torch.allclose(got[..., 1], ref_q, rtol=1e-6, atol=1e-3)
    out = ops.gn_apply(x1, chan, g, b, x2=x2, eps=eps, silu=silu)
    ref = F.group_norm(cat.float().permute(0, 3, 1, 2), 32, g, b, eps)
    if silu:
        ref = F.silu(ref)
    err = (out.float() - ref.permute(0, 2, 3, 1)).abs().max().item()
    assert err <= _tol(dt) * max(1.0, ref.abs().max().item()), err
    one = ops.groupnorm_one_pass(x1, g, b, x2=x2, eps=eps, silu=silu)      # the single-launch form
    err1 = (one.float() - ref.permute(0, 2, 3, 1)).abs().max().item()
    assert err1 <= _tol(dt) * max(1.0, ref.abs().max().item()), err1


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C", [320, 640, 1280, 64])
def test_layernorm(dt, C):
    from d3roma_b200 import ops
    x = _rnd((2, 5, 7, C), dt, 3, 3.0) + 1.0
    g = torch.randn(C, device="cuda") * 0.5 + 1.0
    b = torch.randn(C, device="cuda") * 0.2
    out = ops.layernorm(x, g, b)
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    assert (out.float() - ref).abs().max().item() <= _tol(dt) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dt", DT)
def test_geglu(dt):
    from d3roma_b200 import ops
    x = _rnd((2, 3, 5, 2 * 1280), dt, 4, 2.0)
    out = ops.geglu(x)
    h, gate = x.float().chunk(2, -1)
    ref = h * F.gelu(gate)
    assert (out.float() - ref).abs().max().item() <= _tol(dt) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,hw,heads", [(2, (12, 20), 10), (1, (24, 40), 5), (2, (6, 10), 20), (1, (7, 10), 3),
                                        (1, (1, 64), 1), (1, (5, 13), 2)])
def test_attention_hd64(dt, B, hw, heads):
    from d3roma_b200 import ops
    H, W = hw
    C = 64 * heads
    qkv = _rnd((B, H, W, 3 * C), dt, 5, 1.5)
    out = ops.attention_hd64(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads)
    q, k, v = (t.float().reshape(B, H * W, heads, 64).transpose(1, 2) for t in qkv.split(C, -1))
    ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
    ref = ref.transpose(1, 2).reshape(B, H, W, C)
    err = (out.float() - ref).abs().max().item()
    assert err <= (3e-3 if dt == torch.float16 else 2e-2) * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,hw,heads", [(2, (12, 20), 10), (1, (24, 40), 5), (2, (6, 10), 20), (1, (7, 10), 3),
                                        (1, (1, 128), 1), (1, (5, 13), 2), (2, (45, 80), 5), (1, (16, 16), 1)])
def test_attention_hd64_tcgen05(dt, B, hw, heads):
    """The tcgen05 / TMEM / TMA formulation against the same fp32 reference (and against the mma.sync kernel)."""
    from d3roma_b200 import ops
    H, W = hw
    C = 64 * heads
    qkv = _rnd((B, H, W, 3 * C), dt, 5, 1.5)
    out = ops.attention_hd64_tc(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads)
    q, k, v = (t.float().reshape(B, H * W, heads, 64).transpose(1, 2) for t in qkv.split(C, -1))
    ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
    ref = ref.transpose(1, 2).reshape(B, H, W, C)
    err = (out.float() - ref).abs().max().item()
    assert err <= (3e-3 if dt == torch.float16 else 2e-2) * max(1.0, ref.abs().max().item()), err
    old = ops.attention_hd64(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads)
    assert (out.float() - old.float()).abs().max().item() <= (4e-3 if dt == torch.float16 else 3e-2)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("L", [2, 1, 8, 5])
def test_cross_attention_small(dt, L):
    from d3roma_b200 import ops
    B, H, W, heads = 2, 6, 10, 5
    C = 64 * heads
    q = _rnd((B, H, W, C), dt, 6)
    kc = torch.randn(B, L, C, device="cuda")
    vc = torch.randn(B, L, C, device="cuda")
    out = ops.cross_attention_small(q, kc, vc)
    qh = q.float().reshape(B, H * W, heads, 64).transpose(1, 2)
    kh, vh = (t.reshape(B, L, heads, 64).transpose(1, 2) for t in (kc, vc))
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, -1) @ vh).transpose(1, 2).reshape(B, H, W, C)
    assert (out.float() - ref).abs().max().item() <= _tol(dt) * max(1.0, ref.abs().max().item())


def test_linear_small_and_sincos():
    from d3roma_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    x = torch.randn(5, 1280, device="cuda")
    w = torch.randn(777, 1280, device="cuda") * 0.03
    b = torch.randn(777, device="cuda")
    assert torch.allclose(ops.linear_small(x, w, b), x @ w.t() + b, atol=2e-5, rtol=1e-5)
    assert torch.allclose(ops.linear_small(x, w, None, act_in=1, act_out=1), F.silu(F.silu(x) @ w.t()), atol=2e-5, rtol=1e-5)
    ctx = torch.randn(32, 1024, device="cuda")
    w2 = torch.randn(320, 1024, device="cuda") * 0.03
    assert torch.allclose(ops.linear_small(ctx, w2), ctx @ w2.t(), atol=2e-5, rtol=1e-5)
    from oracle.ldm_ref import timestep_embedding
    t = torch.tensor([0.0, 1.0, 13.0, 999.0, 500.0], device="cuda")
    got = ops.timestep_sincos(t, 320)
    ref = timestep_embedding(t.cpu(), 320)
    assert torch.allclose(got.cpu(), ref, atol=2e-4)      # sin / cos of arguments up to 999 in fp32
    assert torch.allclose(got.cpu()[:3], ref[:3], atol=2e-6)


@pytest.mark.parametrize("dt", DT)
def test_layout_glue_and_resize(dt):
    from d3roma_b200 import ops
    x = torch.randn(2, 16, 6, 10, device="cuda")
    buf = torch.zeros(2, 6, 10, 64, dtype=dt, device="cuda")
    ops.nchw_to_nhwc16(x, buf)
    assert torch.equal(buf[..., :16], x.permute(0, 2, 3, 1).to(dt)) and float(buf[..., 16:].abs().max()) == 0.0
    back = ops.nhwc16_to_nchw(buf, 16)
    assert torch.equal(back, x.to(dt).float())
    y = _rnd((2, 12, 20, 64), dt, 7)
    up = ops.resize_nearest(y, 23, 40)
    ref = F.interpolate(y.float().permute(0, 3, 1, 2), size=(23, 40), mode="nearest").permute(0, 2, 3, 1).to(dt)
    assert torch.equal(up, ref)
    ext = ops.resize_nearest(up, 23, 40, 24, 40)     # zero-extended even copy
    assert torch.equal(ext[:, :23], up) and float(ext[:, 23:].abs().max()) == 0.0


def _conv_ref(x, w, b, stride=1, padding=1):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=stride, padding=padding).permute(0, 2, 3, 1)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,hw,ks", [(320, 320, (24, 40), 3), (640, 320, (12, 20), 3), (320, 960, (1, 480), 1),
                                            (64, 320, (16, 24), 3), (320, 64, (16, 24), 3), (1920, 640, (12, 20), 3)])
def test_conv_clipped_output(dt, cin, cout, hw, ks):
    """Cout_valid < Cout: weight rows padded to a multiple of 128, output tensor with its true channel count."""
    from d3roma_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    B, (H, W) = 2, hw
    x = _rnd((B, H, W, cin), dt, 8)
    w = _rnd((cout, cin, ks, ks), dt, 9, 1.0 / math.sqrt(cin * ks * ks))
    b = torch.randn(cout, device="cuda")
    res = _rnd((B, H, W, cout), dt, 10)
    out = ops.conv2d_igemm(x, ops.pad_rows(ops.pack_conv_weight(w, dtype=dt)), ops.pad_rows(b), ksize=ks, residual=res,
                           cout_valid=cout)
    assert out.shape == (B, H, W, cout)
    ref = _conv_ref(x, w, b, padding=ks // 2) + res.float()
    err = (out.float() - ref).abs().max().item()
    assert err <= _tol(dt) * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("dt", DT)
def test_conv_clipped_output_with_folded_shortcut(dt):
    from d3roma_b200 import ops
    B, H, W = 2, 12, 20
    a = _rnd((B, H, W, 320), dt, 11)
    h = _rnd((B, H, W, 640), dt, 12)
    s = _rnd((B, H, W, 320), dt, 13)
    w = _rnd((320, 320, 3, 3), dt, 14, 0.02)
    wsc = _rnd((320, 960, 1, 1), dt, 15, 0.03)
    b = torch.randn(320, device="cuda")
    out = ops.conv2d_igemm(a, ops.pad_rows(ops.pack_conv_weight(w, wsc, dtype=dt)), ops.pad_rows(b), ksize=3, sc1=h,
                           sc2=s, cout_valid=320)
    ref = _conv_ref(a, w, b) + _conv_ref(torch.cat([h, s], -1), wsc, None, padding=0)
    assert (out.float() - ref).abs().max().item() <= _tol(dt) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("c,hw", [(128, (32, 48)), (256, (16, 24)), (512, (8, 12))])
def test_conv_stride2_asymmetric_pad(dt, c, hw):
    """Downsample2D(padding=0) of the VAE encoder: F.pad(x, (0, 1, 0, 1)) + conv3x3 stride 2."""
    from d3roma_b200 import ops
    B, (H, W) = 2, hw
    x = _rnd((B, H, W, c), dt, 16)
    w = _rnd((c, c, 3, 3), dt, 17, 1.0 / math.sqrt(9 * c))
    b = torch.randn(c, device="cuda")
    out = ops.conv2d_igemm(x, ops.pack_conv_weight(w, dtype=dt), b, ksize=3, stride=2, asym_pad=True)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xp, w.float(), b, stride=2).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert (out.float() - ref).abs().max().item() <= _tol(dt) * max(1.0, ref.abs().max().item())
    sym = ops.conv2d_igemm(x, ops.pack_conv_weight(w, dtype=dt), b, ksize=3, stride=2)
    refs = _conv_ref(x, w, b, stride=2)
    assert (sym.float() - refs).abs().max().item() <= _tol(dt) * max(1.0, refs.abs().max().item())

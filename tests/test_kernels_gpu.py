"""Component parity of the bandwidth / small kernels against plain PyTorch fp32 references of the same ops."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _quad_stats(x):  # x [B,H,W,C] bf16 -> [B, C/4, 2] fp64
    B, H, W, C = x.shape
    o = x.double().reshape(B, H * W, C // 4, 4)
    return torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1).contiguous()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("C1,C2,silu", [(128, 0, True), (256, 128, True), (512, 256, True), (512, 0, False),
                                        (1024, 0, True)])
def test_groupnorm_silu(C1, C2, silu, dt):
    from d3roma_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    B, H, W = 2, 24, 40
    x1 = (torch.randn(B, H, W, C1, device="cuda", generator=g) * 1.5 + 0.3).to(dt)
    x2 = (torch.randn(B, H, W, C2, device="cuda", generator=g) * 0.7 - 0.2).to(dt) if C2 else None
    C = C1 + C2
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    out = ops.groupnorm_silu(x1, _quad_stats(x1), gamma, beta, x2=x2, stats2=None if x2 is None else _quad_stats(x2),
                             silu=silu)
    xc = torch.cat([x1, x2], 3) if C2 else x1
    ref = F.group_norm(xc.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    tol = 1e-2 if dt == torch.bfloat16 else 2e-3
    torch.testing.assert_close(out.float(), ref, rtol=tol, atol=tol)


def test_upsample_and_im2col():
    from d3roma_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2, 6, 10, 128, device="cuda", generator=g).to(torch.bfloat16)
    up = ops.upsample_nearest2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    for Cs, Cc in ((1, 7), (1, 4), (3, 9), (8, 0)):
        s = torch.randn(2, Cs, 8, 12, device="cuda", generator=g)
        c = torch.randn(2, Cc, 8, 12, device="cuda", generator=g) if Cc else None
        col = ops.im2col_input(s, c)
        xin = torch.cat([s, c], 1) if Cc else s
        Cin = Cs + Cc
        unf = F.unfold(xin, 3, padding=1)  # [B, Cin*9, HW] with index c*9 + tap
        unf = unf.view(2, Cin, 9, 8, 12).permute(0, 3, 4, 2, 1).reshape(2, 8, 12, 9 * Cin)
        ref = F.pad(unf, (0, 128 - 9 * Cin)).to(torch.bfloat16)
        assert torch.equal(col, ref)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("dc", [1, 3])
def test_conv_out_fused(dc, dt):
    from d3roma_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    B, H, W = 2, 10, 70
    x = (torch.randn(B, H, W, 128, device="cuda", generator=g) * 2 + 0.5).to(dt)
    gamma = torch.randn(128, device="cuda", generator=g)
    beta = torch.randn(128, device="cuda", generator=g)
    w = torch.randn(dc, 128, 3, 3, device="cuda", generator=g) / 34.0
    bias = torch.randn(dc, device="cuda", generator=g)
    out = ops.conv_out_fused(x, _quad_stats(x), gamma, beta, w.permute(0, 2, 3, 1).reshape(dc, 9, 128).contiguous(),
                             bias, dc)
    torch.backends.cudnn.allow_tf32 = False
    act = F.silu(F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
    # the kernel stages the normalised activation in the 16-bit activation type (like every other layer's GN pass)
    ref16 = F.conv2d(act.to(dt).float(), w.to(dt).float(), bias, padding=1)   # 16-bit operands, fp32 accumulate
    ref32 = F.conv2d(act, w, bias, padding=1)
    tol16 = 3e-3 if dt == torch.bfloat16 else 1e-3   # tanh.approx SiLU can flip a last-place rounding of `act`
    torch.testing.assert_close(out, ref16, rtol=tol16, atol=tol16)
    torch.testing.assert_close(out, ref32, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("N,hw", [(960, (24, 40)), (240, (12, 20)), (2040, (34, 60))])
def test_attention_hd8(N, hw):
    from d3roma_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, C = 2, 512
    qkv = (torch.randn(B, hw[0], hw[1], 3 * C, device="cuda", generator=g) * 1.2).to(torch.bfloat16)
    out = ops.attention_hd8(qkv, C)
    q, k, v = qkv.float().view(B, N, 3, 64, 8).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(8), dim=-1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B, hw[0], hw[1], C)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


def test_time_embedding():
    from d3roma_b200 import ops
    from oracle.unet_ref import TimestepEmbedding, timestep_embedding
    torch.manual_seed(4)
    te = TimestepEmbedding(128, 512)
    proj = torch.nn.Linear(512, 384)
    for tv in ([73.0], [1.0, 109.0, 990.0]):
        t = torch.tensor(tv)
        emb = te(timestep_embedding(t, 128))
        ref_semb = F.silu(emb)
        ref_proj = proj(ref_semb)
        semb = ops.time_embed(t.cuda(), te.linear_1.weight.t().contiguous().cuda(), te.linear_1.bias.cuda(),
                              te.linear_2.weight.t().contiguous().cuda(), te.linear_2.bias.cuda())
        torch.testing.assert_close(semb.cpu(), ref_semb.detach(), rtol=1e-4, atol=1e-5)
        pr = ops.temb_project_all(semb, proj.weight.detach().cuda().contiguous(), proj.bias.detach().cuda())
        torch.testing.assert_close(pr.cpu(), ref_proj.detach(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,Cs,Cc,H,W", [(2, 1, 7, 32, 64), (1, 3, 9, 17, 70), (2, 1, 4, 40, 130)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_conv_in_direct_split_matches_torch(B, Cs, Cc, H, W, dt):
    """conv_in(cat[x, cond]) = conv(x) + {conv(cond) + bias}: the per-loop fp32 conditioning term and the per-step
    sample taps against torch's fp32 conv2d over the concatenated input; statistics are those of the stored values."""
    import torch.nn.functional as F
    from d3roma_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, Cs, H, W, device="cuda", generator=g)
    c = torch.randn(B, Cc, H, W, device="cuda", generator=g)
    w = torch.randn(128, Cs + Cc, 3, 3, device="cuda", generator=g) / (9 * (Cs + Cc)) ** 0.5
    bias = torch.randn(128, device="cuda", generator=g)
    base = ops.conv_in_direct(c, ops.pack_conv_in_direct(w, Cs, Cs + Cc), bias=bias)
    ref_base = F.conv2d(c, w[:, Cs:], bias, padding=1).permute(0, 2, 3, 1)
    torch.testing.assert_close(base, ref_base, rtol=1e-5, atol=2e-5)
    stats = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
    out = ops.conv_in_direct(x, ops.pack_conv_in_direct(w, 0, Cs), base=base, out_dtype=dt, stats=stats)
    ref = F.conv2d(torch.cat([x, c], 1), w, bias, padding=1).permute(0, 2, 3, 1)
    assert out.dtype == dt and out.shape == (B, H, W, 128)
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    assert bool(((out.float() - ref).abs() <= ulp * ref.abs() + 1e-4).all())
    o = out.double().reshape(B, H * W, 32, 4)
    s_ref = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1)
    torch.testing.assert_close(stats, s_ref, rtol=1e-5, atol=1e-3)
    # the per-loop term kept in fp16 (what the UNet does: 4 instead of 6 bytes per element and step)
    base16 = ops.conv_in_direct(c, ops.pack_conv_in_direct(w, Cs, Cs + Cc), bias=bias, out_dtype=torch.float16)
    assert base16.dtype == torch.float16
    assert bool(((base16.float() - ref_base).abs() <= 2.0 ** -11 * ref_base.abs() + 1e-6).all())
    out16 = ops.conv_in_direct(x, ops.pack_conv_in_direct(w, 0, Cs), base=base16, out_dtype=dt)
    assert bool(((out16.float() - ref).abs() <= ulp * ref.abs() + 2.0 ** -11 * ref_base.abs() + 1e-4).all())


@pytest.mark.parametrize("ptype", ["sample", "epsilon", "v_prediction"])
@pytest.mark.parametrize("dc", [1, 3])
def test_conv_out_with_fused_ddim_update_is_bit_identical(ptype, dc):
    """d3_conv_out_fused_ddim (the DDIM update applied inside the conv_out gather) == d3_conv_out_fused followed by
    d3_sched_ddim_step, bit for bit, for every prediction type."""
    import d3roma_b200 as d
    from d3roma_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(13)
    B, H, W = 2, 24, 40
    x = torch.randn(B, H, W, 128, device="cuda", generator=g).to(torch.float16)
    gamma, beta = torch.rand(128, device="cuda", generator=g) + 0.5, torch.randn(128, device="cuda", generator=g) * 0.1
    w = torch.randn(dc, 9, 128, device="cuda", generator=g) / 34.0
    bias = torch.randn(dc, device="cuda", generator=g) * 0.1
    sample = torch.randn(B, dc, H, W, device="cuda", generator=g)
    sch = d.MyDDIMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2",
                                             prediction_type=ptype, clip_sample=True, set_alpha_to_one=False,
                                             steps_offset=1))
    sch.set_timesteps(10)
    for t in (109, 61, 1):
        m = ops.conv_out_fused(x, _quad_stats(x), gamma, beta, w, bias, dc)
        want = sch.step(m, t, sample)
        prev, x0 = ops.conv_out_fused(x, _quad_stats(x), gamma, beta, w, bias, dc, ddim=(sample, sch.step_coeffs(t)))
        assert torch.equal(prev, want.prev_sample) and torch.equal(x0, want.pred_original_sample)
        # the model output can be written too
        mo = torch.empty_like(m)
        ops.conv_out_fused(x, _quad_stats(x), gamma, beta, w, bias, dc, out=mo, ddim=(sample, sch.step_coeffs(t)))
        assert torch.equal(mo, m)

"""Parity of the tcgen05 implicit-GEMM conv (d3_conv2d_igemm) against a plain PyTorch fp32 conv2d
evaluated on the same bf16-rounded operands (floating-point kernel -> torch fp32 reference)."""
import pytest
import torch
import torch.nn.functional as F

from tests._util import check_conv_stats

pytestmark = pytest.mark.gpu

CASES = [
    # B, H, W, Cin, Cout, k, stride, C1, C2, residual
    (1, 16, 128, 64, 128, 3, 1, 0, 0, False),
    (2, 24, 40, 128, 128, 3, 1, 0, 0, True),      # TW=40,TH=3 tile, identity residual
    (1, 12, 20, 512, 512, 3, 1, 0, 0, True),      # tiny map, N tiles
    (2, 48, 80, 256, 256, 3, 1, 0, 0, False),     # BN=256
    (1, 96, 160, 384, 256, 3, 1, 256, 128, False),  # virtual-concat 1x1 shortcut folded in
    (2, 32, 64, 128, 128, 3, 2, 0, 0, False),     # stride-2 downsample
    (1, 34, 60, 256, 256, 3, 2, 0, 0, False),     # stride-2, ragged tiles (544x960 family)
    (2, 24, 40, 512, 1536, 1, 1, 0, 0, False),    # 1x1 (attention QKV linear)
    (1, 17, 30, 512, 512, 1, 1, 0, 0, True),      # 1x1 + residual (attention out-proj), odd sizes
    (1, 384, 640, 128, 128, 3, 1, 0, 0, False),   # the dominant shape
    (1, 48, 64, 256, 128, 3, 1, 128, 128, False),  # Cout=128 with folded shortcut (up5 resnets), pixels-as-N kernel
    (2, 33, 70, 64, 128, 3, 1, 0, 0, True),        # ragged tiles + residual, pixels-as-N kernel
    (1, 1, 20, 128, 128, 1, 1, 0, 0, True),        # single-row image (TH = 1 path)
    (3, 192, 320, 128, 128, 3, 1, 0, 0, True),     # halo-block variant, identity residual
    (1, 8, 320, 256, 128, 3, 1, 128, 128, False),  # halo-block variant with folded 1x1 shortcut slabs
    (2, 5, 200, 64, 128, 3, 1, 0, 0, True),        # ragged width
    (1, 6, 960, 128, 128, 3, 1, 0, 0, False),      # 544x960 family
    (2, 272, 480, 128, 128, 3, 2, 0, 0, False),    # regression: >1 tile per CTA with a tile whose last accumulator
    (40, 24, 40, 128, 128, 3, 1, 0, 0, True),      #   window used to run past TMEM column 512 (2nd stage)
]


def _ref(x, w, b, stride, sc, wsc, res):
    xf = x.float().permute(0, 3, 1, 2)
    k = w.shape[-1]
    y = F.conv2d(xf, w.float(), b, stride=stride, padding=k // 2)
    if sc is not None:
        y = y + F.conv2d(sc.float().permute(0, 3, 1, 2), wsc.float()[:, :, None, None])
    if res is not None:
        y = y + res.float().permute(0, 3, 1, 2)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_igemm_matches_torch(case, dt):
    from d3roma_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, Cin, Cout, k, stride, C1, C2, use_res = case
    g = torch.Generator(device="cuda").manual_seed(1234)
    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, k, k, device=dev, generator=g) / (Cin * k * k) ** 0.5).to(dt)
    bias = torch.randn(B, Cout, device=dev, generator=g)
    Ho, Wo = H // stride, W // stride
    sc1 = sc2 = sc = wsc = None
    if C1:
        sc1 = torch.randn(B, Ho, Wo, C1, device=dev, generator=g).to(dt)
        sc2 = torch.randn(B, Ho, Wo, C2, device=dev, generator=g).to(dt)
        sc = torch.cat([sc1, sc2], dim=3)
        wsc = (torch.randn(Cout, C1 + C2, device=dev, generator=g) / (C1 + C2) ** 0.5).to(dt)
    res = torch.randn(B, Ho, Wo, Cout, device=dev, generator=g).to(dt) if use_res else None
    wp = ops.pack_conv_weight(w, wsc, dtype=dt)
    stats = torch.zeros(B, Cout // 4, 2, dtype=torch.float64, device=dev)
    out = ops.conv2d_igemm(x, wp, bias, ksize=k, stride=stride, sc1=sc1, sc2=sc2, residual=res, stats=stats)
    torch.cuda.synchronize()
    ref = _ref(x, w, None, stride, sc, wsc, res) + bias[:, None, None, :]
    err = (out.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 if dt == torch.bfloat16 else 2.0 ** -10 * ref.abs() + 3e-3
    assert bool((err <= tol).all()), f"max err {err.max().item():.4g} (ref max {ref.abs().max().item():.3g})"
    assert check_conv_stats(stats, out, ref), "GroupNorm quad statistics mismatch"


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("B,H,W,C", [(2, 12, 20, 512), (1, 24, 40, 128), (2, 17, 30, 256), (1, 192, 320, 128)])
def test_upsample_conv_subpixel_phases(B, H, W, C, dt):
    """Upsample2D (nearest 2x + conv3x3) as 4 sub-pixel phase convs == conv over the materialised upsample."""
    from d3roma_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(dt)
    w = (torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5)
    bias = torch.randn(C, device="cuda", generator=g)
    out = torch.empty(B, 2 * H, 2 * W, C, device="cuda", dtype=dt)
    stats = torch.zeros(B, C // 4, 2, dtype=torch.float64, device="cuda")
    for ph, wp in enumerate(ops.pack_upsample_weights(w, dtype=dt)):
        ops.conv2d_igemm(x, wp, bias, ksize=3, out=out, stats=stats, up2x_phase=ph + 1)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.to(dt).float(), bias, padding=1).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    tol = (2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10) * ref.abs() + (3e-2 if dt == torch.bfloat16 else 4e-3)
    assert bool((err <= tol).all()), f"max err {err.max().item():.4g}"
    assert check_conv_stats(stats, out, ref.contiguous()), "GroupNorm quad statistics mismatch"


def test_kernel_variants_match():
    """Routing overrides (halo blocks with each tile width, per-tap tiles, the lane-exchange epilogue, pixels-as-M only)
    are selected through environment variables read once per process; every one of them must give correct results:
    run each in a subprocess over shapes with ragged tiles, residuals, folded shortcuts and statistics
    (tests/_conv_variant_check.py)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "_conv_variant_check.py")
    for env in ({"D3_CONV_VB": "0"}, {"D3_CONV_VB": "8"}, {"D3_CONV_VB": "16"}, {"D3_CONV_VB": "32"},
                {"D3_CONV_VB": "64"}, {"D3_CONV_VB": "64", "D3_CONV_VB_NBLK": "3"}, {"D3_CONV_EPI": "0"},
                {"D3_CONV_EPI": "0", "D3_CONV_VB": "32"}, {"D3_CONV_NO_T": "1"}, {"D3_CONV_T_MAXC": "128"}):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, script], cwd=root, env=e, capture_output=True, text=True, timeout=300)
        assert "VARIANT_OK" in r.stdout, (env, r.stdout[-800:], r.stderr[-800:])


def test_conv_fuzz_random_shapes():
    """Seeded random shapes (ragged tiles, every kernel family, several tiles per persistent CTA so that both
    accumulator stages and every ring wrap are exercised) against torch fp32 on the same rounded operands."""
    import random
    from d3roma_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    rnd = random.Random(1234)
    g = torch.Generator(device="cuda").manual_seed(99)
    n_multi = 0
    for it in range(48):
        cout = rnd.choice([128, 128, 128, 256, 512])
        cin = rnd.choice([64, 128, 256])
        k, stride = rnd.choice([(3, 1), (3, 1), (3, 1), (3, 2), (1, 1)])
        H, W = rnd.randint(3, 60) * stride, rnd.randint(3, 90) * stride
        target = rnd.choice([3_000, 30_000, 60_000, 120_000])   # output pixels: 148 x 256 = 37 888 fill one wave
        B = max(1, min(64, -(-target // ((H // stride) * (W // stride)))))
        use_sc = (k == 3 and stride == 1 and rnd.random() < 0.3)
        use_res = (not use_sc) and rnd.random() < 0.4
        dt = rnd.choice([torch.bfloat16, torch.float16])
        Ho, Wo = H // stride, W // stride
        n_multi += int(B * Ho * Wo > 148 * 256)
        x = torch.randn(B, H, W, cin, device="cuda", generator=g).to(dt)
        w = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).to(dt)
        bias = torch.randn(B, cout, device="cuda", generator=g)
        sc1 = wsc = None
        if use_sc:
            sc1 = torch.randn(B, Ho, Wo, 64, device="cuda", generator=g).to(dt)
            wsc = (torch.randn(cout, 64, device="cuda", generator=g) / 8.0).to(dt)
        res = torch.randn(B, Ho, Wo, cout, device="cuda", generator=g).to(dt) if use_res else None
        stats = torch.zeros(B, cout // 4, 2, dtype=torch.float64, device="cuda")
        out = ops.conv2d_igemm(x, ops.pack_conv_weight(w, wsc, dtype=dt), bias, ksize=k, stride=stride, sc1=sc1,
                               residual=res, stats=stats)
        torch.cuda.synchronize()
        ref = _ref(x, w, None, stride, sc1, wsc, res) + bias[:, None, None, :]
        err = (out.float() - ref).abs()
        tol = 2.0 ** -7 * ref.abs() + 2e-2 if dt == torch.bfloat16 else 2.0 ** -10 * ref.abs() + 3e-3
        case = (it, B, H, W, cin, cout, k, stride, use_sc, use_res, str(dt))
        assert bool((err <= tol).all()), (case, err.max().item())
        assert check_conv_stats(stats, out, ref), case
    assert n_multi >= 12

"""Subprocess body of test_conv_gpu.test_kernel_variants_match: 3x3 convs (ragged tiles, residual, folded 1x1
shortcut, statistics) against torch fp32 on the same rounded operands; prints VARIANT_OK / VARIANT_BAD."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3roma_b200 import ops  # noqa: E402
from tests._util import check_conv_stats  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(3)
SHAPES = [  # B, H, W, Cin, Cout, C1, C2, residual
    (2, 6, 320, 128, 128, 0, 0, True), (1, 4, 640, 64, 128, 0, 0, False), (2, 24, 40, 256, 256, 0, 0, False),
    (1, 48, 80, 256, 512, 0, 0, True), (2, 33, 70, 64, 128, 0, 0, True), (1, 64, 64, 256, 128, 128, 128, False),
    (3, 17, 9, 128, 128, 64, 0, True), (1, 96, 160, 128, 128, 0, 0, False)]
ok = True
for (B, H, W, Ci, Co, C1, C2, use_res) in SHAPES:
    dt = torch.bfloat16
    x = torch.randn(B, H, W, Ci, device="cuda", generator=g).to(dt)
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5).to(dt)
    b = torch.randn(B, Co, device="cuda", generator=g)
    sc1 = torch.randn(B, H, W, C1, device="cuda", generator=g).to(dt) if C1 else None
    sc2 = torch.randn(B, H, W, C2, device="cuda", generator=g).to(dt) if C2 else None
    wsc = (torch.randn(Co, C1 + C2, device="cuda", generator=g) / (C1 + C2) ** 0.5).to(dt) if C1 else None
    res = torch.randn(B, H, W, Co, device="cuda", generator=g).to(dt) if use_res else None
    stats = torch.zeros(B, Co // 4, 2, dtype=torch.float64, device="cuda")
    out = ops.conv2d_igemm(x, ops.pack_conv_weight(w, wsc), b, ksize=3, sc1=sc1, sc2=sc2, residual=res, stats=stats)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1)
    if C1:
        sc = torch.cat([t for t in (sc1, sc2) if t is not None], dim=3)
        ref = ref + F.conv2d(sc.float().permute(0, 3, 1, 2), wsc.float()[:, :, None, None])
    ref = ref.permute(0, 2, 3, 1) + b[:, None, None, :]
    if use_res:
        ref = ref + res.float()
    good = bool(((out.float() - ref).abs() <= 2 ** -7 * ref.abs() + 2e-2).all())
    good = good and check_conv_stats(stats, out, ref)
    if not good:
        print("BAD shape", (B, H, W, Ci, Co, C1, C2, use_res), float((out.float() - ref).abs().max()))
    ok = ok and good
print("VARIANT_OK" if ok else "VARIANT_BAD")

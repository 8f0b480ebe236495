"""TEST INFRASTRUCTURE: torch (CPU, fp32) emulations of the C-ABI ops ``d3roma_b200.training`` calls, with the same
argument conventions (NHWC activations, packed [Cout][tap*Cin + c] weights, virtual concat, sub-pixel phases).
tests/test_training_host.py swaps them in to check the HOST-side orchestration of the backward pass (tape order, gradient
accumulation, transposed weight packs, parameter layout) against autograd on a machine without a GPU.  The GPU tests
check each kernel against these same functions.  Never imported by the package; not a fallback of the product path."""
import math

import torch
import torch.nn.functional as F

from tests._emulated_ops import conv2d_igemm, linear_small, timestep_sincos  # noqa: F401


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).to(dtype).contiguous()


def groupnorm_silu(x1, stats1, gamma, beta, *, x2=None, stats2=None, silu=True, groups=32, eps=1e-5, out=None):
    cat = x1 if x2 is None else torch.cat([x1, x2], -1)
    y = F.group_norm(_nchw(cat), groups, gamma, beta, eps)
    y = F.silu(y) if silu else y
    return _nhwc(y, x1.dtype)


def groupnorm_silu_bwd(x1, stats1, gamma, beta, dy, *, x2=None, stats2=None, silu=True, groups=32, eps=1e-5, addends=(),
                       dgamma=None, dbeta=None, colsum=None):
    cat = (x1 if x2 is None else torch.cat([x1, x2], -1)).float().detach().requires_grad_(True)
    g, b = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    y = F.group_norm(cat.permute(0, 3, 1, 2), groups, g, b, eps)
    y = (F.silu(y) if silu else y).permute(0, 2, 3, 1)
    dx, dg, db = torch.autograd.grad(y, (cat, g, b), dy.float())
    for t in addends:
        if t is not None:
            dx = dx + t.float()
    dx = dx.to(x1.dtype)
    if dgamma is not None:
        dgamma += dg
        dbeta += db
    if colsum is not None:
        s = dx.float().sum(dim=(1, 2))[:, :colsum.shape[-1]]
        colsum += s if colsum.dim() == 2 else s.sum(0)
    return dx.contiguous()


def conv2d_wgrad(x, dy, dw, *, ksize, stride=1, col0=0):
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    g = torch.nn.grad.conv2d_weight(_nchw(x), (Cout, Cin, ksize, ksize), _nchw(dy), stride=stride, padding=ksize // 2)
    dw[:, col0:col0 + ksize * ksize * Cin] += g.permute(0, 2, 3, 1).reshape(Cout, -1)
    return dw


def colsum_nhwc16(x, out):
    s = x.float().sum(dim=(1, 2))
    out += s if out.dim() == 2 else s.sum(0)
    return out


def add_nhwc16(a, b, out=None, colsum=None):
    r = (a.float() + b.float()).to(a.dtype)
    if colsum is not None:
        cs = r.float().sum(dim=(1, 2))
        colsum += cs if colsum.dim() == 2 else cs.sum(0)
    if out is not None:
        out.copy_(r)
        return out
    return r.contiguous()


def upsample_nearest2x(x):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()


def sumpool2x_nhwc16(x, addend=None):
    B, H2, W2, C = x.shape
    r = x.float().view(B, H2 // 2, 2, W2 // 2, 2, C).sum(dim=(2, 4))
    if addend is not None:
        r = r + addend.float()
    return r.to(x.dtype).contiguous()


def im2col_input(sample, cond, out=None, dtype=torch.float32):
    x = sample if cond is None else torch.cat([sample, cond], 1)
    B, C, H, W = x.shape
    cols = F.unfold(x.float(), 3, padding=1).view(B, C, 9, H, W)          # [B, c, tap, H, W]
    cols = cols.permute(0, 3, 4, 2, 1).reshape(B, H, W, 9 * C)             # k = tap * C + c
    return F.pad(cols, (0, 128 - 9 * C)).to(dtype).contiguous()


def _conv_out(h, gamma, beta, w, bias, eps):
    a = F.silu(F.group_norm(h, 32, gamma, beta, eps))
    dc = w.shape[0]
    return a, F.conv2d(a, w.view(dc, 3, 3, 128).permute(0, 3, 1, 2), bias, padding=1)


def conv_out_fused(x, stats, gamma, beta, w, bias, dc, eps=1e-5, out=None, ddim=None):
    return _conv_out(_nchw(x), gamma, beta, w, bias, eps)[1].contiguous()


def conv_out_bwd(h, stats, gamma, beta, w, dy, dw, eps=1e-5):
    a = F.silu(F.group_norm(_nchw(h), 32, gamma, beta, eps)).detach().requires_grad_(True)
    wv = w.detach().clone().requires_grad_(True)
    dc = w.shape[0]
    y = F.conv2d(a, wv.view(dc, 3, 3, 128).permute(0, 3, 1, 2), None, padding=1)
    da, gw = torch.autograd.grad(y, (a, wv), dy.float())
    dw += gw
    return _nhwc(da, h.dtype)


def _attn(qkv, C):
    B, H, W, _ = qkv.shape
    heads = C // 8
    q, k, v = (qkv[..., i * C:(i + 1) * C].reshape(B, H * W, heads, 8).transpose(1, 2) for i in range(3))
    o = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(8.0), -1) @ v
    return o.transpose(1, 2).reshape(B, H, W, C)


def attention_hd8(qkv, C_):
    return _attn(qkv.float(), C_).to(qkv.dtype).contiguous()


def attention_hd8_bwd(qkv, o, d_o):
    q = qkv.float().detach().requires_grad_(True)
    y = _attn(q, qkv.shape[3] // 3)
    (g,) = torch.autograd.grad(y, q, d_o.float())
    return g.to(qkv.dtype).contiguous()


def masked_loss_grad(pred, target, mask, weights, *, l1=False, grad_scale=1.0, dbias=None):
    B = pred.shape[0]
    m = 1.0 if mask is None else mask
    w = torch.ones(B) if weights is None else weights
    e = pred * m - target * m
    if l1:
        l, g = e.abs(), torch.sign(e) * m
    else:
        l, g = e * e, 2.0 * e * m
    dy = g * grad_scale * w.view(-1, 1, 1, 1)
    if dbias is not None:
        dbias += dy.sum(dim=(0, 2, 3))
    return dy.contiguous(), l.double().sum(dim=(1, 2, 3))


def linear_small_bwd(x, w, dy, *, act_in=0, dw=None, db=None, want_dx=True):
    xin = F.silu(x) if act_in else x
    if dw is not None:
        dw += dy.t() @ xin
    if db is not None:
        db += dy.sum(0)
    if not want_dx:
        return None
    dx = dy @ w
    if act_in:
        s = torch.sigmoid(x)
        dx = dx * (s * (1 + x * (1 - s)))
    return dx


def lincomb(a, x, b=1.0, y=None, c=0.0, z=None):
    r = a * x
    if y is not None:
        r = r + b * y
    if z is not None:
        r = r + c * z
    return r


def cast_f32_to_16(x, out):
    out.copy_(x.to(out.dtype))
    return out


def pack_taps16(w, col0, Cin, tap_map, out):
    Cout = w.shape[0]
    o = out.view(Cin, len(tap_map), Cout)
    for j, t in enumerate(tap_map):
        if t < 0:
            o[:, j, :] = 0
        else:
            o[:, j, :] = w[:, col0 + t * Cin:col0 + (t + 1) * Cin].t().to(out.dtype)
    return out


def sumsq(g, out):
    out += (g.double() ** 2).sum()
    return out


def adamw_step(p, g, m, v, *, p16=None, n16=0, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step,
               sumsq_buf=None, inv_loss_scale=1.0, max_norm=0.0, norm_out=None):
    gm = inv_loss_scale
    if sumsq_buf is not None:
        norm = float(sumsq_buf.sqrt()) * inv_loss_scale
        if norm_out is not None:
            norm_out.fill_(norm)
        if max_norm > 0:
            gm *= min(1.0, max_norm / (norm + 1e-6))
    ge = g * gm
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(ge, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(ge, ge, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
    if p16 is not None:
        p16.copy_(p[:n16].to(p16.dtype))


def pad_replicate(x, pads):
    return F.pad(x, tuple(pads), mode="replicate")


def pad_zero(x, pads):
    return F.pad(x, tuple(pads), mode="constant", value=0)


def masked_minmax(x, mask=None):
    return torch.stack([x.amin(dim=1), x.amax(dim=1)], dim=1)


def clamp_target_(x, mask, clip_range):
    x.clamp_(-clip_range, clip_range)
    mask.mul_((x.abs() < clip_range).all(dim=1, keepdim=True).float())
    return x, mask


def bilinear_upsample_acc_(dst, src, mul):
    dst += mul * F.interpolate(src, size=dst.shape[-2:], mode="bilinear")
    return dst


def div_by_std_(x):
    x /= x.std()
    return x


def make_pack_table(entries, device):
    return list(entries)


def pack_taps16_batched(table, w_base, out_base):
    for wo, wp, oo, col0, cin, cout, tm in table:
        rows = cout
        w = w_base[wo:wo + rows * wp].view(rows, wp)
        pack_taps16(w, col0, cin, tm, out_base[oo:oo + cin * len(tm) * cout].view(cin, len(tm) * cout))
    return out_base

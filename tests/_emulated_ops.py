"""TEST INFRASTRUCTURE: torch (CPU, fp32) emulations of the C-ABI ops the latent-family modules call, with the same
argument conventions (NHWC activations, packed [Cout_padded][K] weights, clipped outputs, sub-pixel upsample phases,
virtual concat).  tests/test_ldm_host_logic.py swaps them in to check the HOST-side orchestration on a machine without a
GPU; they are never importable from the package and are not a fallback of the product path."""
import torch
import torch.nn.functional as F

from d3roma_b200 import ops as real_ops

pad_rows = real_ops.pad_rows
pack_conv_weight = real_ops.pack_conv_weight
pack_upsample_weights = real_ops.pack_upsample_weights


def conv2d_igemm(x, w_packed, bias, *, ksize, stride=1, sc1=None, sc2=None, residual=None, out=None, stats=None,
                 up2x_phase=0, cout_valid=0, asym_pad=False):
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    Cv = cout_valid or Cout
    taps = 4 if up2x_phase else ksize * ksize
    wm = w_packed.float()
    xin = x.float().permute(0, 3, 1, 2)
    b = bias.float()
    if up2x_phase:
        py, px = (up2x_phase - 1) // 2, (up2x_phase - 1) % 2
        w = wm[:, :4 * Cin].reshape(Cout, 2, 2, Cin).permute(0, 3, 1, 2)
        xp = F.pad(xin, (1 - px, px, 1 - py, py))       # rows {i+py-1, i+py}, cols {j+px-1, j+px}
        y = F.conv2d(xp, w)
        if b.dim() == 2:
            y = y + b[:, :, None, None]
        else:
            y = y + b[None, :, None, None]
        out[:, py::2, px::2, :] = y.permute(0, 2, 3, 1)[..., :Cv].to(out.dtype)
        return out
    w = wm[:, :taps * Cin].reshape(Cout, ksize, ksize, Cin).permute(0, 3, 1, 2)
    if asym_pad:
        y = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w, stride=2)
    else:
        y = F.conv2d(xin, w, stride=stride, padding=ksize // 2)
    off = taps * Cin
    for sc in (sc1, sc2):
        if sc is not None:
            c = sc.shape[3]
            y = y + F.conv2d(sc.float().permute(0, 3, 1, 2), wm[:, off:off + c, None, None])
            off += c
    assert off == wm.shape[1]
    y = y + (b[:, :, None, None] if b.dim() == 2 and b.shape[0] > 1 else b.reshape(1, -1, 1, 1))
    y = y.permute(0, 2, 3, 1)[..., :Cv]
    if residual is not None:
        y = y + residual.float()
    y = y.to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


def gn_stats(x1, chan, x2=None):
    cat = x1 if x2 is None else torch.cat([x1, x2], -1)
    B, C = cat.shape[0], cat.shape[-1]
    c = chan.view(B, C, 2)
    c[..., 0] += cat.double().sum(dim=(1, 2))
    c[..., 1] += (cat.double() ** 2).sum(dim=(1, 2))
    return chan


def gn_apply(x1, chan, gamma, beta, *, x2=None, groups=32, eps=1e-5, silu=True, out=None):
    cat = x1 if x2 is None else torch.cat([x1, x2], -1)
    B, H, W, C = cat.shape
    c = chan.view(B, groups, C // groups, 2).sum(2)
    n = H * W * (C // groups)
    mean = c[..., 0] / n
    var = (c[..., 1] / n - mean * mean).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + eps)
    y = (cat.double().view(B, H * W, groups, -1) - mean[:, None, :, None]) * rstd[:, None, :, None]
    y = y.view(B, H, W, C).float() * gamma + beta
    return (F.silu(y) if silu else y).to(x1.dtype)


def groupnorm_one_pass(x1, gamma, beta, *, x2=None, groups=32, eps=1e-5, silu=True):
    cat = x1 if x2 is None else torch.cat([x1, x2], -1)
    y = F.group_norm(cat.float().permute(0, 3, 1, 2), groups, gamma, beta, eps).permute(0, 2, 3, 1)
    return (F.silu(y) if silu else y).to(x1.dtype)


def layernorm(x, gamma, beta, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps).to(x.dtype)


def geglu(x):
    h, g = x.float().chunk(2, -1)
    return (h * F.gelu(g)).to(x.dtype)


def attention_hd64(q, k, v, heads):
    B, H, W, C = q.shape
    sp = lambda t: t.float().reshape(B, H * W, heads, 64).transpose(1, 2)  # noqa: E731
    o = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 8.0, -1) @ sp(v)
    return o.transpose(1, 2).reshape(B, H, W, C).to(q.dtype)


attention_hd64_tc = attention_hd64


def cross_attention_small(q, kc, vc):
    B, H, W, C = q.shape
    L, heads = kc.shape[1], C // 64
    qh = q.float().reshape(B, H * W, heads, 64).transpose(1, 2)
    kh, vh = (t.reshape(B, L, heads, 64).transpose(1, 2) for t in (kc, vc))
    o = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, -1) @ vh
    return o.transpose(1, 2).reshape(B, H, W, C).to(q.dtype)


def linear_small(x, w, bias=None, act_in=0, act_out=0):
    y = (F.silu(x) if act_in else x) @ w.t()
    if bias is not None:
        y = y + bias
    return F.silu(y) if act_out else y


def timestep_sincos(t, dim, flip_sin_to_cos=True, freq_shift=0.0):
    import math
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - freq_shift))
    a = t[:, None].float() * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)] if flip_sin_to_cos else [torch.sin(a), torch.cos(a)], -1)


def nchw_to_nhwc16(x, out, c_off=0):
    out[..., c_off:c_off + x.shape[1]] = x.permute(0, 2, 3, 1).to(out.dtype)
    return out


def nhwc16_to_nchw(x, C_, c_off=0):
    return x[..., c_off:c_off + C_].float().permute(0, 3, 1, 2).contiguous()


def resize_nearest(x, Ho, Wo, Hbuf=None, Wbuf=None):
    B, H, W, C = x.shape
    Hbuf, Wbuf = Hbuf or Ho, Wbuf or Wo
    out = x.new_zeros((B, Hbuf, Wbuf, C))
    out[:, :Ho, :Wo] = F.interpolate(x.float().permute(0, 3, 1, 2), size=(Ho, Wo), mode="nearest").permute(0, 2, 3, 1).to(x.dtype)
    return out


# ---- text encoder (d3roma_b200/text_encoder.py) ----------------------------------------------------------------------
def _act_small(x, code):
    if code == 1:
        return F.silu(x)
    if code == 2:
        return F.gelu(x)
    if code == 3:
        return x * torch.sigmoid(1.702 * x)
    return x


def linear_small(x, w, bias=None, act_in=0, act_out=0):  # noqa: F811  (supersedes the SiLU-only form above)
    y = _act_small(x, act_in) @ w.t()
    if bias is not None:
        y = y + bias
    return _act_small(y, act_out)


def embed_tokens(ids, token_embedding, position_embedding):
    B, L = ids.shape
    return (token_embedding[ids] + position_embedding[:L][None]).reshape(B * L, -1)


def layernorm_f32(x, gamma, beta, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)


def causal_attention_f32(qkv, B, L, heads, scale):
    D = qkv.shape[1] // 3
    q, k, v = (t.reshape(B, L, heads, D // heads).transpose(1, 2) for t in qkv.split(D, dim=1))
    s = q @ k.transpose(-1, -2) * scale + torch.full((L, L), float("-inf")).triu(1)
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, D)


def lincomb(a, x, b=1.0, y=None, c=0.0, z=None):
    out = a * x
    if y is not None:
        out = out + b * y
    if z is not None:
        out = out + c * z
    return out

"""``d3roma_b200.AutoencoderKL`` (latent family) against the oracle restatement of diffusers'
AutoencoderKL (oracle/ldm_ref.py, parity unpinned) with the same random-init weights in fp32 on the same GPU, through the
four attributes the reference's pipeline uses (encoder, quant_conv, post_quant_conv, decoder), plus the helper kernels of
the single-head mid-block attention."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DT = [torch.float16, torch.bfloat16]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("rows,n,pitch", [(37, 920, 1024), (5, 128, 128), (3, 60, 128)])
def test_softmax_rows_and_transpose(dt, rows, n, pitch):
    from d3roma_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    s = (torch.randn(rows, pitch, generator=g, device="cuda") * 20).to(dt)
    ref = torch.softmax(s[:, :n].float() * 0.0442, -1)
    got = ops.softmax_rows_(s.clone(), n, 0.0442)
    assert float(got[:, n:].abs().max()) == 0.0 if pitch > n else True
    assert (got[:, :n].float() - ref).abs().max().item() <= (1e-3 if dt == torch.float16 else 8e-3)
    v = torch.randn(n, 256, generator=g, device="cuda").to(dt)
    vt = ops.transpose_pad(v, pitch)
    assert torch.equal(vt[:, :n], v.t()) and (pitch == n or float(vt[:, n:].abs().max()) == 0.0)


def test_conv1x1_small():
    from d3roma_b200 import ops
    x = torch.randn(3, 8, 9, 11, device="cuda")
    w, b = torch.randn(8, 8, device="cuda"), torch.randn(8, device="cuda")
    assert torch.allclose(ops.conv1x1_nchw_small(x, w, b), F.conv2d(x, w[:, :, None, None], b), atol=1e-5)
    w4 = torch.randn(4, 4, device="cuda")
    assert torch.allclose(ops.conv1x1_nchw_small(x[:, :4].contiguous(), w4, None), F.conv2d(x[:, :4], w4[:, :, None, None]), atol=1e-5)


def _pair(cfg, seed=0, dtype="fp16"):
    import d3roma_b200.autoencoder_kl as ak
    from oracle.ldm_ref import make_vae
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = make_vae(seed=seed, **cfg)
    vae = ak.AutoencoderKL(compute_dtype=dtype, **cfg).to("cuda")
    vae.load_state_dict(ref.state_dict())
    return ref.cuda(), vae


def _cmp(got, want, tol, what):
    err = (got - want).abs().max().item()
    print(f"{what}: max-abs {err:.2e} |ref|max {want.abs().max().item():.2f}")
    assert got.shape == want.shape and got.dtype == torch.float32
    assert err <= tol * max(1.0, want.abs().max().item()), (what, err)


@pytest.mark.parametrize("cfg,hw,B", [(dict(block_out_channels=(128, 256)), (32, 48), 2),
                                      (dict(block_out_channels=(128, 256)), (40, 56), 1),      # N = 560: padded keys
                                      (dict(block_out_channels=(128, 256, 512, 512)), (64, 96), 2),
                                      (dict(block_out_channels=(128, 256, 512, 512)), (192, 320), 1)])
def test_vae_encode_decode_vs_oracle(cfg, hw, B):
    ref, vae = _pair(cfg)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(B, 3, *hw, generator=g) * 2 - 1).cuda()
    want_h = ref.encoder(x)
    got_h = vae.encoder(x)
    _cmp(got_h, want_h, 1e-2, "encoder")
    _cmp(vae.quant_conv(got_h), ref.quant_conv(want_h), 1e-2, "quant_conv")
    f = 2 ** (len(cfg["block_out_channels"]) - 1)
    z = torch.randn(B, 4, hw[0] // f, hw[1] // f, generator=g).cuda()
    want_z = ref.post_quant_conv(z)
    got_z = vae.post_quant_conv(z)
    _cmp(got_z, want_z, 1e-5, "post_quant_conv")
    _cmp(vae.decoder(got_z), ref.decoder(want_z), 1e-2, "decoder")


def test_vae_fused_entry_points_match_the_four_call_form():
    """encode_latent / decode_mean (quant conv + latent scale folded; output-channel mean folded into conv_out) against
    the reference's own call sequence (custom_pipelines.py:47-59, 255-272) on the same module."""
    ref, vae = _pair(dict(block_out_channels=(128, 256)))
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(2, 3, 32, 48, generator=g) * 2 - 1).cuda()
    mean, _ = torch.chunk(vae.quant_conv(vae.encoder(x)), 2, dim=1)
    lat = vae.encode_latent(x, 0.18215)
    assert (lat - mean * 0.18215).abs().max().item() <= 2e-3 * max(1.0, float(mean.abs().max()) * 0.18215)
    z = torch.randn(2, 4, 16, 24, generator=g).cuda() * 0.18215
    want = vae.decoder(vae.post_quant_conv(z / 0.18215)).mean(dim=1, keepdim=True)
    got = vae.decode_mean(z, 0.18215)
    assert got.shape == want.shape and (got - want).abs().max().item() <= 5e-3 * max(1.0, float(want.abs().max()))
    want_ref = ref.decoder(ref.post_quant_conv(z / 0.18215)).mean(dim=1, keepdim=True)
    assert (got - want_ref).abs().max().item() <= 1e-2 * max(1.0, float(want_ref.abs().max()))


def test_vae_contract_errors():
    import d3roma_b200.autoencoder_kl as ak
    with pytest.raises(NotImplementedError):
        ak.AutoencoderKL(block_out_channels=(64, 128))
    vae = ak.AutoencoderKL(block_out_channels=(128, 256)).to("cuda")
    with pytest.raises(RuntimeError):
        vae.encoder(torch.zeros(1, 3, 16, 16, device="cuda"))
    from oracle.ldm_ref import make_vae
    vae.load_state_dict(make_vae(block_out_channels=(128, 256)).state_dict())
    with pytest.raises(ValueError):
        vae.encoder(torch.zeros(1, 3, 16, 16))
    with pytest.raises(ValueError):
        vae.encoder(torch.zeros(1, 3, 15, 16, device="cuda"))

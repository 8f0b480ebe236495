"""The oracle restatement of the latent family's two third-party networks (oracle/ldm_ref.py; diffusers is absent from
this image and from /root/reference: parity unpinned) is gated by the PUBLISHED parameter counts and tensor counts of the
Stable-Diffusion-2.1 UNet and the SD VAE, and by the reference's own conv_in widening (train.py:752-768)."""
import torch


def test_published_parameter_counts():
    from oracle.ldm_ref import AutoencoderKL, UNet2DConditionModel
    with torch.device("meta"):
        u, v = UNet2DConditionModel(), AutoencoderKL()
    assert sum(p.numel() for p in u.parameters()) == 865_910_724      # stabilityai/stable-diffusion-2-1 unet
    assert len(u.state_dict()) == 686
    assert sum(p.numel() for p in v.parameters()) == 83_653_863       # SD VAE (AutoencoderKL)
    assert len(v.state_dict()) == 248
    with torch.device("meta"):
        u16 = UNet2DConditionModel(in_channels=16)
    assert sum(p.numel() for p in u16.parameters()) == 865_910_724 + 12 * 320 * 9


def test_widen_conv_in_matches_reference_recipe():
    from oracle.ldm_ref import make_cond_unet, widen_conv_in
    net = make_cond_unet(in_channels=4, block_out_channels=(64, 64), attention_head_dim=(1, 1), cross_attention_dim=8)
    w0, b0 = net.conv_in.weight.clone(), net.conv_in.bias.clone()
    widen_conv_in(net, 3)
    assert net.conv_in.weight.shape == (64, 12, 3, 3) and net.config.in_channels == 12
    assert torch.equal(net.conv_in.weight, w0.repeat(1, 3, 1, 1) / 3) and torch.equal(net.conv_in.bias, b0 / 3)
    # a replicated input through the widened layer reproduces the original layer (up to the bias scaling)
    x = torch.randn(1, 4, 8, 8)
    y = net.conv_in(x.repeat(1, 3, 1, 1))
    ref = torch.nn.functional.conv2d(x, w0, b0 / 3, padding=1)
    assert torch.allclose(y, ref, atol=1e-5)


def test_small_instances_run_odd_and_even_sizes():
    from oracle.ldm_ref import make_cond_unet, make_vae
    net = make_cond_unet(in_channels=8, block_out_channels=(64, 128, 128), attention_head_dim=(1, 2, 2),
                         cross_attention_dim=16)
    ctx = torch.randn(1, 2, 16)
    for hw in ((8, 12), (7, 10)):       # 7x10: diffusers' forward_upsample_size path
        y = net(torch.randn(1, 8, *hw), 5, ctx).sample
        assert y.shape == (1, 4, *hw) and torch.isfinite(y).all()
    vae = make_vae(block_out_channels=(32, 64), norm_num_groups=8)
    z = vae.quant_conv(vae.encoder(torch.randn(1, 3, 16, 24)))
    assert z.shape == (1, 8, 8, 12)
    x = vae.decoder(vae.post_quant_conv(z[:, :4]))
    assert x.shape == (1, 3, 16, 24)

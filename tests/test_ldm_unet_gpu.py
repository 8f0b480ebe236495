"""``d3roma_b200.UNet2DConditionModel`` (latent family) against the oracle restatement of diffusers'
UNet2DConditionModel (oracle/ldm_ref.py, parity unpinned - see its header) with the same random-init weights, run in
fp32 on the same GPU: reduced-width instances of the SD-2.1 block structure, even and odd latent sizes, the reference's
widened conv_in, and (slow) the full 866 M-parameter configuration."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-2   # north-star per-step bound for 16-bit arithmetic (fp16 operands)


def _pair(cfg, cin, seed=0):
    import d3roma_b200.unet_condition as uc
    from oracle.ldm_ref import make_cond_unet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = make_cond_unet(in_channels=cin, seed=seed, **cfg)
    net = uc.UNet2DConditionModel(in_channels=cin, **cfg).to("cuda")
    net.load_state_dict(ref.state_dict())
    return ref.cuda(), net


def _check(ref, net, B, hw, cin, cdim, t, L=2, tol=TOL):
    g = torch.Generator().manual_seed(hash((B, hw, cin)) % 1000)
    x = torch.randn(B, cin, *hw, generator=g).cuda()
    ctx = torch.randn(B, L, cdim, generator=g).cuda()
    want = ref(x, t, ctx).sample
    got = net(x, t, ctx).sample
    assert got.shape == want.shape and got.dtype == torch.float32
    err = (got - want).abs().max().item()
    rms = ((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item()
    print(f"hw={hw} B={B}: max-abs {err:.2e} rel-rms {rms:.2e} |ref|max {want.abs().max().item():.2f}")
    assert err <= tol * max(1.0, want.abs().max().item()), err
    return got


SMALL = dict(block_out_channels=(64, 128, 128), attention_head_dim=(1, 2, 2), cross_attention_dim=64)
SDW = dict(block_out_channels=(320, 640), attention_head_dim=(5, 10), cross_attention_dim=1024)


@pytest.mark.parametrize("hw,B", [((16, 24), 2), ((8, 8), 1), ((12, 20), 3)])
def test_small_config_even_sizes(hw, B):
    ref, net = _pair(SMALL, 8)
    _check(ref, net, B, hw, 8, 64, 37)


@pytest.mark.parametrize("hw", [(23, 40), (7, 9), (10, 15)])
def test_small_config_odd_sizes(hw):
    """diffusers' forward_upsample_size path (the released configs pad 180x320 to 184x320: 23x40 latents)."""
    ref, net = _pair(SMALL, 12)
    _check(ref, net, 2, hw, 12, 64, 501)


def test_sd_widths_two_levels_and_per_sample_timesteps():
    ref, net = _pair(SDW, 16)
    _check(ref, net, 2, (24, 40), 16, 1024, 999)
    _check(ref, net, 2, (24, 40), 16, 1024, torch.tensor([3, 750]))
    # context cache: same tensor object reused, then refilled in place
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 16, 8, 8, generator=g).cuda()
    ctx = torch.randn(1, 2, 1024, generator=g).cuda()
    a = net(x, 10, ctx).sample
    assert torch.equal(a, net(x, 10, ctx).sample)
    ctx.mul_(-1.0)
    b = net(x, 10, ctx).sample
    want = ref(x, 10, ctx).sample
    assert (b - want).abs().max().item() <= TOL * max(1.0, want.abs().max().item()) and not torch.equal(a, b)


def test_widened_conv_in_of_the_reference():
    """train.py:752-768: conv_in repeated over 4 * (k + 1) input channels."""
    import d3roma_b200.unet_condition as uc
    from oracle.ldm_ref import make_cond_unet, widen_conv_in
    ref = widen_conv_in(make_cond_unet(in_channels=4, **SMALL), 4).cuda()
    net = uc.UNet2DConditionModel(in_channels=16, **SMALL).to("cuda")
    net.load_state_dict(ref.state_dict())
    _check(ref, net, 2, (16, 16), 16, 64, 64)


def test_contract_errors():
    import d3roma_b200.unet_condition as uc
    from oracle.ldm_ref import make_cond_unet
    with pytest.raises(NotImplementedError):
        uc.UNet2DConditionModel(block_out_channels=(96, 192), attention_head_dim=(1, 2))
    net = uc.UNet2DConditionModel(in_channels=8, **SMALL).to("cuda")
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 8, 8, 8, device="cuda"), 0, torch.zeros(1, 2, 64, device="cuda"))
    net.load_state_dict(make_cond_unet(in_channels=8, **SMALL).state_dict())
    with pytest.raises(ValueError):
        net(torch.zeros(1, 8, 8, 8), 0, torch.zeros(1, 2, 64, device="cuda"))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 4, 8, 8, device="cuda"), 0, torch.zeros(1, 2, 64, device="cuda"))


def test_full_sd21_configuration():
    """The 866 M-parameter SD-2.1 layout with the reference's 16-channel conv_in (left+right+raw), 24x40 latents."""
    ref, net = _pair({}, 16)
    _check(ref, net, 1, (24, 40), 16, 1024, 321)
    del ref, net
    torch.cuda.empty_cache()


def test_cuda_graph_path_matches_eager_and_survives_shape_changes():
    ref, net = _pair(SMALL, 8)
    g = torch.Generator().manual_seed(4)
    ctx = torch.randn(2, 2, 64, generator=g).cuda()
    xs = [torch.randn(2, 8, 16, 24, generator=g).cuda(), torch.randn(2, 8, 8, 8, generator=g).cuda(),
          torch.randn(2, 8, 16, 24, generator=g).cuda()]
    assert net.use_cuda_graph
    got = [net(x, t, ctx).sample for x, t in zip(xs, (10, 20, torch.tensor([30, 40])))]
    assert len(net._graphs) == 3 and net.graph_kernel_launches > 0
    net.use_cuda_graph = False
    want = [net(x, t, ctx).sample for x, t in zip(xs, (10, 20, torch.tensor([30, 40])))]
    for a, b in zip(got, want):
        assert (a - b).abs().max().item() <= 2e-3     # fp64 atomics of the GroupNorm moments may reorder
    net.use_cuda_graph = True
    again = net(xs[0], 10, ctx).sample            # replay of the first graph after other shapes ran
    assert (again - want[0]).abs().max().item() <= 2e-3
    # reloading weights drops the graphs (they hold raw pointers into the old tensors)
    net.load_state_dict(ref.state_dict())
    assert not net._graphs

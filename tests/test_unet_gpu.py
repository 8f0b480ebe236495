"""Whole-UNet parity: d3roma_b200.UNet2DModel (tcgen05 path) vs the oracle restatement (fp32 CPU) on the same
random-init weights and inputs.

Tolerances.  BASELINE.json's north_star asks for a per-step model-output max-abs error <= 1e-2 with 16-bit tensor-core
operands.  ``tools/error_budget.py`` (run on a B200, profiles/r01_error_budget.txt) shows that on this 113 M-parameter
random-init network the bf16 rounding of the conv operands alone (weights + activations, which ANY bf16 tensor-core
implementation, cuDNN included, must perform) already yields a max-abs error of 0.85-0.92e-2 against fp32, and an ideal
bf16 pipeline (operands + bf16 storage of activations) 1.1-1.3e-2.  The product therefore runs, and bench.py TIMES,
the fp16-operand mode (UNet2DModel's default ``compute_dtype="fp16"``: same tcgen05 kind::f16 rate, 3 more mantissa
bits; the reference's own fork scripts run fp16 autocast, inference_d3roma.py:217):
  * fp16 operands/storage - the timed mode: max-abs <= 1e-2, the north_star tolerance, at every shape tested;
  * bf16 operands/storage - a secondary, non-default mode kept for range-limited checkpoints: max-abs <= 2e-2 and
    rms <= 4e-3, the bf16-inherent bound documented above (NOT claimed to meet the north_star bound).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_NORTH_STAR = 1e-2
TOL_BF16_MAX, TOL_BF16_RMS = 2e-2, 4e-3
TOL_FP32 = 1e-4   # north_star: per-step max-abs <= 1e-4 in fp32 (BASELINE configs[3], reference-precision path)


@pytest.mark.parametrize("dtype", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("cin,dc,hw,ts", [(8, 1, (64, 96), [109, 1]), (5, 1, (32, 64), [49]), (12, 3, (32, 32), [13])])
def test_unet_matches_oracle(cin, dc, hw, ts, dtype):
    from d3roma_b200.unet import UNet2DModel
    from oracle.unet_ref import make_unet
    torch.set_num_threads(8)
    ref = make_unet(cin, dc, seed=0, sample_size=hw)
    net = UNet2DModel(sample_size=hw, in_channels=cin, out_channels=dc, compute_dtype=dtype).to("cuda")
    net.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, cin, *hw, generator=g)
    for t in ts:
        with torch.no_grad():
            y_ref = ref(x, torch.tensor(t)).sample
        y = net(x.cuda(), torch.tensor(t)).sample.cpu()
        e = (y - y_ref).abs()
        err, rms = e.max().item(), e.pow(2).mean().sqrt().item()
        print(f"{dtype} cin={cin} dc={dc} hw={hw} t={t}: max-abs {err:.3e} rms {rms:.3e} "
              f"(ref absmax {y_ref.abs().max():.3f})")
        if dtype == "fp32":
            assert err <= TOL_FP32, err
        elif dtype == "fp16":
            assert err <= TOL_NORTH_STAR, err
        else:
            assert err <= TOL_BF16_MAX and rms <= TOL_BF16_RMS, (err, rms)
    # split entry (sample + cond): conv_in runs in its direct fp32 form (per-loop conditioning term + sample taps)
    # instead of the 16-bit im2col GEMM of the concatenated call - same tolerance against the oracle
    xs, xc = x[:, :dc].contiguous().cuda(), x[:, dc:].contiguous().cuda()
    y2 = net.forward_split(xs, xc, ts[-1]).cpu()
    e2 = (y2 - y_ref).abs()
    print(f"{dtype} split entry: max-abs {e2.max().item():.3e} rms {e2.pow(2).mean().sqrt().item():.3e}")
    if dtype == "fp32":   # same kernels as the concatenated call; fp32 statistics use float atomics
        torch.testing.assert_close(y2, y, rtol=0, atol=2e-5)
    elif dtype == "fp16":
        assert e2.max().item() <= TOL_NORTH_STAR
    else:
        assert e2.max().item() <= TOL_BF16_MAX and e2.pow(2).mean().sqrt().item() <= TOL_BF16_RMS
    if dtype != "fp32":
        # a second call with the same conditioning tensor reuses the cached term and is bit-identical; changing the
        # conditioning IN PLACE must invalidate the cache
        assert torch.equal(net.forward_split(xs, xc, ts[-1]).cpu(), y2)
        xc.mul_(0.5)
        y4 = net.forward_split(xs, xc, ts[-1]).cpu()
        assert not torch.equal(y4, y2)
        assert torch.equal(net.forward_split(xs, xc.clone(), ts[-1]).cpu(), y4)


@pytest.mark.parametrize("dtype,tol", [("fp32", TOL_FP32), ("fp16", TOL_NORTH_STAR)])
def test_unet_sceneflow_shape_544x960(dtype, tol):
    """BASELINE configs[3] shape: 540x960 padded to 544x960, Cin=12 / dc=3 (SceneFlow variant), one forward."""
    from d3roma_b200.unet import UNet2DModel
    from oracle.unet_ref import make_unet
    torch.set_num_threads(16)
    hw = (544, 960)
    ref = make_unet(12, 3, seed=0, sample_size=hw)
    net = UNet2DModel(sample_size=hw, in_channels=12, out_channels=3, compute_dtype=dtype).to("cuda")
    net.load_state_dict(ref.state_dict())
    x = torch.randn(1, 12, *hw, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y_ref = ref(x, torch.tensor(99)).sample
    y = net(x.cuda(), torch.tensor(99)).sample.cpu()
    err = (y - y_ref).abs().max().item()
    print(f"{dtype} 544x960 Cin=12 dc=3: max-abs {err:.3e} (ref absmax {y_ref.abs().max():.3f})")
    assert err <= tol, err


@pytest.mark.parametrize("hw,B", [((544, 960), 3), ((96, 160), 5)])
def test_unet_batch_invariance(hw, B):
    """A batched forward equals the per-sample forwards (no cross-sample arithmetic anywhere).  At
    544x960 with B = 3 every persistent conv CTA walks several tiles of every tile shape the 544x960 family produces
    (this caught a TMEM window overrun that a B = 1 run cannot reach)."""
    from d3roma_b200.unet import UNet2DModel
    from d3roma_b200.weights import random_init_state_dict
    net = UNet2DModel(sample_size=hw, in_channels=8, out_channels=1, compute_dtype="fp16").to("cuda")
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, 1, *hw, device="cuda", generator=g)
    c = torch.randn(B, 7, *hw, device="cuda", generator=g)
    y = net.forward_split(x, c, 60).clone()
    assert torch.isfinite(y).all()
    for i in range(B):
        yi = net.forward_split(x[i:i + 1].contiguous(), c[i:i + 1].contiguous(), 60)
        torch.testing.assert_close(yi, y[i:i + 1], rtol=0, atol=2e-3)


def test_graphs_survive_batch_changes_and_weight_reloads():
    """Captured graphs own their statistics arena and are dropped when the weights change: B = 1 graphed, then a larger
    batch (which used to reallocate the shared arena under the first graph), then B = 1 again must reproduce the first
    result; after load_state_dict() with other weights the graphed forward must follow the new weights."""
    from d3roma_b200.unet import UNet2DModel
    from d3roma_b200.weights import random_init_state_dict
    hw = (32, 64)
    net = UNet2DModel(sample_size=hw, in_channels=8, out_channels=1).to("cuda")
    assert net.compute_dtype == "fp16"
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    g = torch.Generator(device="cuda").manual_seed(1)
    x1, c1 = torch.randn(1, 1, *hw, device="cuda", generator=g), torch.randn(1, 7, *hw, device="cuda", generator=g)
    x6, c6 = torch.randn(6, 1, *hw, device="cuda", generator=g), torch.randn(6, 7, *hw, device="cuda", generator=g)
    y1 = net.forward_graphed(x1, c1, 50).clone()
    torch.testing.assert_close(y1, net.forward_split(x1, c1, 50), rtol=0, atol=1e-5)
    y6 = net.forward_graphed(x6, c6, 50).clone()
    net.forward_split(x6, c6, 50)                                   # eager path with the larger batch in between
    torch.cuda.empty_cache()
    assert torch.equal(net.forward_graphed(x1, c1, 50), y1)
    assert torch.equal(net.forward_graphed(x6, c6, 50), y6)
    # reload other weights on the same object: the old graphs must not be replayed
    net.load_state_dict(random_init_state_dict(8, 1, seed=1))
    y1b = net.forward_graphed(x1, c1, 50).clone()
    assert not torch.equal(y1b, y1)
    torch.testing.assert_close(y1b, net.forward_split(x1, c1, 50), rtol=0, atol=1e-5)


def test_graphed_forward_sees_in_place_conditioning_updates():
    """A caller that refills ONE preallocated conditioning buffer per frame (robot loop through sample_from) must not
    get the previous frame's conditioning from the graph's static copy."""
    from d3roma_b200.unet import UNet2DModel
    from d3roma_b200.weights import random_init_state_dict
    hw = (32, 64)
    net = UNet2DModel(sample_size=hw, in_channels=8, out_channels=1).to("cuda")
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(2, 1, *hw, device="cuda", generator=g)
    buf = torch.randn(2, 7, *hw, device="cuda", generator=g)
    ya = net.forward_graphed(x, buf, 30).clone()
    frame2 = torch.randn(2, 7, *hw, device="cuda", generator=g)
    buf.copy_(frame2)                                               # same storage, new contents
    yb = net.forward_graphed(x, buf, 30).clone()
    assert not torch.equal(ya, yb)
    assert torch.equal(yb, net.forward_graphed(x, frame2.clone(), 30))
    net.prime_conditioning(x.shape, buf)                            # no-op when nothing changed
    assert torch.equal(net.forward_graphed(x, buf, 30), yb)
    # a NEW tensor that the caching allocator places at the address of a freed one (same pointer, version 0 again)
    # must not be mistaken for the old contents
    outs = []
    for i in range(3):
        c_new = torch.full((2, 7, *hw), 0.1 * (i + 1), device="cuda")
        outs.append(net.forward_graphed(x, c_new, 30).clone())
        del c_new
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])


def test_from_pretrained_accepts_reference_style_checkpoints(tmp_path):
    """A pipeline directory as the reference / diffusers saves it: guidance/config.json with only bookkeeping keys
    (FlowGuidance registers no config, guidance.py:94), extra private keys in unet/config.json, and pre-0.14 attention
    parameter names (query / key / value / proj_attn)."""
    import json
    import d3roma_b200 as d
    from safetensors.torch import load_file, save_file
    from d3roma_b200.weights import random_init_state_dict
    hw = (32, 64)
    net = d.UNet2DModel(sample_size=hw, in_channels=8, out_channels=1).to("cuda")
    net.load_state_dict(random_init_state_dict(8, 1, seed=3))
    cfg = dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample",
               clip_sample=True, set_alpha_to_one=False, steps_offset=1)
    pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(cfg), d.FlowGuidance(flow_guidance_weight=0.0))
    pipe.save_pretrained(tmp_path)
    (tmp_path / "guidance" / "config.json").write_text(json.dumps(
        {"_class_name": "FlowGuidance", "_diffusers_version": "0.25.0", "_name_or_path": "experiments/x"}))
    ucfg = json.loads((tmp_path / "unet" / "config.json").read_text())
    ucfg.update({"_diffusers_version": "0.25.0", "_name_or_path": "experiments/x/unet", "act_fn": "silu"})
    (tmp_path / "unet" / "config.json").write_text(json.dumps(ucfg))
    st = tmp_path / "unet" / "diffusion_pytorch_model.safetensors"
    sd = load_file(st)
    old = {}
    for k, v in sd.items():
        for new_n, old_n in ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn.")):
            if ".attentions." in k and new_n in k:
                k = k.replace(new_n, old_n)
        old[k] = v
    assert any(".query." in k for k in old)
    save_file(old, st)
    pipe2 = d.GuidedDiffusionPipeline.from_pretrained(tmp_path)
    assert pipe2.guidance.flow_guidance_weight == 1.0 and pipe2.guidance.flow_guidance_mode == "imputation"
    g = torch.Generator(device="cuda").manual_seed(4)
    x, c = torch.randn(1, 1, *hw, device="cuda", generator=g), torch.randn(1, 7, *hw, device="cuda", generator=g)
    assert torch.equal(pipe2.unet.forward_split(x, c, 10), net.forward_split(x, c, 10))

"""Host-side orchestration of the latent-family modules (weight packing, block order, skip stack, virtual concat, folded
shortcuts, padded output channels, odd-size handling, context cache) checked on the CPU by swapping the C-ABI ops for
torch emulations (tests/_emulated_ops.py) and comparing with the oracle restatement.  The kernels themselves are
checked on the GPU (tests/test_ldm_ops_gpu.py, test_ldm_unet_gpu.py, test_ldm_vae_gpu.py)."""
import pytest
import torch

from tests import _emulated_ops as emu


@pytest.fixture
def emulated(monkeypatch):
    import d3roma_b200.unet_condition as uc
    monkeypatch.setattr(uc, "ops", emu)
    monkeypatch.setattr(uc, "_TEST_DEVICE", "cpu")
    return uc


SMALL = dict(block_out_channels=(64, 128, 128), attention_head_dim=(1, 2, 2), cross_attention_dim=64)


@pytest.mark.parametrize("hw,cin,B", [((8, 16), 8, 2), ((7, 10), 12, 1), ((23, 40), 16, 1)])
def test_cond_unet_orchestration(emulated, hw, cin, B):
    from oracle.ldm_ref import make_cond_unet
    ref = make_cond_unet(in_channels=cin, **SMALL)
    net = emulated.UNet2DConditionModel(in_channels=cin, compute_dtype="fp16", **SMALL)
    net._adt = torch.float32            # emulation keeps fp32 so the comparison isolates the orchestration
    net.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(0)
    x, ctx = torch.randn(B, cin, *hw, generator=g), torch.randn(B, 2, 64, generator=g)
    for t in (7, torch.arange(B) * 100 + 3):
        want = ref(x, t, ctx).sample
        got = net(x, t, ctx).sample
        assert got.shape == want.shape
        assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())


def test_cond_unet_sd_widths(emulated):
    """320 / 640-channel levels: output rows padded to 384 / 640, group widths 10 / 20 / 30."""
    from oracle.ldm_ref import make_cond_unet
    cfg = dict(block_out_channels=(320, 640), attention_head_dim=(5, 10), cross_attention_dim=32, layers_per_block=1)
    ref = make_cond_unet(in_channels=16, **cfg)
    net = emulated.UNet2DConditionModel(in_channels=16, **cfg)
    net._adt = torch.float32
    net.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(1)
    x, ctx = torch.randn(1, 16, 6, 10, generator=g), torch.randn(1, 2, 32, generator=g)
    want, got = ref(x, 500, ctx).sample, net(x, 500, ctx).sample
    assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())


def test_cond_unet_pretrained_directory_roundtrip(emulated, tmp_path):
    """save_pretrained / from_pretrained in the diffusers layout (config.json with bookkeeping and unmodelled keys,
    safetensors or torch-pickle weights); state_dict() / save need keep_state_dict=True at load time."""
    import json
    from oracle.ldm_ref import make_cond_unet
    ref = make_cond_unet(in_channels=8, **SMALL)
    net = emulated.UNet2DConditionModel(in_channels=8, compute_dtype="fp16", **SMALL)
    net.load_state_dict(ref.state_dict())
    with pytest.raises(RuntimeError):
        net.save_pretrained(tmp_path / "x")
    net.keep_state_dict = True
    net.load_state_dict(ref.state_dict())
    net.save_pretrained(tmp_path / "unet")
    cfg = json.loads((tmp_path / "unet" / "config.json").read_text())
    assert cfg["_class_name"] == "UNet2DConditionModel" and cfg["in_channels"] == 8
    cfg.update(_diffusers_version="0.24.0", act_fn="silu", center_input_sample=False, use_linear_projection=True,
               down_block_types=["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
               up_block_types=["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"])
    (tmp_path / "unet" / "config.json").write_text(json.dumps(cfg))
    again = emulated.UNet2DConditionModel.from_pretrained(tmp_path, subfolder="unet", device=None, keep_state_dict=True)
    again._adt = net._adt = torch.float32
    g = torch.Generator().manual_seed(3)
    x, ctx = torch.randn(1, 8, 8, 16, generator=g), torch.randn(1, 2, 64, generator=g)
    assert torch.equal(again(x, 11, ctx).sample, net(x, 11, ctx).sample)
    sd = again.state_dict()
    assert set(sd) == set(ref.state_dict()) and all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())
    # torch-pickle weights
    (tmp_path / "unet" / "diffusion_pytorch_model.safetensors").unlink()
    torch.save(ref.state_dict(), tmp_path / "unet" / "diffusion_pytorch_model.bin")
    emulated.UNet2DConditionModel.from_pretrained(tmp_path / "unet", device=None)
    cfg["use_linear_projection"] = False
    (tmp_path / "unet" / "config.json").write_text(json.dumps(cfg))
    with pytest.raises(NotImplementedError):
        emulated.UNet2DConditionModel.from_pretrained(tmp_path / "unet", device=None)

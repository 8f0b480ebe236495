"""A second, independent source for the arithmetic of the VAE restatement (oracle/ldm_ref.py; diffusers itself is absent
from this image): ``transformers`` carries ports of the ORIGINAL latent-diffusion / taming autoencoder code, the code the
SD VAE and diffusers' ``AutoencoderKL`` descend from -

  * ``ChameleonVQVAEEncoder``: conv_in -> per level {ResnetBlock x num_res_blocks, asymmetric-pad stride-2 downsample} ->
    mid {ResnetBlock, single-head AttnBlock, ResnetBlock} -> GroupNorm(eps 1e-6) -> swish -> conv_out(2 x latent): the SD
    VAE ENCODER for (base 128, multipliers 1-2-4-4, 2 blocks, no level attention, 4 latent channels, double_latent);
  * ``JanusVQVAEMidBlock`` / ``JanusVQVAEResnetBlock`` / ``JanusVQVAEConvUpsample``: the building blocks of the DECODER
    (its composition differs from the SD decoder in one place - attention inside the first up level - so the decoder is
    checked block by block).

The oracle's weights (diffusers key names, Linear attention projections) are mapped onto those modules' names
(1 x 1 conv projections) and both are run on the same input.  Skipped when transformers is not importable."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _map_resnet(sd, src, dst, out):
    for name in ("norm1", "conv1", "norm2", "conv2"):
        for p in ("weight", "bias"):
            out[f"{dst}.{name}.{p}"] = sd[f"{src}.{name}.{p}"]
    if f"{src}.conv_shortcut.weight" in sd:                      # diffusers' 1x1 conv_shortcut = LDM's nin_shortcut
        for p in ("weight", "bias"):
            out[f"{dst}.nin_shortcut.{p}"] = sd[f"{src}.conv_shortcut.{p}"]


def _map_attn(sd, src, dst, out):
    for p in ("weight", "bias"):
        out[f"{dst}.norm.{p}"] = sd[f"{src}.group_norm.{p}"]
    for ours, theirs in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out")):
        out[f"{dst}.{theirs}.weight"] = sd[f"{src}.{ours}.weight"][:, :, None, None]      # Linear [C, C] -> conv 1x1
        out[f"{dst}.{theirs}.bias"] = sd[f"{src}.{ours}.bias"]


@pytest.mark.parametrize("boc,hw", [((128, 256, 512, 512), (64, 96)), ((32, 64, 64), (40, 24)), ((64, 64), (17, 30))])
def test_vae_encoder_matches_the_original_ldm_encoder_port(boc, hw):
    from transformers.models.chameleon.configuration_chameleon import ChameleonVQVAEConfig
    from transformers.models.chameleon.modeling_chameleon import ChameleonVQVAEEncoder
    from oracle.ldm_ref import make_vae
    vae = make_vae(seed=4, block_out_channels=boc).eval()
    sd = {k[len("encoder."):]: v for k, v in vae.state_dict().items() if k.startswith("encoder.")}
    cfg = ChameleonVQVAEConfig(base_channels=boc[0], channel_multiplier=[c // boc[0] for c in boc], num_res_blocks=2,
                               attn_resolutions=None, attn_type="vanilla", latent_channels=4, double_latent=True,
                               in_channels=3, resolution=256, dropout=0.0)
    enc = ChameleonVQVAEEncoder(cfg).eval()
    m = {}
    for p in ("weight", "bias"):
        m[f"conv_in.{p}"], m[f"norm_out.{p}"], m[f"conv_out.{p}"] = sd[f"conv_in.{p}"], sd[f"conv_norm_out.{p}"], sd[f"conv_out.{p}"]
    for i in range(len(boc)):
        for j in range(2):
            _map_resnet(sd, f"down_blocks.{i}.resnets.{j}", f"down.{i}.block.{j}", m)
        if i != len(boc) - 1:
            for p in ("weight", "bias"):
                m[f"down.{i}.downsample.conv.{p}"] = sd[f"down_blocks.{i}.downsamplers.0.conv.{p}"]
    _map_resnet(sd, "mid_block.resnets.0", "mid.block_1", m)
    _map_attn(sd, "mid_block.attentions.0", "mid.attn_1", m)
    _map_resnet(sd, "mid_block.resnets.1", "mid.block_2", m)
    missing, unexpected = enc.load_state_dict(m, strict=True)
    assert not missing and not unexpected
    assert set(m) == set(enc.state_dict())                       # every tensor of the oracle's encoder has a counterpart
    x = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = enc(x.clone())
        got = vae.encoder(x)
    assert got.shape == want.shape == (2, 8, hw[0] // 2 ** (len(boc) - 1), hw[1] // 2 ** (len(boc) - 1))
    assert float((got - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))


def test_vae_decoder_blocks_match_the_original_ldm_block_ports():
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from transformers.models.janus import modeling_janus as J
    from oracle.ldm_ref import make_vae
    vae = make_vae(seed=6, block_out_channels=(64, 128, 128)).eval()
    dec, sd = vae.decoder, {k[len("decoder."):]: v for k, v in vae.state_dict().items() if k.startswith("decoder.")}
    cfg = JanusVQVAEConfig(dropout=0.0)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        # mid block: ResnetBlock -> single-head attention (softmax(q k^T / sqrt(C)) v, 1x1 projections) -> ResnetBlock
        mid = J.JanusVQVAEMidBlock(cfg, 128).eval()
        m = {}
        _map_resnet(sd, "mid_block.resnets.0", "block_1", m)
        _map_attn(sd, "mid_block.attentions.0", "attn_1", m)
        _map_resnet(sd, "mid_block.resnets.1", "block_2", m)
        mid.load_state_dict(m, strict=True)
        h = torch.randn(2, 128, 9, 14, generator=g)
        assert float((dec.mid_block(h) - mid(h.clone())).abs().max()) < 2e-5
        # an up block's resnet with a channel change (1x1 shortcut) and its upsampler (nearest x2 + conv3x3)
        src = next(f"up_blocks.{i}.resnets.0" for i in range(3) if f"up_blocks.{i}.resnets.0.conv_shortcut.weight" in sd)
        cin, cout = sd[src + ".conv1.weight"].shape[1], sd[src + ".conv1.weight"].shape[0]
        rb = J.JanusVQVAEResnetBlock(cfg, cin, cout).eval()
        m = {}
        _map_resnet(sd, src, "rb", m)
        rb.load_state_dict({k[3:]: v for k, v in m.items()}, strict=True)
        i = int(src.split(".")[1])
        h = torch.randn(1, cin, 10, 12, generator=g)
        assert float((dec.up_blocks[i].resnets[0](h) - rb(h.clone())).abs().max()) < 2e-5
        up = J.JanusVQVAEConvUpsample(128).eval()
        up.load_state_dict({"conv.weight": sd["up_blocks.0.upsamplers.0.conv.weight"],
                            "conv.bias": sd["up_blocks.0.upsamplers.0.conv.bias"]}, strict=True)
        h = torch.randn(1, 128, 7, 9, generator=g)
        assert float((dec.up_blocks[0].upsamplers[0](h) - up(h)).abs().max()) < 2e-5


# ---------------------------------------------------------------------------------------------- pixel-space UNet pieces
def test_timestep_embedding_matches_the_adm_form():
    """oracle/unet_ref.timestep_embedding (diffusers ``Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0)``, SURVEY
    A.2) against an independent implementation of the ADM / improved-DDPM ``timestep_embedding`` that ships with
    torchtitan's FLUX model (cos | sin, frequencies ``exp(-ln(1e4) i / half)``)."""
    layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")
    from oracle.unet_ref import timestep_embedding
    t = torch.tensor([0.0, 1.0, 13.0, 109.0, 127.0, 999.0])
    want = layers.timestep_embedding(t, 128, time_factor=1.0)
    got = timestep_embedding(t, 128)
    assert got.shape == want.shape == (6, 128)
    assert float((got - want).abs().max()) < 1e-6


@pytest.mark.parametrize("C,hw", [(256, (6, 10)), (512, (3, 5))])
def test_attention_block_matches_torch_multihead_attention(C, hw):
    """The restated attention block (heads of 8 channels, 1/sqrt(8) scaling, GroupNorm before, residual after) against
    ``torch.nn.MultiheadAttention`` with the same projections: head split convention, scaling and softmax axis."""
    from oracle.unet_ref import Attention
    torch.manual_seed(3)
    blk = Attention(C).eval()
    mha = torch.nn.MultiheadAttention(C, C // 8, batch_first=True).eval()
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.cat([blk.to_q.weight, blk.to_k.weight, blk.to_v.weight]))
        mha.in_proj_bias.copy_(torch.cat([blk.to_q.bias, blk.to_k.bias, blk.to_v.bias]))
        mha.out_proj.weight.copy_(blk.to_out[0].weight)
        mha.out_proj.bias.copy_(blk.to_out[0].bias)
        x = torch.randn(2, C, *hw)
        y = blk.group_norm(x.view(2, C, -1)).transpose(1, 2)
        want = mha(y, y, y, need_weights=False)[0].transpose(1, 2).reshape(2, C, *hw) + x
        got = blk(x)
    assert float((got - want).abs().max()) < 2e-5

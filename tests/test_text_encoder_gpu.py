"""The text encoder of the latent family on the GPU: the three fp32 kernels of csrc/text_encoder.cu and the GELU modes of
d3_linear_small against plain PyTorch fp32 references of the same ops; ``d3roma_b200.CLIPTextModel`` against the fixture
recorded from transformers' own ``CLIPTextModel`` (tests/golden/text_encoder.npz, oracle/gen_text_encoder_golden.py) -
reduced-width models on ragged prompts, and the SD-2.1 architecture (340 M parameters) on the empty prompt, at the fp32
bound of the north star (1e-4)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import text_encoder_ref as T
from tests._golden import load

pytestmark = pytest.mark.gpu
G = load("text_encoder")


def _rnd(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).cuda()


def test_embed_tokens_and_range_check():
    from d3roma_b200 import ops
    tok, pos = _rnd((500, 96), 1), _rnd((77, 96), 2)
    ids = torch.randint(0, 500, (3, 11), generator=torch.Generator().manual_seed(3)).cuda()
    got = ops.embed_tokens(ids, tok, pos)
    assert torch.equal(got.view(3, 11, 96), tok[ids] + pos[:11][None])
    with pytest.raises(IndexError):
        ops.embed_tokens(torch.tensor([[0, 500]]).cuda(), tok, pos)
    with pytest.raises(RuntimeError):           # more tokens than positions
        ops.embed_tokens(torch.zeros((1, 78), dtype=torch.int64).cuda(), tok, pos)


@pytest.mark.parametrize("M,C,eps", [(2, 1024, 1e-5), (27, 64, 1e-5), (5, 1000, 1e-6), (1, 4, 1e-5)])
def test_layernorm_f32(M, C, eps):
    from d3roma_b200 import ops
    x = _rnd((M, C), 4, 3.0) + 1.5
    g, b = _rnd((C,), 5, 0.3) + 1.0, _rnd((C,), 6, 0.2)
    got = ops.layernorm_f32(x, g, b, eps)
    want = F.layer_norm(x.double(), (C,), g.double(), b.double(), eps)
    assert float((got.double() - want).abs().max()) < 5e-6 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("B,L,heads,hd", [(1, 2, 16, 64), (3, 9, 4, 16), (2, 77, 8, 64), (1, 128, 2, 40), (2, 33, 3, 8)])
def test_causal_attention_f32(B, L, heads, hd):
    from d3roma_b200 import ops
    D = heads * hd
    qkv = _rnd((B * L, 3 * D), 7, 1.5)
    got = ops.causal_attention_f32(qkv, B, L, heads, hd ** -0.5)
    q, k, v = (t.reshape(B, L, heads, hd).transpose(1, 2).double() for t in qkv.split(D, dim=1))
    s = q @ k.transpose(-1, -2) * hd ** -0.5 + torch.full((L, L), float("-inf"), device="cuda").triu(1)
    want = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, D)
    assert float((got.double() - want).abs().max()) < 5e-6 * max(1.0, float(want.abs().max()))
    # causality: query row 0 sees only key 0
    assert torch.allclose(got.view(B, L, D)[:, 0], qkv.view(B, L, 3 * D)[:, 0, 2 * D:], rtol=0, atol=1e-6)


@pytest.mark.parametrize("act", [2, 3])
def test_linear_small_gelu_modes(act):
    from d3roma_b200 import ops
    x, w, b = _rnd((5, 256), 8), _rnd((100, 256), 9, 0.1), _rnd((100,), 10)
    f = (lambda t: F.gelu(t)) if act == 2 else (lambda t: t * torch.sigmoid(1.702 * t))
    got_out = ops.linear_small(x, w, b, act_out=act)
    want_out = f(x.double() @ w.double().t() + b.double())
    assert float((got_out.double() - want_out).abs().max()) < 1e-5
    got_in = ops.linear_small(x, w, None, act_in=act)
    want_in = f(x.double()) @ w.double().t()
    assert float((got_in.double() - want_in).abs().max()) < 1e-5
    with pytest.raises(RuntimeError):
        ops.linear_small(x, w, b, act_out=4)


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_small_models_match_library_fixture(act):
    import d3roma_b200 as d
    cfg = dict(T.SMALL_TEXT, hidden_act=act)
    net = d.CLIPTextModel(**cfg).to("cuda").load_state_dict(T.make_state_dict(cfg, seed=3))
    for tag, ids in T.test_ids(cfg).items():
        out = net(ids.cuda())
        assert np.abs(out[0].cpu().numpy() - G[f"small_{act}.{tag}.last"]).max() < 2e-5
        assert np.abs(out.pooler_output.cpu().numpy() - G[f"small_{act}.{tag}.pooled"]).max() < 2e-5
    with pytest.raises(ValueError):             # no CPU fallback
        net(T.test_ids(cfg)["empty"])


def test_sd21_architecture_empty_prompt_matches_library_fixture(tmp_path):
    """The full SD-2.1 text encoder (23 layers, 1024 wide, 340 387 840 parameters) on the tokens of "": the [1, 2, 1024]
    context of every UNet call of the latent pipeline, against transformers' own module, fp32 bound 1e-4."""
    import d3roma_b200 as d
    net = d.CLIPTextModel(**T.SD21_TEXT).to("cuda").load_state_dict(T.make_state_dict(T.SD21_TEXT, seed=11))
    emb = net.encode_empty_prompt()
    assert emb.shape == (1, 2, 1024) and emb.dtype == torch.float32 and emb.is_cuda
    err = np.abs(emb.cpu().numpy() - G["sd21_empty.last"]).max()
    assert err < 1e-4, err
    # the pipeline asks the encoder itself when no tokenizer is given (custom_pipelines.py:260-287)
    pipe = d.GuidedLatentDiffusionPipeline.__new__(d.GuidedLatentDiffusionPipeline)
    pipe.tokenizer, pipe.text_encoder = None, net
    assert torch.equal(pipe._encode_empty_text(), emb)

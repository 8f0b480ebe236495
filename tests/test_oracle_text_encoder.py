"""The text-encoder oracle (oracle/text_encoder_ref.py) pinned to the ``transformers`` library: against the committed
fixture recorded from the library's own ``CLIPTextModel`` (oracle/gen_text_encoder_golden.py), against the library live
when it is importable, and the SD-2.1 architecture's parameter count.  Plus the host logic of
``d3roma_b200.CLIPTextModel`` over torch emulations of its ops (no GPU)."""
import numpy as np
import pytest
import torch

from oracle import text_encoder_ref as T
from tests._golden import load

G = load("text_encoder")


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_oracle_matches_library_fixture(act):
    cfg = dict(T.SMALL_TEXT, hidden_act=act)
    sd = T.make_state_dict(cfg, seed=3)
    for tag, ids in T.test_ids(cfg).items():
        last, pooled = T.text_encoder_forward(sd, cfg, ids)
        assert np.abs(last.numpy() - G[f"small_{act}.{tag}.last"]).max() < 2e-5
        assert np.abs(pooled.numpy() - G[f"small_{act}.{tag}.pooled"]).max() < 2e-5


def test_oracle_sd21_architecture():
    shapes = T.expected_shapes(T.SD21_TEXT)
    assert sum(int(np.prod(s)) for s in shapes.values()) == T.SD21_TEXT_PARAMS
    sd = T.make_state_dict(T.SD21_TEXT, seed=11)
    last, _ = T.text_encoder_forward(sd, T.SD21_TEXT, T.test_ids(T.SD21_TEXT)["empty"])
    assert last.shape == (1, 2, 1024)
    assert np.abs(last.numpy() - G["sd21_empty.last"]).max() < 1e-4


def test_oracle_matches_live_library():
    transformers = pytest.importorskip("transformers")
    from oracle.gen_text_encoder_golden import library_model
    cfg = dict(T.SMALL_TEXT, num_hidden_layers=2)
    sd = T.make_state_dict(cfg, seed=8)
    m = library_model(cfg, sd)
    ids = T.test_ids(cfg)["batch"]
    with torch.no_grad():
        want = m(input_ids=ids)
    last, pooled = T.text_encoder_forward(sd, cfg, ids)
    assert float((last - want.last_hidden_state).abs().max()) < 2e-5
    assert float((pooled - want.pooler_output).abs().max()) < 2e-5
    # rows before a prompt's end-of-text token do not depend on what follows it (causal mask)
    ids2 = ids.clone()
    ids2[2, 3:] = 17
    last2, _ = T.text_encoder_forward(sd, cfg, ids2)
    assert torch.equal(last2[2, :3], last[2, :3])


# ---------------------------------------------------------------------------------------------- host logic (emulated ops)
@pytest.fixture
def emulated(monkeypatch):
    from tests import _emulated_ops as emu
    import d3roma_b200.text_encoder as te
    monkeypatch.setattr(te, "ops", emu)
    monkeypatch.setattr(te, "_TEST_DEVICE", "cpu")
    return te


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_host_logic_matches_fixture(emulated, act):
    cfg = dict(T.SMALL_TEXT, hidden_act=act)
    sd = T.make_state_dict(cfg, seed=3)
    net = emulated.CLIPTextModel(**cfg).load_state_dict(sd)
    for tag, ids in T.test_ids(cfg).items():
        out = net(ids)
        assert np.abs(out[0].numpy() - G[f"small_{act}.{tag}.last"]).max() < 2e-5
        assert np.abs(out.pooler_output.numpy() - G[f"small_{act}.{tag}.pooled"]).max() < 2e-5
        assert out.last_hidden_state is out[0]
    # the stacked q / k / v layout round-trips to transformers' key names
    back = net.state_dict()
    assert set(back) == set(sd)
    for k, v in sd.items():
        assert torch.equal(back[k], v), k


def test_host_logic_errors_and_persistence(emulated, tmp_path):
    cfg = dict(T.SMALL_TEXT)
    sd = T.make_state_dict(cfg, seed=3)
    net = emulated.CLIPTextModel(**cfg)
    with pytest.raises(RuntimeError):
        net(torch.tensor([[998, 999]]))
    bad = dict(sd)
    bad.pop("text_model.final_layer_norm.bias")
    with pytest.raises(KeyError):
        net.load_state_dict(bad)
    net.load_state_dict(dict(sd, **{"text_model.embeddings.position_ids": torch.arange(77)[None]}))   # old buffer: ignored
    with pytest.raises(ValueError):
        net(torch.zeros((1, 78), dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        net(torch.tensor([[998, 999]]), attention_mask=torch.tensor([[1, 0]]))
    with pytest.raises(NotImplementedError):
        emulated.CLIPTextModel(**dict(cfg, hidden_act="relu"))
    net.save_pretrained(tmp_path / "text_encoder")
    again = emulated.CLIPTextModel.from_pretrained(tmp_path, subfolder="text_encoder", device="cpu")
    ids = T.test_ids(cfg)["batch"]
    assert torch.equal(again(ids)[0], net(ids)[0])
    # encode_empty_prompt: BOS, EOS of the CLIP vocabulary need the full-size embedding table
    with pytest.raises(IndexError):
        net.encode_empty_prompt()


def test_latent_pipeline_uses_package_encoder_without_tokenizer(emulated):
    """GuidedLatentDiffusionPipeline._encode_empty_text with tokenizer=None asks the package's encoder for the
    embedding of "" (custom_pipelines.py:260-287)."""
    from d3roma_b200.latent_pipeline import GuidedLatentDiffusionPipeline

    class Enc:
        device = torch.device("cpu")

        def encode_empty_prompt(self):
            return torch.full((1, 2, 8), 3.0)

    pipe = GuidedLatentDiffusionPipeline.__new__(GuidedLatentDiffusionPipeline)
    pipe.tokenizer, pipe.text_encoder = None, Enc()
    assert pipe._encode_empty_text().shape == (1, 2, 8)
    pipe.text_encoder = object()
    with pytest.raises(ValueError):
        pipe._encode_empty_text()

"""ORACLE (test infrastructure only - never imported by the product path): plain-torch restatement of
``transformers.CLIPTextModel`` - the text encoder the reference loads with
``CLIPTextModel.from_pretrained("checkpoint/stable-diffusion/text_encoder")`` (/root/reference/train.py:37, 825) and
calls once per sampling loop on the tokens of the empty prompt (core/custom_pipelines.py:260-287, 299-304).

``transformers`` is an un-vendored, unpinned dependency of the reference (README.md install notes); unlike diffusers it
IS installed in this image (5.5), so this restatement is PINNED: ``oracle/gen_text_encoder_golden.py`` loads the seeded
weights of ``make_state_dict`` into the library's own ``CLIPTextModel`` and records its outputs in
``tests/golden/text_encoder.npz``; ``tests/test_oracle_text_encoder.py`` checks this file against them (and, when
transformers is importable, against the live library).  It follows ``transformers/models/clip/modeling_clip.py``:
``CLIPTextEmbeddings`` (token + position embedding), ``CLIPEncoderLayer`` (pre-LayerNorm, residual attention, residual
MLP), ``CLIPAttention`` (``head_dim ** -0.5`` scaling, causal mask), ``CLIPMLP`` (fc1 -> act -> fc2),
``CLIPTextTransformer`` (final LayerNorm; pooled row at the end-of-text token).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD21_TEXT = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                 num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5,
                 eos_token_id=49407)
SD21_TEXT_PARAMS = 340_387_840
SMALL_TEXT = dict(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                  max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, eos_token_id=999)


def expected_shapes(cfg: dict) -> Dict[str, tuple]:
    D, I = cfg["hidden_size"], cfg["intermediate_size"]
    exp = {"text_model.embeddings.token_embedding.weight": (cfg["vocab_size"], D),
           "text_model.embeddings.position_embedding.weight": (cfg["max_position_embeddings"], D)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            exp[p + f"self_attn.{n}.weight"], exp[p + f"self_attn.{n}.bias"] = (D, D), (D,)
        exp[p + "layer_norm1.weight"], exp[p + "layer_norm1.bias"] = (D,), (D,)
        exp[p + "mlp.fc1.weight"], exp[p + "mlp.fc1.bias"] = (I, D), (I,)
        exp[p + "mlp.fc2.weight"], exp[p + "mlp.fc2.bias"] = (D, I), (D,)
        exp[p + "layer_norm2.weight"], exp[p + "layer_norm2.bias"] = (D,), (D,)
    exp["text_model.final_layer_norm.weight"], exp["text_model.final_layer_norm.bias"] = (D,), (D,)
    return exp


def make_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights under transformers' key names (one CPU generator, drawn in key order): matrices
    N(0, 0.02) (0.05 for the small test sizes so that the attention is not uniform), biases N(0, 0.02), LayerNorm scales
    1 + N(0, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    std = 0.02 if cfg["hidden_size"] >= 512 else 0.05
    sd = {}
    for k, shp in expected_shapes(cfg).items():
        if "layer_norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            sd[k] = std * torch.randn(shp, generator=g)
    return sd


def _act(x: torch.Tensor, name: str) -> torch.Tensor:
    if name == "gelu":
        return F.gelu(x)
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise NotImplementedError(name)


def text_encoder_forward(sd: Dict[str, torch.Tensor], cfg: dict, input_ids: torch.Tensor):
    """-> (last_hidden_state [B, L, D], pooler_output [B, D]); modeling_clip.py CLIPTextTransformer.forward."""
    B, L = input_ids.shape
    D, heads = cfg["hidden_size"], cfg["num_attention_heads"]
    hd = D // heads
    eps = cfg.get("layer_norm_eps", 1e-5)
    # CLIPTextEmbeddings: inputs_embeds + position_embeddings(position_ids[:, :L])
    h = sd["text_model.embeddings.token_embedding.weight"][input_ids] + \
        sd["text_model.embeddings.position_embedding.weight"][:L][None]
    causal = torch.full((L, L), float("-inf")).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        x = F.layer_norm(h, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
        q, k, v = (F.linear(x, sd[p + f"self_attn.{n}.weight"], sd[p + f"self_attn.{n}.bias"])
                   .view(B, L, heads, hd).transpose(1, 2) for n in ("q_proj", "k_proj", "v_proj"))
        a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5 + causal, dim=-1) @ v          # CLIPAttention
        a = a.transpose(1, 2).reshape(B, L, D)
        h = h + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        x = F.layer_norm(h, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
        x = _act(F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]), cfg["hidden_act"])   # CLIPMLP
        h = h + F.linear(x, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    last = F.layer_norm(h, (D,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"], eps)
    if cfg.get("eos_token_id", 2) == 2:
        pos = input_ids.argmax(dim=-1)
    else:
        pos = (input_ids == cfg["eos_token_id"]).int().argmax(dim=-1)
    return last, last[torch.arange(B), pos]


def test_ids(cfg: dict) -> Dict[str, torch.Tensor]:
    """The token batches of the fixtures: the empty prompt (BOS, EOS) and a ragged-looking batch of 9-token prompts
    (BOS, words, EOS, EOS padding as CLIP tokenizers pad)."""
    V = cfg["vocab_size"]
    bos, eos = V - 2, V - 1
    g = torch.Generator().manual_seed(5)
    words = torch.randint(0, V - 2, (3, 9), generator=g)
    words[:, 0] = bos
    words[0, 8] = eos
    words[1, 5:] = eos
    words[2, 2:] = eos
    return {"empty": torch.tensor([[bos, eos]]), "batch": words}

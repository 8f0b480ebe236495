"""``CLIPTextModel`` — the text encoder of the latent family on this package's kernels.

The reference loads ``transformers.CLIPTextModel.from_pretrained("checkpoint/stable-diffusion/text_encoder")``
(``train.py:745-747``, ``evaluate.py`` through the saved pipeline) and uses it for exactly one thing on the sampling
path: the embedding of the EMPTY prompt, ``text_encoder(tokenizer("").input_ids)[0]`` -> ``[1, 2, 1024]``
(``core/custom_pipelines.py:260-287``), repeated over the batch as the cross-attention context of every UNet call.

Architecture (transformers ``CLIPTextTransformer``): token + learned position embeddings; ``num_hidden_layers``
pre-LayerNorm transformer layers (causal self-attention with ``head_dim ** -0.5`` scaling, MLP fc1 -> act -> fc2, both
with residuals); final LayerNorm.  ``[0]`` / ``.last_hidden_state`` is the normalised sequence; ``.pooler_output`` the
row at the end-of-text token.  The SD-2.1 text encoder (OpenCLIP ViT-H text tower without its last layer) is
``hidden_size 1024, 16 heads, intermediate 4096, 23 layers, "gelu"``: 340 387 840 parameters.

The prompt is a handful of tokens and the result a per-checkpoint constant, so this module computes in fp32 on small
CUDA-core kernels (``csrc/text_encoder.cu`` + ``d3_linear_small``): no 16-bit rounding enters the UNet's context.
There is no CPU fallback: the tensors must be on a CUDA device.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops

_ACT = {"gelu": 2, "quick_gelu": 3}
BOS_TOKEN_ID, EOS_TOKEN_ID = 49406, 49407      # <|startoftext|>, <|endoftext|> of the CLIP BPE vocabulary
_TEST_DEVICE = None                            # tests/: run the host logic over emulated ops on the CPU


class CLIPTextOutput(tuple):
    """Indexable like transformers' BaseModelOutputWithPooling: ``out[0]`` is the last hidden state."""

    def __new__(cls, last_hidden_state, pooler_output):
        self = super().__new__(cls, (last_hidden_state, pooler_output))
        self.last_hidden_state, self.pooler_output = last_hidden_state, pooler_output
        return self


class CLIPTextModel:
    def __init__(self, vocab_size: int = 49408, hidden_size: int = 1024, intermediate_size: int = 4096,
                 num_hidden_layers: int = 23, num_attention_heads: int = 16, max_position_embeddings: int = 77,
                 hidden_act: str = "gelu", layer_norm_eps: float = 1e-5, eos_token_id: int = EOS_TOKEN_ID, **unused):
        if hidden_act not in _ACT:
            raise NotImplementedError(f"CLIPTextModel: hidden_act {hidden_act!r} (supported: {sorted(_ACT)})")
        if hidden_size % num_attention_heads or hidden_size % 4 or intermediate_size % 4:
            raise ValueError("CLIPTextModel: hidden_size must be divisible by the head count and by 4")
        if max_position_embeddings > 128:
            raise NotImplementedError("CLIPTextModel: at most 128 positions (d3_causal_attention_f32)")
        self.config = SimpleNamespace(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                                      num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                                      max_position_embeddings=max_position_embeddings, hidden_act=hidden_act,
                                      layer_norm_eps=layer_norm_eps, eos_token_id=eos_token_id)
        self._w: Dict[str, torch.Tensor] = {}
        self._device = torch.device("cpu")
        self.dtype = torch.float32

    # ------------------------------------------------------------------ weights
    def _expected(self) -> Dict[str, tuple]:
        c = self.config
        D, I = c.hidden_size, c.intermediate_size
        exp = {"text_model.embeddings.token_embedding.weight": (c.vocab_size, D),
               "text_model.embeddings.position_embedding.weight": (c.max_position_embeddings, D),
               "text_model.final_layer_norm.weight": (D,), "text_model.final_layer_norm.bias": (D,)}
        for i in range(c.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}."
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                exp[p + f"self_attn.{n}.weight"], exp[p + f"self_attn.{n}.bias"] = (D, D), (D,)
            for n in ("layer_norm1", "layer_norm2"):
                exp[p + n + ".weight"], exp[p + n + ".bias"] = (D,), (D,)
            exp[p + "mlp.fc1.weight"], exp[p + "mlp.fc1.bias"] = (I, D), (I,)
            exp[p + "mlp.fc2.weight"], exp[p + "mlp.fc2.bias"] = (D, I), (D,)
        return exp

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """transformers key names; ``text_model.embeddings.position_ids`` (a buffer of older versions) is ignored.  The
        q / k / v projections of a layer are stacked into one [3D, D] matrix (one Linear launch per layer)."""
        sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
        exp = self._expected()
        missing = sorted(set(exp) - set(sd))
        extra = sorted(set(sd) - set(exp))
        if missing or (strict and extra):
            raise KeyError(f"CLIPTextModel.load_state_dict: missing {missing[:4]} unexpected {extra[:4]}")
        for k, shp in exp.items():
            if tuple(sd[k].shape) != shp:
                raise ValueError(f"CLIPTextModel.load_state_dict: {k} has shape {tuple(sd[k].shape)}, expected {shp}")
        dev = self._device

        def f(k):
            return sd[k].detach().to(dev, torch.float32).contiguous()

        w = {"tok": f("text_model.embeddings.token_embedding.weight"),
             "pos": f("text_model.embeddings.position_embedding.weight"),
             "lnf.g": f("text_model.final_layer_norm.weight"), "lnf.b": f("text_model.final_layer_norm.bias")}
        for i in range(self.config.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}."
            w[f"{i}.qkv.w"] = torch.cat([f(p + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")], 0)
            w[f"{i}.qkv.b"] = torch.cat([f(p + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")], 0)
            w[f"{i}.o.w"], w[f"{i}.o.b"] = f(p + "self_attn.out_proj.weight"), f(p + "self_attn.out_proj.bias")
            w[f"{i}.ln1.g"], w[f"{i}.ln1.b"] = f(p + "layer_norm1.weight"), f(p + "layer_norm1.bias")
            w[f"{i}.ln2.g"], w[f"{i}.ln2.b"] = f(p + "layer_norm2.weight"), f(p + "layer_norm2.bias")
            w[f"{i}.fc1.w"], w[f"{i}.fc1.b"] = f(p + "mlp.fc1.weight"), f(p + "mlp.fc1.bias")
            w[f"{i}.fc2.w"], w[f"{i}.fc2.b"] = f(p + "mlp.fc2.weight"), f(p + "mlp.fc2.bias")
        self._w = w
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        w, D = self._w, self.config.hidden_size
        sd = {"text_model.embeddings.token_embedding.weight": w["tok"],
              "text_model.embeddings.position_embedding.weight": w["pos"]}
        for i in range(self.config.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}."
            for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                sd[p + f"self_attn.{n}.weight"] = w[f"{i}.qkv.w"][j * D:(j + 1) * D]
                sd[p + f"self_attn.{n}.bias"] = w[f"{i}.qkv.b"][j * D:(j + 1) * D]
            sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"] = w[f"{i}.o.w"], w[f"{i}.o.b"]
            sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"] = w[f"{i}.ln1.g"], w[f"{i}.ln1.b"]
            sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"] = w[f"{i}.ln2.g"], w[f"{i}.ln2.b"]
            sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = w[f"{i}.fc1.w"], w[f"{i}.fc1.b"]
            sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = w[f"{i}.fc2.w"], w[f"{i}.fc2.b"]
        sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"] = w["lnf.g"], w["lnf.b"]
        return {k: v.detach().cpu().clone() for k, v in sd.items()}

    def to(self, device=None, dtype=None):
        if device is not None and not isinstance(device, torch.dtype):
            self._device = torch.device(device)
            self._w = {k: v.to(self._device) for k, v in self._w.items()}
        return self

    def eval(self):
        return self

    def parameters(self):           # train.py:827-828 freezes them (`param.requires_grad = False`)
        return iter(self._w.values())

    def requires_grad_(self, flag: bool = False):
        return self

    @property
    def device(self) -> torch.device:
        return self._device

    # ------------------------------------------------------------------ persistence (transformers layout)
    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = None, device="cuda"):
        """``config.json`` + ``model.safetensors`` (or ``pytorch_model.bin``) as ``transformers`` saves them
        (train.py:745-747 ``CLIPTextModel.from_pretrained(config.pretrained_model_name_or_path, subfolder="text_encoder")``)."""
        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, "config.json")) as fh:
            cfg = json.load(fh)
        cfg = cfg.get("text_config", cfg) if "hidden_size" not in cfg else cfg
        net = cls(**{k: v for k, v in cfg.items() if not k.startswith("_")}).to(device)
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        return net.load_state_dict(sd, strict=False)      # text_projection.* of a CLIPTextModelWithProjection is unused

    def save_pretrained(self, path):
        from safetensors.torch import save_file
        os.makedirs(path, exist_ok=True)
        cfg = dict(vars(self.config), _class_name="CLIPTextModel", architectures=["CLIPTextModel"], model_type="clip_text_model")
        with open(os.path.join(path, "config.json"), "w") as fh:
            json.dump(cfg, fh, indent=2)
        save_file({k: v.contiguous() for k, v in self.state_dict().items()}, os.path.join(path, "model.safetensors"))

    # ------------------------------------------------------------------ forward
    def __call__(self, input_ids: torch.Tensor, attention_mask=None, position_ids=None, **unused) -> CLIPTextOutput:
        """``input_ids`` int [B, L] -> (last_hidden_state fp32 [B, L, D], pooler_output fp32 [B, D]).  Padding masks are
        not supported (the reference tokenises with ``padding="do_not_pad"``, custom_pipelines.py:265-271): with the
        causal mask, padded positions after the end-of-text token cannot influence the rows before them anyway."""
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise NotImplementedError("CLIPTextModel: padding attention masks are not supported")
        if position_ids is not None:
            raise NotImplementedError("CLIPTextModel: explicit position_ids are not supported")
        if not self._w:
            raise RuntimeError("CLIPTextModel: load_state_dict() first")
        c, w = self.config, self._w
        ids = torch.as_tensor(input_ids)
        if ids.dim() != 2 or ids.dtype not in (torch.int64, torch.int32):
            raise ValueError("CLIPTextModel: input_ids must be an integer [B, L] tensor")
        if not (ids.is_cuda or _TEST_DEVICE is not None):
            raise ValueError("CLIPTextModel: input_ids must be on the CUDA device of the model (no CPU fallback)")
        ids = ids.to(self._device, torch.int64).contiguous()
        B, L = ids.shape
        if L > c.max_position_embeddings:
            raise ValueError(f"CLIPTextModel: {L} tokens exceed max_position_embeddings = {c.max_position_embeddings}")
        heads, D = c.num_attention_heads, c.hidden_size
        scale, act = float(D // heads) ** -0.5, _ACT[c.hidden_act]
        h = ops.embed_tokens(ids, w["tok"], w["pos"])                                       # [B*L, D]
        for i in range(c.num_hidden_layers):
            x = ops.layernorm_f32(h, w[f"{i}.ln1.g"], w[f"{i}.ln1.b"], c.layer_norm_eps)
            qkv = ops.linear_small(x, w[f"{i}.qkv.w"], w[f"{i}.qkv.b"])                     # [B*L, 3D]
            a = ops.causal_attention_f32(qkv, B, L, heads, scale)
            h = ops.lincomb(1.0, h, 1.0, ops.linear_small(a, w[f"{i}.o.w"], w[f"{i}.o.b"]))
            x = ops.layernorm_f32(h, w[f"{i}.ln2.g"], w[f"{i}.ln2.b"], c.layer_norm_eps)
            x = ops.linear_small(x, w[f"{i}.fc1.w"], w[f"{i}.fc1.b"], act_out=act)
            h = ops.lincomb(1.0, h, 1.0, ops.linear_small(x, w[f"{i}.fc2.w"], w[f"{i}.fc2.b"]))
        last = ops.layernorm_f32(h, w["lnf.g"], w["lnf.b"], c.layer_norm_eps).view(B, L, D)
        # pooled row: the first end-of-text token (transformers >= 4.30 with eos_token_id != 2), else the largest id
        if c.eos_token_id == 2:
            pos = ids.argmax(dim=-1)
        else:
            pos = (ids == c.eos_token_id).int().argmax(dim=-1)
        return CLIPTextOutput(last, last[torch.arange(B, device=last.device), pos])

    forward = __call__

    def encode_empty_prompt(self, tokenizer=None) -> torch.Tensor:
        """custom_pipelines.py:260-287: the ``[1, 2, D]`` embedding of ``""`` - with the caller's tokenizer when given
        (``padding="do_not_pad"``), else the two ids every CLIP tokenizer produces for it (BOS, EOS)."""
        if tokenizer is not None:
            ids = tokenizer("", padding="do_not_pad", max_length=tokenizer.model_max_length, truncation=True,
                            return_tensors="pt").input_ids
        else:
            ids = torch.tensor([[BOS_TOKEN_ID, EOS_TOKEN_ID]])
        return self(ids.to(self._device))[0]

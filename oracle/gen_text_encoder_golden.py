"""ORACLE tooling: record ``tests/golden/text_encoder.npz`` from the ``transformers`` library itself (the dependency
the reference imports at train.py:37) - seeded weights of ``oracle/text_encoder_ref.make_state_dict`` loaded into the
library's ``CLIPTextModel``, run on the token batches of ``text_encoder_ref.test_ids``:

    python oracle/gen_text_encoder_golden.py        # transformers 5.5 in this image

  small_gelu / small_quick_gelu : 3-layer, 64-wide models, empty prompt + a batch of three 9-token prompts
  sd21_empty                    : the SD-2.1 text-encoder architecture (340 387 840 parameters) on the empty prompt
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import text_encoder_ref as T  # noqa: E402


def library_model(cfg: dict, sd):
    from transformers import CLIPTextConfig, CLIPTextModel
    m = CLIPTextModel(CLIPTextConfig(**cfg, bos_token_id=cfg["vocab_size"] - 2, pad_token_id=cfg["vocab_size"] - 1)).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if not k.endswith("position_ids")] and not unexpected, (missing, unexpected)
    return m


def main():
    import transformers
    out = {"transformers_version": np.array(transformers.__version__)}
    for name, act in (("small_gelu", "gelu"), ("small_quick_gelu", "quick_gelu")):
        cfg = dict(T.SMALL_TEXT, hidden_act=act)
        sd = T.make_state_dict(cfg, seed=3)
        m = library_model(cfg, sd)
        for tag, ids in T.test_ids(cfg).items():
            with torch.no_grad():
                o = m(input_ids=ids)
            out[f"{name}.{tag}.last"] = o.last_hidden_state.numpy()
            out[f"{name}.{tag}.pooled"] = o.pooler_output.numpy()
    sd = T.make_state_dict(T.SD21_TEXT, seed=11)
    m = library_model(T.SD21_TEXT, sd)
    assert sum(p.numel() for p in m.parameters()) == T.SD21_TEXT_PARAMS
    with torch.no_grad():
        o = m(input_ids=T.test_ids(T.SD21_TEXT)["empty"])
    out["sd21_empty.last"] = o.last_hidden_state.numpy()
    np.savez_compressed(ROOT / "tests" / "golden" / "text_encoder.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

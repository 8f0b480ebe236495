"""Gate the UNet restatement (oracle/unet_ref.py) by the published parameter counts and layer census
 and pin oracle/pipeline_ref.py against the reference's own pipeline class
(tests/golden/pipeline.npz)."""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import pipeline_ref
from oracle.unet_ref import make_unet

G = Path(__file__).parent / "golden"


@pytest.mark.parametrize("cin,cout,n", [(3, 3, 113673219), (8, 1, 113676673), (5, 1, 113673217), (12, 3, 113683587)])
def test_parameter_counts(cin, cout, n):
    m = make_unet(cin, cout)
    assert sum(p.numel() for p in m.parameters()) == n


def test_layer_census_and_keys():
    m = make_unet(8, 1)
    convs = [x for x in m.modules() if isinstance(x, nn.Conv2d)]
    assert len(convs) == 96
    assert sum(1 for c in convs if c.kernel_size == (3, 3) and c.stride == (1, 1)) == 71
    assert sum(1 for c in convs if c.stride == (2, 2)) == 5
    assert sum(1 for c in convs if c.kernel_size == (1, 1)) == 20
    assert sum(1 for x in m.modules() if isinstance(x, nn.GroupNorm)) == 71
    assert sum(1 for x in m.modules() if isinstance(x, nn.Linear)) == 58
    sd = m.state_dict()
    for k in ("conv_in.weight", "time_embedding.linear_2.bias", "down_blocks.4.attentions.1.to_out.0.weight",
              "down_blocks.2.resnets.0.conv_shortcut.weight", "down_blocks.0.downsamplers.0.conv.weight",
              "mid_block.attentions.0.group_norm.weight", "up_blocks.1.attentions.2.to_q.bias",
              "up_blocks.4.upsamplers.0.conv.bias", "up_blocks.5.resnets.2.time_emb_proj.weight",
              "conv_norm_out.weight", "conv_out.bias"):
        assert k in sd, k
    assert "down_blocks.5.downsamplers.0.conv.weight" not in sd and "up_blocks.5.upsamplers.0.conv.weight" not in sd


def test_skip_stack_needs_padding():
    """the 6-level UNet cannot run at unpadded 360x640 (45 -> 23 -> 12, 12*2 != 23)."""
    m = make_unet(8, 1)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, 360 // 4, 640 // 4), torch.tensor(1))
    with torch.no_grad():
        assert m(torch.zeros(1, 8, 96, 160), torch.tensor(1)).sample.shape == (1, 1, 96, 160)


def test_pipeline_restatement_matches_reference_pipeline():
    z = np.load(G / "pipeline.npz")
    torch.set_num_threads(8)
    unet = make_unet(8, 1, seed=0, sample_size=(32, 64))
    conds = {k: torch.from_numpy(z[k]) for k in ("left", "right", "raw")}
    x_T = torch.randn((1, 1, 32, 64), generator=torch.Generator().manual_seed(42))
    rec = []
    img = pipeline_ref.run_pipeline(unet, x_T, conds, "left+right+raw", sampler="ddim", num_inference_steps=10,
                                    record=rec)
    np.testing.assert_allclose(torch.stack(rec).numpy(), z["ddim10/model_outputs"], atol=1e-5)
    np.testing.assert_allclose(img.numpy(), z["ddim10/images"], atol=1e-5)
    # DDPM: variance noise comes from the default generator, one draw per step with t > 0
    torch.manual_seed(99)
    img = pipeline_ref.run_pipeline(unet, x_T, conds, "left+right+raw", sampler="ddpm", num_inference_steps=8,
                                    ddpm_noise=lambda i, shape: torch.randn(shape))
    np.testing.assert_allclose(img.numpy(), z["ddpm8/images"], atol=1e-5)


def test_oracle_forward_regression_fixture():
    """The restatement's own outputs are frozen in tests/golden/unet_oracle.npz (oracle/gen_golden.py:gen_unet_oracle):
    not a pin to diffusers (absent), but a guard against the oracle drifting while kernels are compared against it."""
    from pathlib import Path
    import numpy as np
    from oracle.unet_ref import make_unet
    z = np.load(Path(__file__).resolve().parent / "golden" / "unet_oracle.npz")
    torch.set_num_threads(8)
    for tag in ("c8", "c5", "c12"):
        cin, dc, h, w, t = (int(v) for v in z[f"{tag}/meta"])
        net = make_unet(cin, dc, seed=0, sample_size=(h, w))
        assert sum(p.numel() for p in net.parameters()) == int(z[f"{tag}/n_params"])
        with torch.no_grad():
            y = net(torch.from_numpy(z[f"{tag}/x"]), torch.tensor(t)).sample
        np.testing.assert_allclose(y.numpy(), z[f"{tag}/y"], rtol=0, atol=5e-5)

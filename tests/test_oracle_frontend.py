"""The oracle's restatement of the frame preprocessing (inference.py:108-213) against the cv2-generated golden vectors."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden" / "frontend.npz"
CASES = ["down", "up", "same", "odd", "gray"]


@pytest.mark.parametrize("name", CASES)
def test_frontend_oracle_matches_cv2_golden(name):
    cv2 = pytest.importorskip("cv2")
    from oracle import frontend_ref as F
    g = np.load(GOLD)
    if tuple(int(v) for v in cv2.__version__.split(".")[:2]) != tuple(int(v) for v in g["cv2_version"][:2]):
        pytest.skip("different OpenCV minor version than the one that generated the fixture")
    h, w = (int(v) for v in g[f"{name}/size"])
    img = g[f"{name}/img"]
    np.testing.assert_array_equal(F.preprocess_image(img, (h, w)).numpy(), g[f"{name}/normalized"])
    raw, mask, norm = F.preprocess_depth(g[f"{name}/depth"], (h, w), 24.547)
    np.testing.assert_array_equal(raw.numpy(), g[f"{name}/raw_disp"])
    np.testing.assert_array_equal(norm.numpy(), g[f"{name}/norm_disp"])
    np.testing.assert_array_equal(mask.numpy(), (g[f"{name}/raw_disp"] > 0).astype(np.float32))


def test_pad_replicate_matches_input_padder_geometry():
    import torch
    from oracle import frontend_ref as F
    x = torch.arange(36 * 64, dtype=torch.float32).view(1, 36, 64)
    y = F.pad_replicate(x, 32)
    assert y.shape == (1, 64, 64)
    assert torch.equal(y[:, 14:50], x) and torch.equal(y[:, 0], x[:, 0]) and torch.equal(y[:, -1], x[:, -1])

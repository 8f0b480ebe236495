"""Pin oracle/guided_pipeline_ref.py against tests/golden/guided.npz: the guided DDIM loop recorded from the
reference's own scheduler class (hook at core/scheduler_ddim.py:461-472), SSIM module, Normalizer and camera maps."""
from pathlib import Path

import numpy as np
import torch

from oracle import guided_pipeline_ref as GP

G = Path(__file__).resolve().parent / "golden"


def _load():
    z = np.load(G / "guided.npz")
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    return z, t


def test_x0_gradient_matches_reference_chain():
    z, t = _load()
    mind, maxd, shift = (float(v) for v in z["cam"])
    for i in (0, 4, 9):
        loss, g = GP.x0_loss_and_grad(t(f"guided/x0{i}"), t("left"), t("right"), 128.0, mind, maxd, shift)
        ref = z[f"guided/g{i}"]
        np.testing.assert_allclose(g.numpy(), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
        assert abs(float(loss) - float(z[f"guided/loss{i}"])) <= 1e-4 * float(z[f"guided/loss{i}"])


def test_guided_loop_teacher_forced_matches_reference():
    z, t = _load()
    mind, maxd, shift = (float(v) for v in z["cam"])
    rec = []
    conds = {"left": t("left"), "right": t("right"), "raw": t("raw")}
    out = GP.run_guided_pipeline(None, t("x_T"), conds, "left+right+raw", t("left"), t("right"),
                                 weight=float(z["weight"]), bound=128.0, min_disp=mind, max_disp=maxd, shift=shift,
                                 num_inference_steps=10, record=rec, teacher_outputs=[t(f"guided/m{i}") for i in range(10)])
    for i, r in enumerate(rec):
        # teacher forcing feeds the recorded model outputs; the sample trajectory then has to follow the reference's
        np.testing.assert_allclose(r["sample"].numpy(), z[f"guided/x{i}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(r["x0_perturbed"].numpy(), z[f"guided/x0p{i}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(r["prev"].numpy(), z[f"guided/prev{i}"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.numpy(), z["guided/images"], rtol=0, atol=2e-5)


def test_guidance_lowers_the_photometric_loss_of_the_result():
    z, _ = _load()
    assert float(z["guided/final_loss"]) < float(z["plain/final_loss"])

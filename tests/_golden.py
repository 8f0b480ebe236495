"""Golden-vector access for the CPU tests.

``tests/golden/*.npz`` were recorded by ``oracle/gen_golden.py`` from the reference's own classes on ONE host.  The
restatements under ``oracle/`` reproduce the reference bit for bit *on the host that runs both*; across hosts torch's
CPU kernels for ``pow`` / ``sqrt`` / ``exp`` take different SIMD paths (``torch.backends.cpu.get_cpu_capability()``:
AVX2 on the recording host, AVX512 on others) that differ by an ulp per call and compound to < 1e-5 over a 10-step
trajectory.  So:

  * ``source = "live"``      - the reference is imported from /root/reference (when present) and re-run on THIS host by
                               the same generator functions: comparisons are bit-exact (``exact=True``).
  * ``source = "committed"`` - the committed fixtures: bit-exact for everything that does not go through those kernels
                               (tables, timesteps, gamma = 1 normalisation, padding), a few ulp otherwise.
"""
from __future__ import annotations

import functools
from pathlib import Path

import numpy as np
import pytest

G = Path(__file__).parent / "golden"
REF = Path("/root/reference")
HAVE_REF = (REF / "core" / "custom_pipelines.py").exists()

SOURCES = ["committed", pytest.param("live", marks=pytest.mark.skipif(not HAVE_REF, reason="reference tree absent"))]


@functools.lru_cache(maxsize=None)
def load(name: str, source: str = "committed"):
    """The arrays of ``tests/golden/<name>.npz`` - or, for ``source="live"``, the same arrays re-made here by
    ``oracle.gen_golden.gen_<name>`` from the imported reference."""
    if source == "committed":
        return np.load(G / f"{name}.npz")
    import torch
    from oracle import gen_golden
    store: dict = {}
    with torch.no_grad():
        getattr(gen_golden, f"gen_{name}")(store)
    return {k: np.asarray(v) for k, v in store.items()}


def same(got, want, source: str = "committed", rtol: float = 1e-5, atol: float = 1e-5) -> bool:
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        return False
    if np.array_equal(got, want):
        return True
    return source == "committed" and bool(np.allclose(got, want, rtol=rtol, atol=atol))

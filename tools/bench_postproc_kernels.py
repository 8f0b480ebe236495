"""CUDA-event timing of the bandwidth kernels added for the evaluation side (one B200): the gather-fused denormalise
(single rank: the local store only), the two ensemble_depths kernels, against their algorithmic bytes and the measured
copy bandwidth in MEASURED_PEAKS.json."""
import ctypes
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import _lib, ops  # noqa: E402
from d3roma_b200.distributed import PeerGather  # noqa: E402

peaks = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
peak = 6538.0
if peaks.exists():
    pk = json.loads(peaks.read_text())
    peak = float(pk.get("hbm_gbs", peak))


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


def report(name, sec, nbytes):
    print(f"{name:58s} {sec * 1e6:8.1f} us  {nbytes / 1e6:8.1f} MB algorithmic  {nbytes / sec / 1e9:7.0f} GB/s = "
          f"{nbytes / sec / 1e9 / peak:.2f} of {peak:.0f}", flush=True)


B, H, W = 64, 360, 640           # 64 maps: 59 MB in + 59 MB out, beyond what one launch keeps in flight
norm = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
y = torch.rand(B, 1, H, W, device="cuda") * 2 - 1
pg = PeerGather({"disp": ((B, 1, H, W), torch.float32)}, rank=0, world=1)


def fused():
    pg.denormalize_into(y, norm)
    pg.finish()


report("denormalize_gather (1 rank) + signal + wait", timed(fused), 2 * y.numel() * 4)
report("Normalizer.denormalize (d3_denormalize_disp)", timed(lambda: norm.denormalize(y)), 2 * y.numel() * 4)
pg.close()

f = ctypes.POINTER(ctypes.c_float)
for n in (4, 8, 16):
    x = torch.rand(n, 4 * H, W, device="cuda")          # 4 samples' worth of pixels per map: > L2 for n >= 8
    P = x[0].numel()
    s, t = np.ones(n, np.float32), np.zeros(n, np.float32)
    ws = torch.empty((_lib.lib().d3_ensemble_workspace_bytes(),), dtype=torch.uint8, device="cuda")
    out3 = torch.empty(3, dtype=torch.float64, device="cuda")
    al, un = torch.empty(4 * H, W, device="cuda"), torch.empty(4 * H, W, device="cuda")
    for median in (1, 0):
        tag = "median" if median else "mean"
        report(f"ensemble_objective n={n} {tag}", timed(lambda: _lib.check(_lib.lib().d3_ensemble_objective(
            x.data_ptr(), n, P, s.ctypes.data_as(f), t.ctypes.data_as(f), median, ws.data_ptr(), out3.data_ptr(),
            ops._stream()))), n * P * 4)
        report(f"ensemble_reduce n={n} {tag} (+ rescale pass)", timed(lambda: _lib.check(_lib.lib().d3_ensemble_reduce(
            x.data_ptr(), n, P, s.ctypes.data_as(f), t.ctypes.data_as(f), median, ws.data_ptr(), al.data_ptr(),
            un.data_ptr(), ops._stream()))), (n + 2 + 4) * P * 4)

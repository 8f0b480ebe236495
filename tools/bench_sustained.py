"""Developer benchmark: one conv / GN shape launched back to back for several seconds (power-capped steady state),
with nvidia-smi clocks and power sampled during the run.  Usage: bench_sustained.py [seconds] [shape-index ...]"""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from d3roma_b200 import ops  # noqa: E402

SHAPES = [
    (16, 384, 640, 128, 128, 3, 1, 0, 0),
    (16, 384, 640, 256, 128, 3, 1, 128, 128),
    (16, 96, 160, 256, 256, 3, 1, 0, 0),
]


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                                    "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True,
                                   timeout=5).stdout.strip().split(",")
                self.rows.append((float(o[0]), float(o[1]), o[2].strip()))
            except Exception:
                pass
            time.sleep(0.2)


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) / 10
    n = max(20, int(seconds * 1e3 / burst))
    s = Sampler(); s.start()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    s.stop = True
    ms = e0.elapsed_time(e1) / n
    rows = s.rows[len(s.rows) // 3:] or [(0, 0, "")]
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    pw = sorted(r[1] for r in rows)[len(rows) // 2]
    return burst, ms, clk, pw


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    which = [int(a) for a in sys.argv[2:]] or list(range(len(SHAPES)))
    for i in which:
        (B, H, W, Cin, Cout, k, s, C1, C2) = SHAPES[i]
        x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
        Kt = k * k * Cin + C1 + C2
        wp = (torch.randn(Cout, Kt, device="cuda") / Kt ** 0.5).to(torch.bfloat16)
        bias = torch.zeros(Cout, device="cuda")
        sc1 = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16) if C1 else None
        sc2 = torch.randn(B, H, W, C2, device="cuda").to(torch.bfloat16) if C2 else None
        out = torch.empty(B, H // s, W // s, Cout, device="cuda", dtype=torch.bfloat16)
        stats = torch.zeros(B, Cout // 4, 2, dtype=torch.float64, device="cuda")
        fn = lambda: ops.conv2d_igemm(x, wp, bias, ksize=k, stride=s, sc1=sc1, sc2=sc2, out=out, stats=stats)  # noqa
        burst, ms, clk, pw = sustained(fn, seconds)
        flops = 2.0 * B * (H // s) * (W // s) * Cout * Kt
        print(f"conv {H}x{W} {Cin}->{Cout} sc={C1}+{C2}: burst {flops / burst / 1e9:.0f} TF, sustained "
              f"{flops / ms / 1e9:.0f} TF ({ms:.3f} ms) @ {clk:.0f} MHz {pw:.0f} W", flush=True)
    if not sys.argv[2:]:
        a = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
        b = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
        c = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
        burst, ms, clk, pw = sustained(lambda: torch.matmul(a, b, out=c), seconds)
        fl = 2 * 8192.0 ** 3
        print(f"cuBLAS 8192^3: burst {fl / burst / 1e9:.0f} TF, sustained {fl / ms / 1e9:.0f} TF @ {clk:.0f} MHz {pw:.0f} W")
        x = torch.randn(16, 384, 640, 128, device="cuda").to(torch.bfloat16)
        y = torch.empty_like(x)
        st = torch.zeros(16, 32, 2, dtype=torch.float64, device="cuda")
        st[:, :, 1] = 4.0 * 384 * 640
        g = torch.ones(128, device="cuda"); bt = torch.zeros(128, device="cuda")
        burst, ms, clk, pw = sustained(lambda: ops.groupnorm_silu(x, st, g, bt, groups=32, eps=1e-5, out=y), seconds)
        by = 2.0 * x.numel() * 2
        print(f"GN 128ch full-res: burst {by / burst / 1e6:.0f} GB/s, sustained {by / ms / 1e6:.0f} GB/s @ {clk:.0f} MHz {pw:.0f} W")


if __name__ == "__main__":
    main()

"""Developer tool: per-kernel time shares of the UNet forward in the power-capped steady state.  Runs forwards back to
back for a few seconds, then records 3 forwards with torch.profiler (CUPTI activity trace: kernel durations without
serialisation or replay) and prints the per-kernel totals.  usage: profile_sustained.py [B] [dtype] [seconds]"""
import sys
import time
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200.weights import random_init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
net = d.UNet2DModel(sample_size=(384, 640), in_channels=8, out_channels=1, compute_dtype=dtype).to("cuda")
net.load_state_dict(random_init_state_dict(8, 1))
x = torch.randn(B, 1, 384, 640, device="cuda")
c = torch.randn(B, 7, 384, 640, device="cuda")
t0 = time.time()
n = 0
while time.time() - t0 < seconds:
    net.forward_split(x, c, 73)
    n += 1
    if n % 4 == 0:
        torch.cuda.synchronize()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net.forward_split(x, c, 73)
e1.record()
torch.cuda.synchronize()
print(f"sustained forward: {e0.elapsed_time(e1) / 5:.2f} ms (eager launches)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        net.forward_split(x, c, 73)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        tot[name][0] += ev.device_time_total / 3e3 if hasattr(ev, "device_time_total") else ev.cuda_time_total / 3e3
        tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"kernel time per forward: {total:.2f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {v[0]:7.3f} ms {100 * v[0] / total:5.1f}%  x{v[1] // 3:4d}  {k}")
if len(sys.argv) > 4:  # per-launch list of the last recorded forward (kernel order)
    evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 3
    for ev in evs[2 * per:]:
        nm = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:48]
        print(f"    {ev.device_time_total:9.1f} us  {nm}")

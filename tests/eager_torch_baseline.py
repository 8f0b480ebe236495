"""Not a test: the "same box" GPU baseline SURVEY.md §8(d) asks for - the oracle's plain torch.nn restatement of the
denoiser run by eager PyTorch (cuDNN / cuBLAS / ATen) on the same B200, bf16 + channels_last and fp32 (TF32 off / on),
B = 16 at 384x640, sustained for a few seconds.  Lives under tests/ because it executes oracle/ code.
usage: python tests/eager_torch_baseline.py [B] [seconds]  ->  one line per mode: ms per UNet forward."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.unet_ref import make_unet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
torch.backends.cudnn.benchmark = True


def run(dtype, channels_last, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    net = make_unet(8, 1).to("cuda", dtype)
    x = torch.randn(B, 8, 384, 640, device="cuda", dtype=dtype)
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    t = torch.tensor(73, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            net(x, t)
        torch.cuda.synchronize()
        t0, n = time.time(), 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < seconds:
            net(x, t)
            n += 1
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"eager torch {str(dtype).split('.')[-1]}{' channels_last' if channels_last else ''}"
          f"{' tf32' if tf32 else ''}: {ms:.2f} ms per UNet forward (B={B}), {B / ms * 100:.2f} maps/s at 10 steps, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net, x
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


run(torch.bfloat16, True, False)
run(torch.bfloat16, False, False)
run(torch.float32, False, True)
run(torch.float32, False, False)

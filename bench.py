#!/usr/bin/env python
"""Benchmark of the D3RoMa denoising hot path (BASELINE.json metric: disparity maps/sec at 360x640, 10 DDIM steps).

    python bench.py --gpus N --steps K --warmup W          # this implementation (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W  # the reference path on the host CPU cores (oracle port)

One "step" = one pass of the hot path over one batch: the left+right+raw variant, 16 stereo pairs per GPU at
360x640 (computed at the padded 384x640 like the reference's eval path), 10 DDIM steps (BASELINE configs[1]).
For N > 1 the driver launches this under torchrun; independent stereo pairs are sharded by rank (weak scaling,
16 pairs per GPU) and the final disparity maps are all-gathered over NCCL inside the timed step.

JSON keys beyond the base contract:
  roofline      dominant kernel (tcgen05 implicit-GEMM conv, shape 3x3 128->128 @384x640, 31 % of the FLOPs) timed
                alone with CUDA events: achieved TFLOP/s vs the measured cuBLAS bf16 burst peak.
  roofline_hbm  the bandwidth kernel (GroupNorm+SiLU, 128 ch @384x640) vs the measured copy bandwidth.
  unet          whole-UNet algorithmic TFLOP/s inside the timed region vs the sustained peak.
  cpu_baseline  the oracle port of the reference pipeline timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H0, W0 = 360, 640           # image size the metric is quoted on
HP, WP = 384, 640           # padded size the UNet runs at (InputPadder(divis_by=32))
B_PER_GPU = 16
DDIM_STEPS = 10
UNET_GFLOP = 1872.53        # per sample per UNet forward at 384x640, Cin=8, dc=1
DDIM_CFG = dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample",
                clip_sample=True, set_alpha_to_one=False, steps_offset=1)
METRIC = "disparity maps/sec at 360x640, 10 DDIM steps"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- synthetic inputs
def synth_inputs(n: int, first_index: int = 0, h: int = H0, w: int = W0):
    """seeded per global sample index; images uint8-like in [-1,1], raw depth uniform [0.3,2] m with
    20 % holes -> disparity -> Normalizer(average, bound 128) -> [-1,1].  Host tensors (CPU)."""
    lefts, rights, depths = [], [], []
    for i in range(n):
        g = torch.Generator().manual_seed(1234 + first_index + i)
        lefts.append((torch.randint(0, 256, (3, h, w), generator=g).float() / 255.0 - 0.5) * 2)
        rights.append((torch.randint(0, 256, (3, h, w), generator=g).float() / 255.0 - 0.5) * 2)
        d = torch.rand((1, h, w), generator=g) * 1.7 + 0.3
        d[torch.rand((1, h, w), generator=g) < 0.2] = 0.0
        depths.append(d)
    return torch.stack(lefts), torch.stack(rights), torch.stack(depths)


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_steps(n_steps: int, n_warm: int, threads: int):
    """Time the oracle port of the reference pipeline (core/custom_pipelines.py loop + restated UNet) on the host:
    each timed unit = 1 stereo pair, ONE DDIM step (UNet forward + scheduler update) at 384x640, fp32."""
    from oracle import pipeline_ref, prepost_ref as P
    from oracle.unet_ref import make_unet
    torch.set_num_threads(threads)
    unet = make_unet(8, 1, seed=0, sample_size=(HP, WP))
    left, right, depth = synth_inputs(1)
    raw = P.normalize_average(P.depth_to_disp(depth[0], 24.5471958), 128.0)[None]
    conds = {"left": P.pad_replicate(left), "right": P.pad_replicate(right), "raw": P.pad_replicate(raw)}
    x_T = torch.randn((1, 1, HP, WP), generator=torch.Generator().manual_seed(42))
    times = []
    for i in range(n_warm + n_steps):
        t0 = time.perf_counter()
        pipeline_ref.run_pipeline(unet, x_T, conds, "left+right+raw", sampler="ddim", num_inference_steps=1)
        dt = time.perf_counter() - t0
        if i >= n_warm:
            times.append(dt)
    sec_per_unet_step = sum(times) / len(times)
    return sec_per_unet_step, 1.0 / (DDIM_STEPS * sec_per_unet_step)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sec, maps = cpu_reference_steps(max(1, args.steps), max(1, min(args.warmup, 2)), threads)
    sample = ("1 synthetic stereo pair 384x640 (padded 360x640), ONE DDIM step (UNet forward + scheduler update) per "
              "timed step, fp32, scaled x10 steps to maps/s")
    line = {
        "impl": "reference", "metric": METRIC, "value": maps, "unit": "maps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * DDIM_STEPS,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": maps, "unit": "maps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": maps, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = jie0530/d3roma core/custom_pipelines.py loop restated in oracle/ (diffusers.UNet2DModel "
                "is an un-vendored third-party dependency, so the CPU path is the oracle port, torch CPU fp32)",
    }
    emit(line)


def workload_config(n_gpus: int):
    return {"workload": "left+right+raw variant, batch 16 per GPU at 360x640 (UNet at padded 384x640), DDIM 10 steps, "
                        "random-init UNet2DModel(8->1, 113.68 M params), BASELINE configs[1]",
            "pairs_per_gpu": B_PER_GPU, "global_pairs": B_PER_GPU * n_gpus, "ddim_steps": DDIM_STEPS,
            "parallelism": f"batch-shard x{n_gpus} (replicated weights, output all-gather)",
            "l2": "per-step working set (>1 GB activations per layer) exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------------------------- GPU arm
def build_pipeline(dtype: str, device):
    import d3roma_b200 as d
    from d3roma_b200.weights import random_init_state_dict
    net = d.UNet2DModel(sample_size=(HP, WP), in_channels=8, out_channels=1, compute_dtype=dtype).to(device)
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    pipe = d.GuidedDiffusionPipeline(net, d.MyDDIMScheduler.from_config(DDIM_CFG),
                                     d.FlowGuidance(flow_guidance_weight=0.0))
    return d, pipe


def time_dominant_kernels(dtype, peaks):
    """Roofline legs: the dominant conv shape and the GroupNorm+SiLU pass, each timed alone with CUDA events."""
    from d3roma_b200 import ops
    adt = torch.bfloat16 if dtype == "bf16" else torch.float16
    B = B_PER_GPU
    x = torch.randn(B, HP, WP, 128, device="cuda").to(adt)
    w = (torch.randn(128, 1152, device="cuda") / 34.0).to(adt)
    bias = torch.zeros(128, device="cuda")
    out = torch.empty_like(x)
    stats = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
    gamma, beta = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")

    def timeit(fn, n=10, warm=3):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3

    t_conv = timeit(lambda: ops.conv2d_igemm(x, w, bias, ksize=3, out=out, stats=stats))
    flops = 2.0 * B * HP * WP * 128 * 1152
    t_gn = timeit(lambda: ops.groupnorm_silu(x, stats, gamma, beta, out=out))
    gn_bytes = 2.0 * x.numel() * 2
    roof = {"bound": "tensor", "kernel": "conv_igemm_t_kernel (pixels-as-N, 16x16 halo blocks) 3x3 128->128 @16x384x640",
            "achieved": flops / t_conv / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
            "frac": flops / t_conv / 1e12 / peaks["tf_burst"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r01b_ncu_full_summary.csv)
            "traffic": 1.014290e9 + 0.964388e9, "algorithmic_bytes": 2.0 * x.numel() * 2,
            "ncu_tensor_pipe_active_pct": 69.5,
            "flops_per_launch": flops, "ms_per_launch": t_conv * 1e3, "peak_source": peaks["src"] + " burst"}
    roof_hbm = {"bound": "hbm", "kernel": "gn_apply_kernel 128ch @16x384x640 (GroupNorm+SiLU)",
                "achieved": gn_bytes / t_gn / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": gn_bytes / t_gn / 1e9 / peaks["hbm"], "traffic": 1.006704e9 + 0.950754e9,
                "bytes_per_launch": gn_bytes,
                "ms_per_launch": t_gn * 1e3, "peak_source": peaks["src"]}
    return roof, roof_hbm


def run_gpu_arm(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the d3roma_b200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    from d3roma_b200 import _lib
    d, pipe = build_pipeline(args.dtype, dev)
    from d3roma_b200.evaluate import denoise_batch
    normalizer = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    cam = d.SIM_CAMERA

    # ---- synthetic inputs: host (pinned) copies for the e2e leg, device-resident padded copies for `value`
    left_h, right_h, depth_h = synth_inputs(B_PER_GPU, first_index=rank * B_PER_GPU)
    left_h, right_h, depth_h = left_h.pin_memory(), right_h.pin_memory(), depth_h.pin_memory()
    padder = d.InputPadder(left_h.shape, divis_by=32)

    def prep_device():
        dd = depth_h.to(dev)
        raw = normalizer.normalize(d.prepost.depth_to_disp(dd, cam.fxb).view(B_PER_GPU, H0, W0))[0]
        l, r, rw = padder.pad(left_h.to(dev), right_h.to(dev), raw.view(B_PER_GPU, 1, H0, W0))
        return torch.cat([l, r, rw], 1).contiguous()

    cond = prep_device()
    gens = [torch.Generator().manual_seed(10_000 + rank * B_PER_GPU + i) for i in range(B_PER_GPU)]
    x_T = torch.cat([torch.randn((1, 1, HP, WP), generator=g) for g in gens]).to(dev)
    gather_buf = torch.empty((world * B_PER_GPU, 1, H0, W0), device=dev) if world > 1 else None

    def one_step():
        img = pipe.sample_from(x_T, cond, DDIM_STEPS)
        disp = normalizer.denormalize(padder.unpad(img))
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, disp)
        return disp

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        one_step()
    sync()
    lib = _lib.lib()
    lib.d3_reset_launch_count()
    g0 = pipe.unet.graph_kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.d3_launch_count()) + (pipe.unet.graph_kernel_launches - g0)
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    value = world * B_PER_GPU * args.steps / sec

    # ---- e2e: the public call chain with HOST buffers (pad -> pipeline.__call__ -> unpad -> denormalise -> D2H)
    def e2e_step():
        dd = depth_h.to(dev, non_blocking=True)
        raw = normalizer.normalize(d.prepost.depth_to_disp(dd, cam.fxb).view(B_PER_GPU, H0, W0))[0]
        return denoise_batch(pipe, normalizer, left=left_h, right=right_h, sim_disp=raw.view(B_PER_GPU, 1, H0, W0),
                             cond_channels="left+right+raw", num_inference_steps=DDIM_STEPS,
                             num_intermediate_images=1, generator=gens, to_host=True)

    for _ in range(2):
        e2e_step()
    sync()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n_e2e):
        out_h = e2e_step()
    e1.record()
    sync()
    wall = time.perf_counter() - t0
    te = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B_PER_GPU * n_e2e / float(te.item())
    h2d = left_h.numel() * 4 + right_h.numel() * 4 + depth_h.numel() * 4 + x_T.numel() * 4
    d2h = out_h.numel() * 4

    if rank == 0:
        roof, roof_hbm = time_dominant_kernels(args.dtype, peaks)
        unet_tf = world * B_PER_GPU * DDIM_STEPS * UNET_GFLOP * 1e9 * args.steps / sec / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(world),
            "ms_per_unet_step": sec / args.steps / DDIM_STEPS * 1e3,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "maps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e, "api": "d3roma_b200.evaluate.denoise_batch -> GuidedDiffusionPipeline.__call__"},
            "gpu_launches": launches,
            "roofline": roof, "roofline_hbm": roof_hbm,
            "unet": {"algorithmic_tflops": unet_tf / world, "peak": peaks["tf_sust"],
                     "frac_of_sustained_peak": unet_tf / world / peaks["tf_sust"], "gflop_per_sample_step": UNET_GFLOP},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sec_cpu, maps_cpu = cpu_reference_steps(3, 1, threads)
            line["cpu_baseline"] = {"value": maps_cpu, "unit": "maps/s", "cores": threads, "kind": "port",
                                    "sample": "oracle port of the reference loop: 1 pair 384x640, 3 timed single "
                                              "DDIM steps (UNet forward + scheduler), fp32, scaled x10 to maps/s"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; see main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main():
    # Libraries write banners to file descriptor 1 (e.g. "NCCL version ..." at communicator creation).  Keep a private
    # handle on the real stdout for the JSON line and point fd 1 at stderr for everything else.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="d3roma_b200", choices=["d3roma_b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

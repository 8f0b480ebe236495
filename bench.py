#!/usr/bin/env python
"""Benchmark of the D3RoMa denoising hot path (BASELINE.json metric: disparity maps/sec at 360x640, 10 DDIM steps).

    python bench.py --gpus N --steps K --warmup W [--config c2|c1|c3|c4|c5|ldm|train] [--dtype fp16|bf16]
    python bench.py --impl reference --steps K --warmup W     # the reference path on the host CPU cores (oracle port)

One "step" = one pass of the hot path over one batch.  The default workload is BASELINE configs[1] ("c2"): the
left+right+raw variant, 16 stereo pairs per GPU at 360x640 (computed at the padded 384x640 like the reference's eval
path), 10 DDIM steps, in the fp16-operand mode (the mode that meets the north-star per-step tolerance; `--dtype bf16`
times the secondary bf16 mode).  For N > 1 the driver launches this under torchrun; independent stereo pairs are
sharded by rank (weak scaling) and the final disparity maps are all-gathered inside the timed step: over NCCL by
default, or (`--gather peer`) as stores into NVLink peer memory issued by the denormalise kernel itself
(d3roma_b200.distributed.PeerGather; `config.gather` names what ran).

Other BASELINE configs (parity-test cases made measurable; `config.workload` names what ran):
  c1  1 pair, DDIM 10 (configs[0] on the GPU: the B = 1 latency case)
  c3  rgb+raw variant with stereo-geometry guidance, 8 pairs per GPU (configs[2]: 64 pairs on 8 GPUs)
  c4  540x960 (UNet at 544x960), DDPM 100 steps, fp32 reference-precision path, 1 pair per GPU (configs[3])
  c5  32 pairs per GPU, DDIM 10, disparity maps AND the [n,11] metric rows all-gathered (configs[4]: 256 pairs on 8 GPUs)
  ldm the latent family: rgb+raw, SD-2.1 UNet2DConditionModel + SD VAE, DDPM 10 of 1000 steps, 16 pairs per GPU

JSON keys beyond the base contract:
  roofline            dominant kernel (tcgen05 implicit-GEMM conv, 3x3 128->128 @384x640, 31 % of the FLOPs) timed alone
                      with CUDA events in this run: achieved TFLOP/s vs the measured cuBLAS bf16 burst peak; `traffic` is
                      read from profiles/ncu_traffic.json (written from an `ncu --set full` capture, with its commit).
  roofline_hbm        the bandwidth kernel (GroupNorm+SiLU, 128 ch @384x640) vs the measured copy bandwidth.
  unet                whole-UNet algorithmic TFLOP/s inside the timed region vs the sustained peak.
  gpu_eager_baseline  the oracle's torch.nn UNet run by eager PyTorch (cuDNN / cuBLAS, same dtype, channels_last) on the
                      SAME GPU in this run, outside the timed region: the "same-box library" bar.
  cpu_baseline        the oracle port of the reference pipeline timed on the host cores on a bounded sample.
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

DDIM_CFG = dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample",
                clip_sample=True, set_alpha_to_one=False, steps_offset=1)
DDPM_CFG = dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2", prediction_type="sample",
                thresholding=True, clip_sample=True)
METRIC = "disparity maps/sec at 360x640, 10 DDIM steps"

# per-sample UNet forward GFLOP: (Cin, dc, padded H, padded W) -> GFLOP
UNET_GFLOP = {(8, 1, 384, 640): 1872.53, (5, 1, 384, 640): 1870.83, (8, 1, 544, 960): 4001.95}

CONFIGS = {
    # name: pairs/GPU, (h, w), cond_channels, Cin, sampler, steps, guidance weight, dtype override, BASELINE.json index
    "c1": dict(pairs=1, hw=(360, 640), cond="left+right+raw", cin=8, sampler="ddim", n_steps=10, guid=0.0, dtype=None, idx=0),
    "c2": dict(pairs=16, hw=(360, 640), cond="left+right+raw", cin=8, sampler="ddim", n_steps=10, guid=0.0, dtype=None, idx=1),
    "c3": dict(pairs=8, hw=(360, 640), cond="rgb+raw", cin=5, sampler="ddim", n_steps=10, guid=2e-4, dtype=None, idx=2),
    "c4": dict(pairs=1, hw=(540, 960), cond="left+right+raw", cin=8, sampler="ddpm", n_steps=100, guid=0.0, dtype="fp32", idx=3),
    "c5": dict(pairs=32, hw=(360, 640), cond="left+right+raw", cin=8, sampler="ddim", n_steps=10, guid=0.0, dtype=None, idx=4),
}


def pad32(v: int) -> int:
    return (v + 31) // 32 * 32


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two roofline kernels, transcribed from a committed
    `ncu --set full` capture (the file names the capture and the commit it was taken at).  None when absent."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return {}
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


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
def synth_inputs(n: int, first_index: int = 0, h: int = 360, w: int = 640, stereo_consistent: bool = False):
    """seeded per global sample index; images uint8-like in [-1,1], raw depth uniform [0.3,2] m with
    20 % holes.  Host tensors (CPU).  ``stereo_consistent``: the right image is the left one displaced by the true
    disparity (so that the stereo-geometry guidance of config c3 sees a real correspondence), rgb = left."""
    import torch.nn.functional as F
    lefts, rights, depths = [], [], []
    for i in range(n):
        g = torch.Generator().manual_seed(1234 + first_index + i)
        left = (torch.randint(0, 256, (3, h, w), generator=g).float() / 255.0 - 0.5) * 2
        right = (torch.randint(0, 256, (3, h, w), generator=g).float() / 255.0 - 0.5) * 2
        d = torch.rand((1, h, w), generator=g) * 1.7 + 0.3
        holes = torch.rand((1, h, w), generator=g) < 0.2
        if stereo_consistent:
            d = F.interpolate(d[None, :, ::24, ::24], size=(h, w), mode="bilinear", align_corners=True)[0]
            left = F.avg_pool2d(F.pad(left[None], (1, 1, 1, 1), mode="replicate"), 3, 1)[0]
            disp = 24.5471958 / d
            xx, yy = torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy")
            grid = torch.stack([(xx + disp[0]) / ((w - 1) / 2.0) - 1, yy / ((h - 1) / 2.0) - 1], -1)[None]
            right = F.grid_sample(left[None], grid, align_corners=True, mode="bilinear", padding_mode="border")[0]
        d = d.clone()
        d[holes] = 0.0
        lefts.append(left)
        rights.append(right)
        depths.append(d)
    return torch.stack(lefts), torch.stack(rights), torch.stack(depths)


def workload_config(name: str, n_gpus: int, dtype: str):
    c = CONFIGS[name]
    hp, wp = pad32(c["hw"][0]), pad32(c["hw"][1])
    sampler = f"{'DDIM' if c['sampler'] == 'ddim' else 'DDPM'} {c['n_steps']} steps"
    guid = f", stereo-geometry guidance (weight {c['guid']:g})" if c["guid"] > 0 else ""
    return {"workload": f"BASELINE configs[{c['idx']}] ({name}): {c['cond']} variant, batch {c['pairs']} per GPU at "
                        f"{c['hw'][0]}x{c['hw'][1]} (UNet at padded {hp}x{wp}), {sampler}{guid}, random-init "
                        f"UNet2DModel({c['cin']}->1, 113.68 M params), {dtype} operands",
            "config": name, "pairs_per_gpu": c["pairs"], "global_pairs": c["pairs"] * n_gpus,
            "sampler_steps": c["n_steps"],
            "parallelism": f"batch-shard x{n_gpus} (replicated weights, output all-gather)",
            "l2": "per-step working set (>1 GB of activations per layer at batch 16) exceeds the 126 MB L2; no explicit "
                  "flush" if c["pairs"] >= 8 else "small batch: layers of the deep levels fit in the 126 MB L2 (as they "
                  "do in the real single-pair loop); inputs of every step are rewritten by the previous kernel"}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_steps(n_steps: int, n_warm: int, threads: int):
    """Time the oracle port of the reference pipeline (core/custom_pipelines.py loop + restated UNet) on the host:
    each timed unit = 1 stereo pair, ONE DDIM step (UNet forward + scheduler update) at 384x640, fp32.
    Returns (seconds per timed unit, maps/s equivalent = 1 / (10 units))."""
    from oracle import pipeline_ref, prepost_ref as P
    from oracle.unet_ref import make_unet
    torch.set_num_threads(threads)
    unet = make_unet(8, 1, seed=0, sample_size=(384, 640))
    left, right, depth = synth_inputs(1)
    raw = P.normalize_average(P.depth_to_disp(depth[0], 24.5471958), 128.0)[None]
    conds = {"left": P.pad_replicate(left), "right": P.pad_replicate(right), "raw": P.pad_replicate(raw)}
    x_T = torch.randn((1, 1, 384, 640), generator=torch.Generator().manual_seed(42))
    times = []
    for i in range(n_warm + n_steps):
        t0 = time.perf_counter()
        pipeline_ref.run_pipeline(unet, x_T, conds, "left+right+raw", sampler="ddim", num_inference_steps=1)
        dt = time.perf_counter() - t0
        if i >= n_warm:
            times.append(dt)
    sec = sum(times) / len(times)
    return sec, 1.0 / (10 * sec)


CPU_SAMPLE = ("bounded sample of BASELINE configs[1]: each timed step = 1 stereo pair x 1 DDIM step (UNet forward + "
              "scheduler update, fp32, 384x640) out of the workload's 16 pairs x 10 steps; maps/s = 1 / (10 x seconds "
              "per timed step)")


TRAIN_CPU_SAMPLE = ("bounded sample of the training workload: each timed step = ONE optimisation step (add_noise, UNet "
                    "forward, masked MSE, autograd backward, clip_grad_norm_, AdamW) on 1 sample at 384x640, fp32, out of "
                    "the workload's 16 samples per step")


def cpu_train_steps(n_steps: int, n_warm: int, threads: int):
    """The oracle port of the reference's train_step (oracle/train_ref.py) on the host cores: seconds per 1-sample step."""
    from oracle import schedulers_ref as S, train_ref
    from oracle.unet_ref import make_unet
    torch.set_num_threads(threads)
    unet = make_unet(8, 1, seed=0, sample_size=(384, 640))
    opt = torch.optim.AdamW(unet.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(7)
    clean, cond = torch.rand(1, 1, 384, 640, generator=g) * 2 - 1, torch.rand(1, 7, 384, 640, generator=g) * 2 - 1
    noise, t = torch.randn(1, 1, 384, 640, generator=g), torch.randint(0, 128, (1,), generator=g)
    mask = (torch.rand(1, 1, 384, 640, generator=g) > 0.1).float()
    ac = torch.cumprod(1 - S.make_betas("squaredcos_cap_v2", 128), 0)
    times = []
    for i in range(n_warm + n_steps):
        t0 = time.perf_counter()
        train_ref.train_step(unet, opt, clean, cond, noise, t, ac, mask, None, "sample", "mse", 1.0)
        if i >= n_warm:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return sec, 1.0 / sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.config == "train":
        sec, sps = cpu_train_steps(max(1, min(args.steps, 3)), 1, threads)
        emit({"impl": "reference", "metric": "training samples/sec at 360x640 (forward + backward + AdamW step)",
              "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
              "data": "synthetic",
              "config": {"workload": "CPU reference arm, " + TRAIN_CPU_SAMPLE, "config": "train",
                         "timed_unit": {"samples": 1, "optimisation_steps": 1}},
              "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port", "sample": TRAIN_CPU_SAMPLE},
              "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0,
              "note": "reference = train.py:157-538 restated in oracle/train_ref.py over the oracle UNet (accelerate and "
                      "diffusers are un-vendored third-party dependencies); ms_per_step is the wall time of ONE timed unit"})
        return
    sec, maps = cpu_reference_steps(max(1, args.steps), max(1, min(args.warmup, 2)), threads)
    cfg = workload_config("c2", args.gpus, "f32")
    cfg["workload"] = ("CPU reference arm, " + CPU_SAMPLE + " | config it samples: " + cfg["workload"])
    cfg["timed_unit"] = {"pairs": 1, "ddim_steps": 1, "units_per_map": 10}
    line = {
        "impl": "reference", "metric": METRIC, "value": maps, "unit": "maps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": maps, "unit": "maps/s", "cores": threads, "kind": "port", "sample": CPU_SAMPLE},
        "e2e": {"value": maps, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = jie0530/d3roma core/custom_pipelines.py loop restated in oracle/ (diffusers.UNet2DModel "
                "is an un-vendored third-party dependency, so the CPU path is the oracle port, torch CPU fp32); "
                "ms_per_step is the wall time of ONE timed unit (1 pair x 1 DDIM step), value is scaled to maps/s",
    }
    emit(line)


# --------------------------------------------------------------------------------------------- GPU arm
def time_dominant_kernels(dtype, peaks, B=16):
    """Roofline legs: the dominant conv shape and the GroupNorm+SiLU pass, each timed alone with CUDA events on the
    stream the kernels are launched on (torch's current stream)."""
    from d3roma_b200 import ops
    adt = torch.bfloat16 if dtype == "bf16" else torch.float16
    HP, WP = 384, 640
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
    tr = load_ncu_traffic()
    roof = {"bound": "tensor",
            "kernel": "conv_igemm_t_kernel (pixels-as-N, 16x16 halo blocks, stmatrix epilogue) 3x3 128->128 @16x384x640",
            "achieved": flops / t_conv / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
            "frac": flops / t_conv / 1e12 / peaks["tf_burst"],
            "traffic": tr.get("conv", {}).get("dram_bytes"), "traffic_source": tr.get("conv", {}).get("source"),
            "algorithmic_bytes": 2.0 * x.numel() * 2, "flops_per_launch": flops, "ms_per_launch": t_conv * 1e3,
            "peak_source": peaks["src"] + " burst (kernel timed alone)"}
    roof_hbm = {"bound": "hbm", "kernel": "gn_apply_kernel 128ch @16x384x640 (GroupNorm+SiLU)",
                "achieved": gn_bytes / t_gn / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": gn_bytes / t_gn / 1e9 / peaks["hbm"],
                "traffic": tr.get("groupnorm", {}).get("dram_bytes"),
                "traffic_source": tr.get("groupnorm", {}).get("source"),
                "bytes_per_launch": gn_bytes, "ms_per_launch": t_gn * 1e3, "peak_source": peaks["src"]}
    return roof, roof_hbm


def gpu_eager_baseline(dtype: str, B: int, hw, cin: int, n_steps: int):
    """The oracle's torch.nn UNet through eager PyTorch (cuDNN convs, SDPA-free attention as restated) on this GPU in
    the same 16-bit dtype, channels_last: 1 warm-up + 3 timed forwards, outside the timed region of the bench."""
    from oracle.unet_ref import make_unet
    adt = torch.bfloat16 if dtype == "bf16" else torch.float16
    try:
        net = make_unet(cin, 1, seed=0, sample_size=hw).to("cuda", adt).to(memory_format=torch.channels_last)
        x = torch.randn(B, cin, *hw, device="cuda", dtype=adt).to(memory_format=torch.channels_last)
        t = torch.tensor(61, device="cuda")
        with torch.no_grad():
            net(x, t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                net(x, t)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        del net, x
        torch.cuda.empty_cache()
        return {"ms_per_unet_step": ms, "value": B / (ms * 1e-3 * n_steps), "unit": "maps/s",
                "what": f"eager PyTorch {torch.__version__} (cuDNN/cuBLAS) on the oracle's torch.nn UNet, {dtype}, "
                        f"channels_last, batch {B}, UNet forwards only (scheduler excluded), same GPU, this run"}
    except Exception as e:  # noqa: BLE001  (an out-of-memory or cuDNN failure must not kill the bench line)
        return {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}


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
    import d3roma_b200 as d
    from d3roma_b200 import _lib
    from d3roma_b200.evaluate import denoise_batch
    from d3roma_b200.weights import random_init_state_dict
    from d3roma_b200 import metrics as M

    C = CONFIGS[args.config]
    dtype = C["dtype"] or args.dtype
    Bp, (H0, W0), n_steps = C["pairs"], C["hw"], C["n_steps"]
    HP, WP = pad32(H0), pad32(W0)
    guided = C["guid"] > 0
    net = d.UNet2DModel(sample_size=(HP, WP), in_channels=C["cin"], out_channels=1, compute_dtype=dtype).to(dev)
    net.load_state_dict(random_init_state_dict(C["cin"], 1, seed=0))
    sched = (d.MyDDIMScheduler.from_config(DDIM_CFG) if C["sampler"] == "ddim"
             else d.MyDDPMScheduler.from_config(DDPM_CFG))
    guidance = d.FlowGuidance(flow_guidance_weight=C["guid"], flow_guidance_mode="stereo" if guided else "imputation")
    pipe = d.GuidedDiffusionPipeline(net, sched, guidance)
    normalizer = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    cam = d.SIM_CAMERA

    # ---- synthetic inputs: host (pinned) copies for the e2e leg, device-resident padded copies for `value`
    left_h, right_h, depth_h = synth_inputs(Bp, first_index=rank * Bp, h=H0, w=W0, stereo_consistent=guided)
    left_h, right_h, depth_h = left_h.pin_memory(), right_h.pin_memory(), depth_h.pin_memory()
    padder = d.InputPadder(left_h.shape, divis_by=32)
    names = ("rgb", "raw") if C["cond"] == "rgb+raw" else ("left", "right", "raw")

    def prep_device():
        dd = depth_h.to(dev)
        raw = normalizer.normalize(d.prepost.depth_to_disp(dd, cam.fxb).view(Bp, H0, W0))[0]
        l, r, rw = padder.pad(left_h.to(dev), right_h.to(dev), raw.view(Bp, 1, H0, W0))
        srcs = {"rgb": l, "left": l, "right": r, "raw": rw}
        return torch.cat([srcs[n] for n in names], 1).contiguous(), l, r

    cond, left_d, right_d = prep_device()
    gens = [torch.Generator().manual_seed(10_000 + rank * Bp + i) for i in range(Bp)]
    x_T = torch.cat([torch.randn((1, 1, HP, WP), generator=g) for g in gens]).to(dev)
    gather_buf = torch.empty((world * Bp, 1, H0, W0), device=dev) if world > 1 else None
    rows_buf = torch.empty((world * Bp, 11), dtype=torch.float64, device=dev) if world > 1 else None
    want_rows = args.config == "c5"
    gt_disp = d.prepost.depth_to_disp(depth_h.to(dev), cam.fxb).view(Bp, H0, W0) if want_rows else None
    fxb_t = torch.full((Bp,), float(cam.fxb), device=dev) if want_rows else None

    peer = None
    if world > 1 and args.gather == "peer":
        from d3roma_b200.distributed import PeerGather
        regions = {"disp": ((Bp, 1, H0, W0), torch.float32)}
        if want_rows:
            regions["rows"] = ((Bp, 11), torch.float64)
        peer = PeerGather(regions, rank=rank, world=world)

    def one_step():
        cond.mul_(1.0)   # a new batch every step: the per-loop conditioning term of conv_in is recomputed, not cached
        img = pipe.sample_from(x_T, cond, n_steps, left_images=left_d if guided else None,
                               right_images=right_d if guided else None,
                               denorm=normalizer.denormalize if guided else None)
        if peer is not None:
            # the exchange is part of the last kernel: denormalise + stores into every rank's gather buffer over NVLink
            peer.denormalize_into(padder.unpad(img), normalizer)
            if want_rows:
                mine = peer._local_slot("disp")
                peer.push(M.compute_errors_device(gt_disp, mine[:, 0], gt_disp > 0, fxb_t), "rows")
            return peer.finish()["disp"]
        disp = normalizer.denormalize(padder.unpad(img))
        rows = None
        if want_rows:   # distributed_evaluate.py:153-163: per-sample metric rows gathered with the maps
            rows = M.compute_errors_device(gt_disp, disp[:, 0], gt_disp > 0, fxb_t)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, disp)
            if rows is not None:
                dist.all_gather_into_tensor(rows_buf, rows)
        return disp

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(3, args.warmup)
    for _ in range(warm):
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
    value = world * Bp * args.steps / sec

    # ---- e2e: the public call chain with HOST buffers (pad -> pipeline.__call__ -> unpad -> denormalise -> D2H)
    def e2e_step():
        dd = depth_h.to(dev, non_blocking=True)
        raw = normalizer.normalize(d.prepost.depth_to_disp(dd, cam.fxb).view(Bp, H0, W0))[0]
        kw = dict(rgb=left_h, sim_disp=raw.view(Bp, 1, H0, W0)) if C["cond"] == "rgb+raw" else \
            dict(left=left_h, right=right_h, sim_disp=raw.view(Bp, 1, H0, W0))
        if guided:
            kw.update(left=left_h, right=right_h)
        return denoise_batch(pipe, normalizer, cond_channels=C["cond"], num_inference_steps=n_steps,
                             num_intermediate_images=1, generator=gens, to_host=True, **kw)

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
    e2e_value = world * Bp * n_e2e / float(te.item())
    n_img = 2 if (C["cond"] != "rgb+raw" or guided) else 1
    h2d = (left_h.numel() * n_img + depth_h.numel() + x_T.numel()) * 4
    d2h = out_h.numel() * 4

    if rank == 0:
        gflop = UNET_GFLOP.get((C["cin"], 1, HP, WP))
        unet_tf = (Bp * n_steps * gflop * 1e9 * args.steps / sec / 1e12) if gflop else None
        line = {
            "metric": METRIC if args.config != "c4" else "disparity maps/sec at 540x960, 100 DDPM steps (fp32 path)",
            "value": value, "unit": "maps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(args.config, world, dtype),
            "ms_per_unet_step": sec / args.steps / n_steps * 1e3,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "maps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e, "api": "d3roma_b200.evaluate.denoise_batch -> GuidedDiffusionPipeline.__call__"},
            "gpu_launches": launches,
        }
        if world > 1:
            line["config"]["gather"] = ("stores into NVLink peer memory fused into the denormalise kernel + flag handshake "
                                        "(d3roma_b200.distributed.PeerGather)" if peer is not None else
                                        "NCCL all_gather_into_tensor after the loop")
        if unet_tf is not None and dtype != "fp32":
            line["unet"] = {"algorithmic_tflops": unet_tf, "peak": peaks["tf_sust"],
                            "frac_of_sustained_peak": unet_tf / peaks["tf_sust"], "gflop_per_sample_step": gflop,
                            "peak_source": peaks["src"] + " sustained (kernels timed inside a long step)"}
        if dtype != "fp32":
            roof, roof_hbm = time_dominant_kernels(dtype, peaks)
            line["roofline"], line["roofline_hbm"] = roof, roof_hbm
            if not args.no_gpu_baseline:
                line["gpu_eager_baseline"] = gpu_eager_baseline(dtype, Bp, (HP, WP), C["cin"], n_steps)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sec_cpu, maps_cpu = cpu_reference_steps(3, 1, threads)
            line["cpu_baseline"] = {"value": maps_cpu, "unit": "maps/s", "cores": threads, "kind": "port",
                                    "sample": CPU_SAMPLE + f" (3 timed steps of {sec_cpu:.2f} s)"}
        emit(line)
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- latent family (f1)
LDM_SCHED = dict(num_train_timesteps=1000, beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012,
                 prediction_type="v_prediction", clip_sample=False, thresholding=False)
LDM_PAIRS, LDM_STEPS, LDM_PLOTS = 16, 10, 5


def ldm_eager_baseline(dtype: str, Bp: int, hw, n_steps: int, plots: int):
    """The oracle's torch.nn restatements of the SD-2.1 UNet and the SD VAE run by eager PyTorch (cuDNN / cuBLAS / SDPA
    math, same 16-bit dtype, channels_last) on the SAME GPU, outside the timed region: one UNet forward, one encoder
    pass, one decoder pass at the workload's batch, combined with the loop's call counts."""
    from oracle.ldm_ref import AutoencoderKL, UNet2DConditionModel
    tdt = torch.float16 if dtype == "fp16" else torch.bfloat16
    h, w = hw[0] // 8, hw[1] // 8
    with torch.device("cuda"):
        unet = UNet2DConditionModel(in_channels=12).to(tdt).to(memory_format=torch.channels_last).eval()
        vae = AutoencoderKL().to(tdt).to(memory_format=torch.channels_last).eval()

    def t_of(fn, reps=2):
        with torch.no_grad():
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    x = torch.randn(Bp, 12, h, w, device="cuda", dtype=tdt).contiguous(memory_format=torch.channels_last)
    ctx = torch.randn(Bp, 2, 1024, device="cuda", dtype=tdt)
    img = torch.randn(2 * Bp, 3, hw[0], hw[1], device="cuda", dtype=tdt).contiguous(memory_format=torch.channels_last)
    z = torch.randn(Bp, 4, h, w, device="cuda", dtype=tdt)
    ms_u = t_of(lambda: unet(x, 500, ctx))
    ms_e = t_of(lambda: vae.quant_conv(vae.encoder(img)))
    ms_d = t_of(lambda: vae.decoder(vae.post_quant_conv(z)))
    ms = n_steps * ms_u + ms_e + plots * ms_d
    del unet, vae
    torch.cuda.empty_cache()
    return {"value": Bp / (ms * 1e-3), "unit": "maps/s", "ms_unet_forward": ms_u, "ms_vae_encode_2B": ms_e,
            "ms_vae_decode_B": ms_d, "kind": "eager PyTorch (cuDNN/cuBLAS) on the same GPU, oracle torch.nn modules, "
            f"{dtype}, channels_last; loop time composed as {n_steps} UNet + 1 encode (rgb+raw = 2B images) + "
            f"{plots} decode passes", "data": "synthetic"}


def run_ldm_arm(args):
    """--config ldm: the reference's latent family (GuidedLatentDiffusionPipeline, custom_pipelines.py:208-460) in the
    shape of conf/task/eval_ldm_mixed_rgb+raw.yaml - rgb+raw conditioning, SD-2.1 UNet2DConditionModel (12-channel
    conv_in), SD AutoencoderKL, DDPM sampler with 10 inference steps of 1000 (v-prediction, scaled-linear betas), 5
    decoded intermediate images - at the BASELINE metric's 360x640 with 16 pairs per GPU."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import d3roma_b200 as d
    from d3roma_b200 import _lib
    from d3roma_b200.autoencoder_kl import AutoencoderKL
    from d3roma_b200.unet_condition import UNet2DConditionModel
    from d3roma_b200.weights import random_init_cond_unet_state_dict, random_init_vae_state_dict
    dtype = args.dtype
    Bp, (H0, W0), n_steps, plots = LDM_PAIRS, (360, 640), LDM_STEPS, LDM_PLOTS
    unet = UNet2DConditionModel(in_channels=12, compute_dtype=dtype).to(dev)
    unet.load_state_dict(random_init_cond_unet_state_dict(12, seed=0, device=dev))
    vae = AutoencoderKL(compute_dtype=dtype).to(dev)
    vae.load_state_dict(random_init_vae_state_dict(seed=1, device=dev))
    pipe = d.GuidedLatentDiffusionPipeline(unet, vae, None, None, d.MyDDPMScheduler.from_config(LDM_SCHED),
                                           d.FlowGuidance(flow_guidance_weight=0.0, flow_guidance_mode="gradient"))
    normalizer = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    cam = d.SIM_CAMERA
    ete = torch.randn(1, 2, 1024, generator=torch.Generator().manual_seed(3)).to(dev)   # empty-prompt embedding stand-in
    left_h, _, depth_h = synth_inputs(Bp, first_index=rank * Bp, h=H0, w=W0)
    left_h, depth_h = left_h.pin_memory(), depth_h.pin_memory()
    gens = [torch.Generator().manual_seed(10_000 + rank * Bp + i) for i in range(Bp)]
    gather_buf = torch.empty((world * Bp, 1, H0, W0), device=dev) if world > 1 else None

    def raw_of(depth_dev):
        return normalizer.normalize(d.prepost.depth_to_disp(depth_dev, cam.fxb).view(Bp, H0, W0))[0].view(Bp, 1, H0, W0)

    rgb_d, raw_d = left_h.to(dev), raw_of(depth_h.to(dev))

    def loop(rgb, raw):
        out = pipe(rgb, None, None, raw, None, None, empty_text_embed=ete, num_inference_steps=n_steps,
                   num_intermediate_images=plots, cond_channels="rgb+raw", denorm=normalizer.denormalize, generator=gens)
        disp = normalizer.denormalize(out.images)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, disp)
        return disp

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(3, args.warmup)
    for _ in range(warm):
        loop(rgb_d, raw_d)
    sync()
    lib = _lib.lib()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    lib.d3_reset_launch_count()
    g0 = unet.graph_kernel_launches
    e0.record()
    for _ in range(args.steps):
        loop(rgb_d, raw_d)
    e1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    eager_launches = int(lib.d3_launch_count()) + (unet.graph_kernel_launches - g0)
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    value = world * Bp * args.steps / sec

    def e2e_step():
        rgb = left_h.to(dev, non_blocking=True)
        return loop(rgb, raw_of(depth_h.to(dev, non_blocking=True))).to("cpu")

    for _ in range(2):
        e2e_step()
    sync()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        out_h = e2e_step()
    sync()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Bp * n_e2e / float(te.item())

    if rank == 0:
        # component times (CUDA events, after the timed region): one UNet forward (graph replay), encode, decode
        def t_of(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        x = torch.randn(Bp, 12, H0 // 8, W0 // 8, device=dev)
        ctx = ete.repeat(Bp, 1, 1)
        img2 = torch.cat([rgb_d, raw_d.repeat(1, 3, 1, 1)])
        z = torch.randn(Bp, 4, H0 // 8, W0 // 8, device=dev) * 0.18215
        ms_u = t_of(lambda: unet(x, 500, ctx))
        ms_e = t_of(lambda: vae.encode_latent(img2, 0.18215))
        ms_d = t_of(lambda: vae.decode_mean(z, 0.18215))
        g = next(iter(unet._graphs.values()), None)
        line = {
            "metric": "disparity maps/sec at 360x640, 10 DDPM steps (latent family: SD-2.1 UNet + VAE)",
            "value": value, "unit": "maps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": "latent family (SURVEY 8 f1; reference conf/task/eval_ldm_mixed_rgb+raw.yaml loop shape): "
                                   f"rgb+raw variant, batch {Bp} per GPU at {H0}x{W0} (latents {H0 // 8}x{W0 // 8}), DDPM "
                                   f"{n_steps} of 1000 steps, v-prediction, {plots} decoded intermediate images, random-init "
                                   "UNet2DConditionModel (SD-2.1 layout, 12-channel conv_in, 865.9 M params) + AutoencoderKL "
                                   f"(SD VAE, 83.7 M params), {dtype} operands",
                       "config": "ldm", "pairs_per_gpu": Bp, "global_pairs": Bp * world, "sampler_steps": n_steps,
                       "parallelism": f"batch-shard x{world} (replicated weights, output all-gather)",
                       "l2": "weights alone (1.9 GB in fp16) exceed the 126 MB L2; no explicit flush"},
            "ms_per_unet_step": ms_u, "ms_vae_encode_2B_images": ms_e, "ms_vae_decode_B_images": ms_d,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "maps/s", "h2d_bytes_per_step": (left_h.numel() + depth_h.numel()) * 4
                    + Bp * 4 * (H0 // 8) * (W0 // 8) * 4, "d2h_bytes_per_step": out_h.numel() * 4, "steps": n_e2e,
                    "api": "d3roma_b200.GuidedLatentDiffusionPipeline.__call__ (host buffers in, host maps out)"},
            "gpu_launches": eager_launches,
        }
        if not args.no_gpu_baseline:
            line["gpu_eager_baseline"] = ldm_eager_baseline(dtype, Bp, (H0, W0), n_steps, plots)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


TRAIN_PAIRS = 16     # samples per GPU and optimisation step (reference default train_batch_size is 16 per device)


def train_eager_baseline(dtype: str, B: int, hw, cin: int, n_steps: int):
    """The same optimisation step under eager PyTorch (autograd + cuDNN / cuBLAS, autocast to ``dtype``, channels_last,
    torch.optim.AdamW, clip_grad_norm_) on the oracle's torch.nn UNet, same GPU, this run."""
    from oracle.unet_ref import make_unet
    adt = torch.float16 if dtype == "fp16" else torch.bfloat16
    dev = torch.device("cuda", torch.cuda.current_device())
    try:
        net = make_unet(cin, 1, seed=0, sample_size=hw).to(dev).to(memory_format=torch.channels_last)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
        scaler = torch.amp.GradScaler("cuda", enabled=dtype == "fp16")
        x = torch.randn(B, cin, *hw, device=dev).to(memory_format=torch.channels_last)
        tgt = torch.randn(B, 1, *hw, device=dev)
        t = torch.randint(0, 128, (B,), device=dev)

        def step():
            with torch.autocast("cuda", dtype=adt):
                out = net(x, t)
                out = out.sample if hasattr(out, "sample") else out
            loss = torch.nn.functional.mse_loss(out.float(), tgt)
            scaler.scale(loss).backward()
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
            scaler.step(opt)
            scaler.update()
            opt.zero_grad(set_to_none=True)
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_steps
        del net, opt, x
        torch.cuda.empty_cache()
        return {"ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "samples/s",
                "what": f"eager PyTorch {torch.__version__} (autograd, cuDNN/cuBLAS, autocast {dtype}, channels_last, AdamW, "
                        f"clip_grad_norm_) on the oracle's torch.nn UNet, batch {B}, same GPU, this run"}
    except Exception as e:   # noqa: BLE001 - the baseline is informative; never let it take the bench line down
        torch.cuda.empty_cache()
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


def run_train_arm(args):
    """--config train: the reference's training step for the pixel-space model (train.py:157-538) -
    add_noise, UNet forward with one timestep per sample, masked MSE loss, backward, gradient all-reduce across the GPUs
    (DDP), clip_grad_norm_, AdamW - on 16 synthetic left+right+raw samples per GPU at 360x640 (padded 384x640)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import d3roma_b200 as d
    from d3roma_b200 import _lib, ops
    from d3roma_b200.training import UNetTrainer
    from d3roma_b200.weights import random_init_state_dict
    dtype = args.dtype if args.dtype_given else "bf16"
    B, (H, W) = TRAIN_PAIRS, (pad32(360), pad32(640))
    tr = UNetTrainer(in_channels=8, out_channels=1, compute_dtype=dtype, lr=1e-4, max_grad_norm=1.0,
                     loss_scale=65536.0 if dtype == "fp16" else 1.0, device=dev)
    tr.load_state_dict(random_init_state_dict(8, 1, seed=0))
    sched = d.MyDDPMScheduler.from_config(dict(num_train_timesteps=128, beta_schedule="squaredcos_cap_v2",
                                               prediction_type="sample", clip_sample=True))
    g = torch.Generator().manual_seed(100 + rank)
    host = {"clean": torch.rand(B, 1, H, W, generator=g) * 2 - 1, "cond": torch.rand(B, 7, H, W, generator=g) * 2 - 1,
            "noise": torch.randn(B, 1, H, W, generator=g), "mask": (torch.rand(B, 1, H, W, generator=g) > 0.1).float()}
    host = {k: v.pin_memory() for k, v in host.items()}
    t_host = torch.randint(0, 128, (B,), generator=g).pin_memory()
    devt = {k: v.to(dev) for k, v in host.items()}
    t_dev = t_host.to(dev)

    def step(x, t):
        return tr.train_step(x["clean"], x["cond"], x["noise"], t, sched, mask=x["mask"], prediction_type="sample",
                             world_size=world)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(3, args.warmup)
    losses = []
    for _ in range(warm):
        losses.append(step(devt, t_dev))
    sync()
    lib = _lib.lib()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    lib.d3_reset_launch_count()
    e0.record()
    for _ in range(args.steps):
        losses.append(step(devt, t_dev))
    e1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.d3_launch_count())
    tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec = float(tt.item())
    value = world * B * args.steps / sec

    def e2e_step():
        x = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return float(step(x, t_host.to(dev, non_blocking=True)).item())      # D2H read of the loss

    e2e_step()
    sync()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        last_loss = e2e_step()
    sync()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * n_e2e / float(te.item())

    if rank == 0:
        def t_of(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        # phases of one step (CUDA events, after the timed region)
        noisy = sched.add_noise(devt["clean"], devt["noise"], t_dev)
        ms_fwd = t_of(lambda: tr.forward(noisy, devt["cond"], t_dev))
        pred = tr.forward(noisy, devt["cond"], t_dev)
        dy, _ = ops.masked_loss_grad(pred, devt["clean"], devt["mask"], None, grad_scale=tr.loss_scale / pred.numel())

        def fb():
            tr.zero_grad()
            tr.forward(noisy, devt["cond"], t_dev)
            tr.backward(dy)
        ms_fb = t_of(fb)
        ms_opt = t_of(lambda: tr.optimizer_step(1))
        # dominant backward kernel timed alone: weight gradient of a 3x3 128->128 conv at the full resolution
        adt = torch.float16 if dtype == "fp16" else torch.bfloat16
        xa = torch.randn(B, H, W, 128, device=dev).to(adt)
        ga = torch.randn(B, H, W, 128, device=dev).to(adt)
        gw = torch.zeros(128, 9 * 128, device=dev)
        ms_wg = t_of(lambda: ops.conv2d_wgrad(xa, ga, gw, ksize=3), reps=10)
        flops = 2.0 * B * H * W * 9 * 128 * 128
        peaks = load_peaks()
        peak = float(peaks["tf_burst"])
        unet_tflop = 3 * 1872.53e9 * B / 1e12    # forward + data gradients + weight gradients, per step and GPU
        line = {
            "metric": "training samples/sec at 360x640 (forward + backward + AdamW step)",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": "training step of the pixel-space denoiser (SURVEY 8 f4; reference train.py:157-538): "
                                   f"left+right+raw variant, {B} samples per GPU at 360x640 (UNet at padded {H}x{W}), one "
                                   "timestep per sample, masked MSE on x0 ('sample' prediction), clip_grad_norm_ 1.0, AdamW "
                                   f"lr 1e-4, random-init UNet2DModel(8->1, 113.68 M params), {dtype} activations / "
                                   "gradients with fp32 master weights and moments",
                       "config": "train", "samples_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"data parallel x{world} (replicated weights, NCCL all-reduce of the flat fp32 "
                                      "gradient buffer)",
                       "l2": "per-step working set (>1 GB per activation at batch 16) exceeds the 126 MB L2; no explicit flush"},
            "ms_forward": ms_fwd, "ms_forward_backward": ms_fb, "ms_optimizer": ms_opt,
            "loss_first_last": [float(losses[0]), float(losses[-1])],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s",
                    "h2d_bytes_per_step": sum(v.numel() for v in host.values()) * 4 + t_host.numel() * 8,
                    "d2h_bytes_per_step": 8, "steps": n_e2e, "last_loss": last_loss,
                    "api": "d3roma_b200.training.UNetTrainer.train_step (pinned host batch in, loss out)"},
            "gpu_launches": launches,
            "unet": {"algorithmic_tflops": unet_tflop * args.steps / sec, "peak": float(peaks["tf_sust"]),
                     "frac_of_sustained_peak": unet_tflop * args.steps / sec / float(peaks["tf_sust"]),
                     "gflop_per_sample_step": 3 * 1872.53, "peak_source": "measured sustained (cuBLAS bf16 inside a long step)"},
            "roofline": {"bound": "tensor", "kernel": f"conv_wgrad_kernel (split-K, MN-major TMA operands) 3x3 128->128 @{B}x{H}x{W}",
                         "achieved": flops / (ms_wg * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (ms_wg * 1e-3) / 1e12 / peak,
                         "traffic": load_ncu_traffic().get("wgrad", {}).get("dram_bytes"),
                         "traffic_source": load_ncu_traffic().get("wgrad", {}).get("source"),
                         "algorithmic_bytes": 2.0 * B * H * W * 128 * 2, "flops_per_launch": flops,
                         "ms_per_launch": ms_wg,
                         "peak_source": ("measured" if peaks["src"] == "measured" else "fallback") + " burst (kernel timed alone)"},
        }
        if not args.no_gpu_baseline:
            del xa, ga
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = train_eager_baseline(dtype, B, (H, W), 8, 2)
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sec_c, sps_c = cpu_train_steps(2, 1, threads)
            line["cpu_baseline"] = {"value": sps_c, "unit": "samples/s", "cores": threads, "kind": "port",
                                    "sample": TRAIN_CPU_SAMPLE + f" (2 timed steps of {sec_c:.2f} s)"}
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + ["ldm", "train"])
    ap.add_argument("--dtype", default=None, choices=["fp16", "bf16"])
    ap.add_argument("--gather", default="nccl", choices=["nccl", "peer"],
                    help="output gather of the N > 1 runs: NCCL all_gather_into_tensor after the loop, or stores into NVLink "
                         "peer memory fused into the denormalise kernel (d3roma_b200.distributed.PeerGather)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.dtype_given = args.dtype is not None      # --config train defaults to bf16, everything else to fp16
    if args.dtype is None:
        args.dtype = "fp16"
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config == "ldm":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the d3roma_b200 path has no CPU fallback)")
        run_ldm_arm(args)
    elif args.config == "train":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the d3roma_b200 path has no CPU fallback)")
        run_train_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

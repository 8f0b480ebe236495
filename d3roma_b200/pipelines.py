"""``GuidedDiffusionPipeline`` — drop-in for the reference's pixel-space denoising loop
(``/root/reference/core/custom_pipelines.py:90-205``): same constructor, same ``__call__`` signature and positional
order, same ``GuidedPipelineOutput`` fields, same RNG consumption (x_T drawn with ``randn_tensor`` on the
generator's device then moved, :117-123), same assertions and ``NotImplementedError`` for unknown
``cond_channels`` (:158-171).

What differs is how the loop body executes:
  * the conditioning channels are concatenated ONCE (they are constant across steps) and never concatenated with the
    noisy sample — the conv_in front end gathers from both tensors (``d3_im2col_input``);
  * ``unet(model_input, t)`` is one CUDA-graph replay of the sm_100a kernels (``UNet2DModel.forward_graphed``);
  * ``scheduler.step`` is one fused kernel; the five intermediate-image stacks the reference fills with identical
    ``image.clamp(-1, 1)`` copies (:176-182) are one tensor written in place and aliased five times.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import torch

from . import ops
from ._config import BaseOutput
from .scheduler_ddim import randn_tensor

COND_ORDER = {
    "rgb": ("rgb",),
    "rgb+raw": ("rgb", "raw"),
    "rgb+right": ("left", "right"),          # sic: the reference concatenates left+right for this key (:162-163)
    "left+right+raw": ("left", "right", "raw"),
    "rgb+left+right": ("rgb", "left", "right"),
    "rgb+left+right+raw": ("rgb", "left", "right", "raw"),
}


@dataclass
class GuidedPipelineOutput(BaseOutput):
    """custom_pipelines.py:63-85 (field names kept, including the reference's spelling)."""

    images: torch.Tensor
    images_pred_orig: Optional[torch.Tensor] = None
    images_perturbed_orig: Optional[torch.Tensor] = None
    images_pred_prev: Optional[torch.Tensor] = None
    images_purturbed_pred_prev: Optional[torch.Tensor] = None
    images_sampled_prev: Optional[torch.Tensor] = None


class GuidedDiffusionPipeline:
    def __init__(self, unet, scheduler, guidance):
        self.register_modules(unet=unet, scheduler=scheduler, guidance=guidance)
        self._progress_bar_config = {}
        self.use_cuda_graph = True
        self.graph_whole_loop = True    # deterministic DDIM loops: ONE graph for all steps (UNet forwards + scheduler steps)
        self._loop_graphs = {}

    # ---- DiffusionPipeline-ish plumbing used by the callers (inference.py:71-91,215; evaluate.py:314-317)
    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self) -> torch.device:
        return self.unet.device

    def to(self, device):
        self.unet.to(device)
        return self

    def save_pretrained(self, path):
        """diffusers directory layout: model_index.json, unet/, scheduler/, guidance/ (train.py:682, 701)."""
        import json
        import os
        from ._config import save_config
        os.makedirs(path, exist_ok=True)
        self.unet.save_pretrained(os.path.join(path, "unet"))
        save_config(self.scheduler, os.path.join(path, "scheduler"), "scheduler_config.json")
        g = self.guidance
        os.makedirs(os.path.join(path, "guidance"), exist_ok=True)
        with open(os.path.join(path, "guidance", "config.json"), "w") as f:
            json.dump({"_class_name": "FlowGuidance", "flow_guidance_weight": g.flow_guidance_weight,
                       "perturb_start_ratio": g.perturb_start_ratio, "flow_guidance_mode": g.flow_guidance_mode}, f)
        with open(os.path.join(path, "model_index.json"), "w") as f:
            json.dump({"_class_name": "GuidedDiffusionPipeline", "unet": ["d3roma_b200", "UNet2DModel"],
                       "scheduler": ["d3roma_b200", type(self.scheduler).__name__],
                       "guidance": ["d3roma_b200", "FlowGuidance"]}, f, indent=2)

    @classmethod
    def from_pretrained(cls, path, compute_dtype=None, device="cuda"):
        """evaluate.py:314 / inference.py:73: ``GuidedDiffusionPipeline.from_pretrained(path).to("cuda")``."""
        import os
        from ._config import load_config
        from .guidance import FlowGuidance
        from .scheduler_ddim import MyDDIMScheduler
        from .scheduler_ddpm import MyDDPMScheduler
        from .unet import UNet2DModel
        unet = UNet2DModel.from_pretrained(os.path.join(path, "unet"), compute_dtype=compute_dtype, device=device)
        scfg = load_config(os.path.join(path, "scheduler"), "scheduler_config.json")
        sched_cls = MyDDPMScheduler if "DDPM" in scfg.get("_class_name", "DDIM") else MyDDIMScheduler
        scheduler = sched_cls.from_config(scfg)
        gpath = os.path.join(path, "guidance", "config.json")
        gcfg = load_config(os.path.join(path, "guidance"), "config.json") if os.path.exists(gpath) else {}
        # the reference's FlowGuidance registers no config (guidance.py:94 `@register_to_config` is commented out), so the
        # config.json it saves holds only bookkeeping keys (_class_name, _diffusers_version, _name_or_path): keep what
        # the constructor knows, like ConfigMixin.from_config does
        import inspect
        known = set(inspect.signature(FlowGuidance.__init__).parameters) - {"self", "camera"}
        return cls(unet, scheduler, FlowGuidance(**{k: v for k, v in gcfg.items() if k in known}))

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable):
        if self._progress_bar_config.get("disable", True) and "desc" not in self._progress_bar_config:
            return iterable
        try:
            from tqdm.auto import tqdm
            return tqdm(iterable, **self._progress_bar_config)
        except Exception:  # pragma: no cover
            return iterable

    # ---- the loop body, on device-resident tensors
    @torch.no_grad()
    def _loop(self, image, cond, srcs, build_cond, num_inference_steps, plot_every, add_noise_rgb, raw_depth,
              denorm=None):
        """custom_pipelines.py:134-196 for tensors already on the device. Returns (clamped image, intermediates)."""
        dev = image.device
        batch_size, depth_channels = image.shape[:2]
        self.scheduler.set_timesteps(num_inference_steps)
        guidance = self.guidance
        guided = (guidance is not None and getattr(guidance, "flow_guidance_mode", None) == "stereo"
                  and guidance.flow_guidance_weight > 0.0 and srcs.get("left") is not None
                  and srcs.get("right") is not None)
        if guided:
            guidance.num_train_timesteps = self.scheduler.config.num_train_timesteps
            if denorm is not None:
                guidance.denormer = denorm   # the DDPM hook passes no denormer (scheduler_ddpm.py:507-511)
            import inspect
            step_kw = {"denormer": denorm} if "denormer" in inspect.signature(self.scheduler.step).parameters else {}
        n_plots = num_inference_steps // plot_every
        if self._whole_loop_ok(image, cond, add_noise_rgb, guided):
            return self._loop_graphed(image, cond, num_inference_steps, plot_every)
        stack = torch.empty((batch_size, depth_channels * n_plots, image.shape[2], image.shape[3]),
                            dtype=torch.float32, device=dev)
        fwd = self.unet.forward_graphed if self.use_cuda_graph else self.unet.forward_split
        step = 0
        for t in self.progress_bar(self.scheduler.timesteps):
            if add_noise_rgb:
                rgb = srcs["rgb"]
                noise_rgb = torch.randn(rgb.shape).to(rgb.device)
                timesteps = (t * torch.ones(rgb.shape[0]).to(rgb.device)).long()
                noisy_rgb = self.scheduler.add_noise(rgb, noise_rgb, timesteps)
                cond = build_cond(rgb * 0.5 + noisy_rgb * 0.5)
            model_output = fwd(image, cond, t)
            if guided:
                g = self.scheduler.step(model_output, t, image, guidance=guidance, left_image=srcs["left"],
                                        right_image=srcs["right"], raw_depth=raw_depth, **step_kw)
            else:
                g = self.scheduler.step(model_output, t, image)
            image = g.prev_sample
            if (step + 1) % plot_every == 0:
                ops.clamp_into(image, stack, (step + 1) // plot_every - 1)
            step += 1
        image = ops.clamp_into(image, torch.empty_like(image), 0)
        return image, stack

    # ---- whole-loop CUDA graph ------------------------------------------------------------------------------------
    def _whole_loop_ok(self, image, cond, add_noise_rgb, guided) -> bool:
        """The unguided DDIM loop is a fixed sequence of kernels (no host-side randomness, every scalar known up front:
        the timestep embeddings and the scheduler coefficients of all steps are baked into the graph), so all N x (UNet
        forward + scheduler step) replay as ONE graph launch - no per-step host work at all (the B = 1 latency case)."""
        from .scheduler_ddim import MyDDIMScheduler
        u = self.unet
        return (self.use_cuda_graph and self.graph_whole_loop and not add_noise_rgb and not guided and cond is not None
                and type(self.scheduler) is MyDDIMScheduler and hasattr(u, "forward_split")
                and getattr(u, "compute_dtype", "fp32") != "fp32" and image.shape[1] <= 3 and image.is_cuda
                and self._progress_bar_config.get("disable", True) and "desc" not in self._progress_bar_config)

    def _loop_graphed(self, image, cond, num_inference_steps, plot_every):
        from . import _lib
        from .unet import _StatsArena
        u, sch = self.unet, self.scheduler
        B, dc, H, W = image.shape
        key = self._loop_key(image.shape, cond, num_inference_steps, plot_every)
        g = self._loop_graphs.get(key)
        if g is None:
            for k in [k for k in self._loop_graphs if k[:3] == key[:3]]:   # same shapes, stale weights / scheduler
                del self._loop_graphs[k]
            n_plots = num_inference_steps // plot_every
            g = {"x": image.clone(), "c": cond.clone(), "arena": _StatsArena(image.device, B * 2 * 40000),
                 "ts": [torch.full((1,), float(t), dtype=torch.float32, device=image.device) for t in sch.timesteps],
                 "stack": torch.empty((B, dc * n_plots, H, W), dtype=torch.float32, device=image.device),
                 "img": torch.empty_like(image)}
            g["tb"] = [u.time_bias(t) for t in g["ts"]]    # all steps' time-embedding rows, once per graph
            g["base"] = u.cond_base(g["c"], dc, out=torch.empty((B, H, W, 128), dtype=u._base_dtype, device=image.device))
            g["cond_key"], g["cond_ref"] = u._cond_key(cond), cond   # the tensor is held: see UNet2DModel._cond_key

            def run():
                x = g["x"]
                for i, t in enumerate(sch.timesteps):
                    # the DDIM update rides in the UNet's last kernel (d3_conv_out_fused_ddim): same bits as sch.step(m, t, x)
                    x, _ = u.forward_split(x, g["c"], g["ts"][i], cond_base=g["base"], arena=g["arena"], tb=g["tb"][i],
                                           ddim=sch.step_coeffs(int(t)))
                    if (i + 1) % plot_every == 0:
                        ops.clamp_into(x, g["stack"], (i + 1) // plot_every - 1)
                ops.clamp_into(x, g["img"], 0)

            side = torch.cuda.Stream(device=image.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                u.forward_split(g["x"], g["c"], g["ts"][0], cond_base=g["base"], arena=g["arena"])   # warm-up
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.lib().d3_launch_count()
            with torch.cuda.graph(graph):
                run()
            g["n_kernels"] = int(_lib.lib().d3_launch_count() - n0)
            g["graph"] = graph
            self._loop_graphs[key] = g
        g["x"].copy_(image)
        self._refresh_loop_cond(g, cond, dc)
        g["graph"].replay()
        u.graph_kernel_launches += g["n_kernels"]
        return g["img"].clone(), g["stack"].clone()

    def _loop_key(self, image_shape, cond, num_inference_steps, plot_every):
        u, sch = self.unet, self.scheduler
        return (tuple(image_shape), tuple(cond.shape), cond.device.index, num_inference_steps, plot_every, id(u),
                getattr(u, "_weights_epoch", 0), id(sch), tuple(sorted((k, str(v)) for k, v in dict(sch.config).items())))

    def _refresh_loop_cond(self, g, cond, dc):
        ck = self.unet._cond_key(cond)
        if g["cond_key"] != ck:
            g["c"].copy_(cond)
            g["cond_key"], g["cond_ref"] = ck, cond
            self.unet.cond_base(g["c"], dc, out=g["base"])

    def reset_graphs(self):
        self._loop_graphs.clear()
        if hasattr(self.unet, "reset_graphs"):
            self.unet.reset_graphs()

    @torch.no_grad()
    def sample_from(self, x_T: torch.Tensor, cond: torch.Tensor, num_inference_steps: int = 10,
                    left_images=None, right_images=None, denorm=None) -> torch.Tensor:
        """Device-resident entry: start noise ``x_T`` [B,dc,H,W] and pre-concatenated conditioning ``cond``
        [B,Cc,H,W] already in HBM (fp32, contiguous). Returns the clamped final sample.  ``left_images`` /
        ``right_images`` (+ ``denorm``, the x0 -> pixels map) enable stereo guidance when the guidance object asks."""
        srcs = {"left": left_images, "right": right_images}
        image, _ = self._loop(x_T, cond, srcs, None, num_inference_steps, num_inference_steps, False, None, denorm)
        return image

    @torch.no_grad()
    def __call__(self, rgb_images=None, left_images=None, right_images=None, sim_disp=None, raw_depth=None, mask=None,
                 num_inference_steps: int = 128, num_intermediate_images: int = 128, add_noise_rgb: bool = False,
                 depth_channels: int = 1, cond_channels: str = "rgb", denorm: Callable = None,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None):
        """ in & out images are preferably on the same device """
        assert rgb_images is not None or left_images is not None
        assert num_intermediate_images > 0 and num_inference_steps % num_intermediate_images == 0
        plot_every = num_inference_steps // num_intermediate_images
        image_shape = rgb_images.shape if rgb_images is not None else left_images.shape
        batch_size = image_shape[0]
        ss = self.unet.sample_size
        if type(ss) == int:
            shape = (batch_size, depth_channels, ss, ss)
        elif type(ss) == tuple or type(ss) == list:
            shape = (batch_size, depth_channels, image_shape[-2], image_shape[-1])
        else:
            raise ValueError("sample_size must be int or tuple of ints")
        dev = self.device
        srcs = {"rgb": rgb_images, "left": left_images, "right": right_images, "raw": sim_disp}
        srcs = {k: (None if v is None else v.to(dev, torch.float32)) for k, v in srcs.items()}
        if cond_channels not in COND_ORDER:
            raise NotImplementedError
        names = COND_ORDER[cond_channels]
        if any(srcs[n] is None for n in names):
            raise ValueError(f"cond_channels={cond_channels!r} needs {names}")

        def build_cond(rgb_override=None):
            parts = [(rgb_override if (n == "rgb" and rgb_override is not None) else srcs[n]) for n in names]
            return torch.cat(parts, dim=1).contiguous() if len(parts) > 1 else parts[0].contiguous()

        cond = None if add_noise_rgb else build_cond()
        if cond is not None and hasattr(self.unet, "prime_conditioning"):
            # the GPU starts on the per-loop conditioning work while the host still draws x_T
            g = self._loop_graphs.get(self._loop_key(shape, cond, num_inference_steps, plot_every)) \
                if self.use_cuda_graph and self.graph_whole_loop else None
            if g is not None:
                self._refresh_loop_cond(g, cond, shape[1])
            else:
                self.unet.prime_conditioning(shape, cond, self.use_cuda_graph)
        # CPU draw, then moved (custom_pipelines.py:117-123); with CPU generator(s) the values do not depend on the target
        # device, so the per-sample draws go straight into a pinned buffer and across asynchronously
        image = randn_tensor(shape, generator=generator, device=dev if generator is not None else None)
        image = image.to(dev, torch.float32).contiguous()
        image, stack = self._loop(image, cond, srcs, build_cond, num_inference_steps, plot_every, add_noise_rgb,
                                  raw_depth, denorm)
        return GuidedPipelineOutput(images=image, images_pred_orig=stack, images_perturbed_orig=stack,
                                    images_pred_prev=stack, images_purturbed_pred_prev=stack,
                                    images_sampled_prev=stack)

"""``GuidedLatentDiffusionPipeline`` — the sampling LOOP of the reference's latent family
(``/root/reference/core/custom_pipelines.py:208-460``) with the same constructor, ``__call__`` signature, conditioning
keys, RNG consumption and output semantics, over this package's scheduler kernels.

Networks: ``unet`` is ``d3roma_b200.UNet2DConditionModel`` and ``vae`` ``d3roma_b200.AutoencoderKL``
(the SD-2.1 UNet / SD VAE architectures on the sm_100a kernels, ``unet_condition.py`` / ``autoencoder_kl.py``); any object
with the same call surface works too (the loop only uses ``unet(x, t, ctx).sample`` and ``vae.encoder / quant_conv /
post_quant_conv / decoder``, exactly what the reference touches, ``train.py:743-772, 821-833``).  ``text_encoder`` is
``d3roma_b200.CLIPTextModel`` (``text_encoder.py``; fp32 kernels) or the caller's transformers object: its only use on
this path is the embedding of the empty prompt, a constant ``[1, 2, 1024]`` tensor, which callers may also pass
directly as ``empty_text_embed`` (the argument the reference provides for exactly that, custom_pipelines.py:229,
299-301).  With ``tokenizer=None`` the package's encoder supplies the two token ids of ``""`` itself (the BPE tokenizer,
host-side vocabulary files, is not rebuilt).  This file is the reference's own part of
the path: the loop, its latent encode / decode conventions (``encode_rgb`` :47-59, ``encode_disp`` :35-45,
``__decode_depth`` :255-272), and the scheduler steps (``d3_sched_ddim_step`` / ``d3_sched_ddpm_step``) with the guidance
hook wiring of :389-399.

Differences from the reference, none of which change a returned value:
  * ``encode_disp(sim_disp)`` is evaluated once instead of at every step of the ``"rgb+raw"`` variant (:356 re-encodes
    a constant with a deterministic encoder);
  * the final ``__decode_depth(image)`` of :428 is skipped (its result is discarded: the pipeline returns
    ``image_final``, the decoded x0 of the last step, :414-415, 454);
  * the five intermediate lists (:403-411) hold one decoded tensor per plotted step, aliased five times.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Union

import torch

from .pipelines import GuidedPipelineOutput
from .scheduler_ddim import randn_tensor

LATENT_SCALE = 0.18215


def encode_rgb(vae, x: torch.Tensor, rgb_latent_scale_factor: float = LATENT_SCALE) -> torch.Tensor:
    """custom_pipelines.py:47-59: mean of the VAE posterior, scaled."""
    moments = vae.quant_conv(vae.encoder(x))
    mean, _ = torch.chunk(moments, 2, dim=1)
    return mean * rgb_latent_scale_factor


def encode_disp(vae, x: torch.Tensor, depth_latent_scale_factor: float = LATENT_SCALE) -> torch.Tensor:
    """custom_pipelines.py:35-45: the one-channel map replicated to three channels, then ``encode_rgb``."""
    return encode_rgb(vae, x.repeat(1, 3, 1, 1), depth_latent_scale_factor)


class LatentGradientGuidance:
    """The latent-space form of ``FlowGuidance.perturb`` in "gradient" mode (core/guidance.py:225-241): ten Adam steps
    (lr 1e-3) on the predicted latent, minimising the masked L1 distance between the decoded + denormalised map and the
    raw depth.  It differentiates THROUGH the decoder, so it needs a decoder torch.autograd can see (the caller's torch
    ``AutoencoderKL``); with the forward-only CUDA decoder of this package it raises.  Every LDM config the reference
    ships evaluates with ``flow_guidance_weights: [0]`` (conf/task/eval_ldm*.yaml), i.e. never reaches this."""

    def __init__(self, guidance):
        self.flow_guidance_weight = guidance.flow_guidance_weight
        self.flow_guidance_mode = guidance.flow_guidance_mode

    def perturb(self, pred_original_sample, t, decoder, denormer, raw_mask=None, left_image=None, right_image=None,
                raw_depth=None):
        if self.flow_guidance_mode != "gradient":
            raise NotImplementedError
        lat = pred_original_sample.clone().requires_grad_(True)
        raw_depth = raw_depth.to(lat.device)
        mask = raw_depth > 0.0
        opt = torch.optim.Adam([lat], lr=0.001)
        with torch.enable_grad():
            for _ in range(10):
                opt.zero_grad()
                pred = denormer(decoder(lat))
                if not pred.requires_grad:
                    raise NotImplementedError("latent gradient guidance needs a differentiable (torch) VAE decoder")
                torch.nn.functional.l1_loss(pred * mask, raw_depth * mask).backward()
                opt.step()
        return lat.detach()


class GuidedLatentDiffusionPipeline:
    def __init__(self, unet, vae, tokenizer, text_encoder, scheduler, guidance):
        self.register_modules(unet=unet, vae=vae, tokenizer=tokenizer, text_encoder=text_encoder, scheduler=scheduler,
                              guidance=guidance)
        self._progress_bar_config = {}
        self.empty_text_embed = None   # optional cached [1, 2, D] embedding of the empty prompt (skips the text encoder)

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self) -> torch.device:
        for m in (self.unet, self.vae):
            if isinstance(m, torch.nn.Module):
                for p in m.parameters():
                    return p.device
            elif isinstance(getattr(m, "device", None), torch.device) and m.device.type == "cuda":
                return m.device
        return torch.device("cuda", torch.cuda.current_device())

    def to(self, device):
        for m in (self.unet, self.vae, self.text_encoder):
            if m is not None and hasattr(m, "to"):
                m.to(device)
        return self

    # ---- diffusers directory layout (train.py:682, 701 `pipeline.save_pretrained`; evaluate.py:313-314 / inference.py:72-73
    # `GuidedLatentDiffusionPipeline.from_pretrained(path).to("cuda")`)
    def save_pretrained(self, path):
        """model_index.json + unet/ vae/ text_encoder/ scheduler/ guidance/ (tokenizer/ when it can save itself).  The
        networks must have been loaded with ``keep_state_dict=True`` (they keep their weights in kernel layouts)."""
        import json
        import os
        from ._config import save_config
        os.makedirs(path, exist_ok=True)
        index = {"_class_name": "GuidedLatentDiffusionPipeline"}
        for name in ("unet", "vae", "text_encoder", "tokenizer"):
            m = getattr(self, name)
            if m is None or not hasattr(m, "save_pretrained"):
                index[name] = [None, None]
                continue
            m.save_pretrained(os.path.join(path, name))
            index[name] = [type(m).__module__.split(".")[0], type(m).__name__]
        save_config(self.scheduler, os.path.join(path, "scheduler"), "scheduler_config.json")
        index["scheduler"] = ["d3roma_b200", type(self.scheduler).__name__]
        g = self.guidance
        os.makedirs(os.path.join(path, "guidance"), exist_ok=True)
        with open(os.path.join(path, "guidance", "config.json"), "w") as f:
            json.dump({"_class_name": "FlowGuidance", "flow_guidance_weight": g.flow_guidance_weight,
                       "perturb_start_ratio": getattr(g, "perturb_start_ratio", 0.0),
                       "flow_guidance_mode": g.flow_guidance_mode}, f)
        index["guidance"] = ["d3roma_b200", "FlowGuidance"]
        with open(os.path.join(path, "model_index.json"), "w") as f:
            json.dump(index, f, indent=2)

    @classmethod
    def from_pretrained(cls, path, compute_dtype=None, device="cuda", keep_state_dict=False, tokenizer=None):
        """A pipeline directory as diffusers writes it for the reference (``unet/``, ``vae/``, ``text_encoder/``,
        ``tokenizer/``, ``scheduler/``, ``guidance/``): the three networks on this package's kernels, the reference's
        scheduler class named by ``scheduler_config.json``.  The tokenizer is the caller's (``tokenizer=``), or
        ``transformers.CLIPTokenizer`` when that library and a ``tokenizer/`` folder are there, else ``None`` (the text
        encoder then encodes the empty prompt from its two token ids).  A directory without ``text_encoder/`` gives a
        pipeline that needs ``empty_text_embed``."""
        import inspect
        import os
        from ._config import load_config
        from .autoencoder_kl import AutoencoderKL
        from .guidance import FlowGuidance
        from .scheduler_ddim import MyDDIMScheduler
        from .scheduler_ddpm import MyDDPMScheduler
        from .text_encoder import CLIPTextModel
        from .unet_condition import UNet2DConditionModel
        kw = dict(compute_dtype=compute_dtype, device=device, keep_state_dict=keep_state_dict)
        unet = UNet2DConditionModel.from_pretrained(os.path.join(path, "unet"), **kw)
        vae = AutoencoderKL.from_pretrained(os.path.join(path, "vae"), **kw)
        text_encoder = None
        if os.path.exists(os.path.join(path, "text_encoder", "config.json")):
            text_encoder = CLIPTextModel.from_pretrained(os.path.join(path, "text_encoder"), device=device)
        if tokenizer is None and os.path.isdir(os.path.join(path, "tokenizer")):
            try:
                from transformers import CLIPTokenizer
                tokenizer = CLIPTokenizer.from_pretrained(os.path.join(path, "tokenizer"))
            except Exception:       # library or vocabulary files absent: the encoder knows the ids of ""
                tokenizer = None
        scfg = load_config(os.path.join(path, "scheduler"), "scheduler_config.json")
        sched_cls = MyDDPMScheduler if "DDPM" in scfg.get("_class_name", "DDPM") else MyDDIMScheduler
        gpath = os.path.join(path, "guidance", "config.json")
        gcfg = load_config(os.path.join(path, "guidance"), "config.json") if os.path.exists(gpath) else {}
        known = set(inspect.signature(FlowGuidance.__init__).parameters) - {"self", "camera"}
        guidance = FlowGuidance(**{k: v for k, v in gcfg.items() if k in known})
        return cls(unet, vae, tokenizer, text_encoder, sched_cls.from_config(scfg), guidance)

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

    def _encode_empty_text(self) -> torch.Tensor:
        """custom_pipelines.py:241-253: embedding of the empty prompt ([1, 2, 1024] for the SD-2.1 text encoder)."""
        if self.tokenizer is None:
            # d3roma_b200.CLIPTextModel knows the two token ids of "" (BOS, EOS); any other encoder needs its tokenizer
            if not hasattr(self.text_encoder, "encode_empty_prompt"):
                raise ValueError("GuidedLatentDiffusionPipeline: pass empty_text_embed, or a tokenizer with the text encoder")
            return self.text_encoder.encode_empty_prompt()
        tok = self.tokenizer("", padding="do_not_pad", max_length=self.tokenizer.model_max_length, truncation=True,
                             return_tensors="pt")
        return self.text_encoder(tok.input_ids.to(self.text_encoder.device))[0]

    def _decode_depth(self, depth_latent: torch.Tensor, depth_latent_scale_factor: float = LATENT_SCALE) -> torch.Tensor:
        """custom_pipelines.py:255-272: undo the latent scale, decode, mean over the output channels.  With this
        package's ``AutoencoderKL`` the three steps are one pass (``decode_mean``: scale and channel mean folded into
        the 1x1 / output convs)."""
        if hasattr(self.vae, "decode_mean"):
            return self.vae.decode_mean(depth_latent, depth_latent_scale_factor)
        z = self.vae.post_quant_conv(depth_latent / depth_latent_scale_factor)
        return self.vae.decoder(z).mean(dim=1, keepdim=True)

    def _encode_all(self, rgb, left, right, sim):
        """Latents of every conditioning image the call was given (custom_pipelines.py:312-325).  With this package's
        ``AutoencoderKL`` the images go through the encoder as ONE batch (the reference runs up to four encoder passes)
        and the mean / scale selection is folded into the quant conv (``encode_latent``)."""
        if not hasattr(self.vae, "encode_latent"):
            return (None if rgb is None else encode_rgb(self.vae, rgb), None if left is None else encode_rgb(self.vae, left),
                    None if right is None else encode_rgb(self.vae, right), None if sim is None else encode_disp(self.vae, sim))
        imgs = [rgb, left, right, None if sim is None else sim.repeat(1, 3, 1, 1)]
        present = [t for t in imgs if t is not None]
        lat = self.vae.encode_latent(torch.cat(present, dim=0).float().contiguous(), LATENT_SCALE)
        parts = iter(torch.split(lat, [t.shape[0] for t in present], dim=0))
        return tuple(None if t is None else next(parts).contiguous() for t in imgs)

    @torch.no_grad()
    def __call__(self, rgb_images, left_images=None, right_images=None, sim_disp=None, raw_depth=None, raw_mask=None,
                 empty_text_embed=None, num_inference_steps: int = 128, num_intermediate_images: int = 128,
                 add_noise_rgb: bool = False, depth_channels: int = 1, cond_channels: str = "rgb",
                 denorm: Callable = None,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None):
        """ in & out images are preferably on the same device """
        assert num_intermediate_images > 0 and num_inference_steps % num_intermediate_images == 0
        plot_every = num_inference_steps // num_intermediate_images
        shapes = [t.shape for t in (rgb_images, left_images, right_images, sim_disp) if t is not None]
        assert shapes, "no input image"
        input_shape = shapes[0]
        dev = self.device
        if empty_text_embed is None:        # callers that cannot pass it (evaluate.py:95 eval_batch) may set the attribute
            empty_text_embed = getattr(self, "empty_text_embed", None)
        if empty_text_embed is None:
            empty_text_embed = self._encode_empty_text()
        empty_text_embed = empty_text_embed.to(dev).repeat((input_shape[0], 1, 1))

        image = randn_tensor((input_shape[0], 4, input_shape[2] // 8, input_shape[3] // 8), generator=generator)
        image = image.to(dev)
        up = lambda t: None if t is None else t.to(dev)  # noqa: E731
        rgb_images, left_images, right_images, sim_disp = up(rgb_images), up(left_images), up(right_images), up(sim_disp)
        cond_images, left_lat, right_lat, sim_lat = self._encode_all(rgb_images, left_images, right_images, sim_disp)

        orders = {"rgb": ("rgb",), "rgb+raw": ("rgb", "raw"), "left+right": ("left", "right"),
                  "left+right+raw": ("left", "right", "raw"), "rgb+left+right": ("rgb", "left", "right"),
                  "rgb+left+right+raw": ("rgb", "left", "right", "raw")}
        if cond_channels not in orders:
            raise ValueError(f"Unknown cond_channels: {cond_channels}")
        lat = {"left": left_lat, "right": right_lat, "raw": sim_lat}

        self.scheduler.set_timesteps(num_inference_steps)
        is_ddim = self.scheduler.__class__.__name__ == "MyDDIMScheduler"
        guided = self.guidance is not None and self.guidance.flow_guidance_weight > 0.0
        guidance = LatentGradientGuidance(self.guidance) if guided else None
        decoded: List[torch.Tensor] = []
        image_final = None
        for step, t in enumerate(self.progress_bar(self.scheduler.timesteps)):
            final_rgb = cond_images
            if "rgb" in cond_channels and add_noise_rgb:
                noise_rgb = torch.randn(cond_images.shape).to(cond_images.device)
                timesteps = (t * torch.ones(cond_images.shape[0]).to(cond_images.device)).long()
                noisy_rgb = self.scheduler.add_noise(cond_images, noise_rgb, timesteps)
                final_rgb = cond_images * 0.5 + noisy_rgb * 0.5
            parts = [image] + [final_rgb if k == "rgb" else lat[k] for k in orders[cond_channels]]
            model_output = self.unet(torch.cat(parts, dim=1), t, empty_text_embed).sample
            if is_ddim:
                assert self.guidance.flow_guidance_mode != "imputation", \
                    "latent diffusion does not support imputation mode"
                g = self.scheduler.step(model_output, t, image, guidance=guidance, decoder=self._decode_depth,
                                        denormer=denorm, raw_mask=raw_mask, left_image=left_images,
                                        right_image=right_images, raw_depth=raw_depth)
            else:
                g = self.scheduler.step(model_output, t, image)
            image = g.prev_sample
            last = step == num_inference_steps - 1
            d_hat = None
            if (step + 1) % plot_every == 0:
                d_hat = self._decode_depth(g.pred_original_sample).clamp(-1, 1)
                decoded.append(d_hat)
            if last:
                image_final = d_hat if d_hat is not None else self._decode_depth(g.pred_original_sample).clamp(-1, 1)
        stack = torch.concat(decoded, dim=1)
        return GuidedPipelineOutput(images=image_final, images_pred_orig=stack, images_perturbed_orig=stack,
                                    images_pred_prev=stack, images_purturbed_pred_prev=stack,
                                    images_sampled_prev=stack)

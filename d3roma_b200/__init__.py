"""d3roma_b200 — B200-native (sm_100a) implementation of D3RoMa's iterative disparity-denoising loop.

Public surface mirrors the reference's classes for this path (jie0530/d3roma ``core/``):
``GuidedDiffusionPipeline``, ``MyDDIMScheduler``, ``MyDDPMScheduler``, ``FlowGuidance``, ``UNet2DModel``,
``Normalizer``, ``InputPadder``, ``D3RoMa`` (single-frame inference front end); the latent family adds
``GuidedLatentDiffusionPipeline``, ``UNet2DConditionModel``, ``AutoencoderKL`` and ``CLIPTextModel``.  All arithmetic runs in ``libd3roma_b200.so`` (build: ``python -m d3roma_b200.build``).
"""
from .guidance import FlowGuidance  # noqa: F401
from .pipelines import GuidedDiffusionPipeline, GuidedPipelineOutput  # noqa: F401
from .latent_pipeline import GuidedLatentDiffusionPipeline  # noqa: F401
from .prepost import SIM_CAMERA, Camera, InputPadder, Normalizer  # noqa: F401
from .scheduler_ddim import DDIMSchedulerOutput, MyDDIMScheduler  # noqa: F401
from .scheduler_ddpm import DDPMSchedulerOutput, MyDDPMScheduler  # noqa: F401
from .unet import UNet2DModel  # noqa: F401
from .unet_condition import UNet2DConditionModel  # noqa: F401
from .autoencoder_kl import AutoencoderKL  # noqa: F401
from .text_encoder import CLIPTextModel  # noqa: F401
from .metrics import compute_errors, compute_scale_and_shift, metrics_to_dict  # noqa: F401
from .inference import D3RoMa  # noqa: F401

__version__ = "0.1.0"

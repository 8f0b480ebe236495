"""ORACLE tooling (test infrastructure only): a minimal stand-in for the ``diffusers`` API surface that
the reference's own ``core/*.py`` import, so that those files can be imported UNMODIFIED from
``/root/reference`` in this container (where ``diffusers``, ``open3d`` and ``matplotlib`` are not
installed) to generate golden vectors.  Nothing here is used at run time by the product, the GPU
tests, ``smoke()`` or ``bench.py`` — only by ``oracle/gen_golden.py``.

Stubbed surface: ``ConfigMixin.from_config/.config``, ``register_to_config``,
``SchedulerMixin``, ``KarrasDiffusionSchedulers``, ``BaseOutput``, ``randn_tensor``,
``DiffusionPipeline.register_modules/progress_bar/device``, ``ModelMixin``,
``pipelines.ImagePipelineOutput``; inert ``open3d`` / ``matplotlib(.pyplot)`` modules.
"""
from __future__ import annotations

import enum
import functools
import inspect
import sys
import types
from collections import OrderedDict
from dataclasses import fields, is_dataclass

import torch


class _FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = list(sig.parameters.values())[1:]
        cfg = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update({k: v for k, v in kwargs.items() if k in sig.parameters})
        self._internal_dict = _FrozenDict(cfg)
        init(self, *args, **{k: v for k, v in kwargs.items() if k in sig.parameters})

    return inner


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config)
        cfg.update(kwargs)
        sig = inspect.signature(cls.__init__)
        return cls(**{k: v for k, v in cfg.items() if k in sig.parameters})


class SchedulerMixin:
    pass


class KarrasDiffusionSchedulers(enum.Enum):
    DDIMScheduler = 1
    DDPMScheduler = 2


class BaseOutput(OrderedDict):
    """dataclass-style output with attribute access and mutable fields (what the callers rely on)."""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    OrderedDict.__setitem__(self, f.name, v)

    def __setattr__(self, k, v):
        object.__setattr__(self, k, v)
        if is_dataclass(self) and k in {f.name for f in fields(self)} and v is not None:
            OrderedDict.__setitem__(self, k, v)


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on the generator's device (CPU by default), then move."""
    device = device or torch.device("cpu")
    rand_device = device
    if generator is not None:
        g0 = generator[0] if isinstance(generator, (list, tuple)) else generator
        if g0.device.type != torch.device(device).type and g0.device.type == "cpu":
            rand_device = "cpu"
    if isinstance(generator, (list, tuple)):
        s = (1,) + tuple(shape[1:])
        lat = torch.cat([torch.randn(s, generator=g, device=rand_device, dtype=dtype) for g in generator], 0)
        return lat.to(device)
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)


class ModelMixin(torch.nn.Module):
    pass


class DiffusionPipeline:
    def __init__(self):
        self._pb = {}

    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def device(self):
        for v in vars(self).values():
            if isinstance(v, torch.nn.Module):
                for p in v.parameters():
                    return p.device
        return torch.device("cpu")

    def to(self, device):
        for v in vars(self).values():
            if isinstance(v, torch.nn.Module):
                v.to(device)
        return self

    def set_progress_bar_config(self, **kw):
        self._pb = kw

    def progress_bar(self, it):
        return it


class ImagePipelineOutput(BaseOutput):
    pass


def install() -> None:
    """Seed ``sys.modules`` so that ``import diffusers...`` inside the reference resolves to this stub."""
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "__d3_stub__", False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    d = mod("diffusers", __d3_stub__=True, DiffusionPipeline=DiffusionPipeline, ModelMixin=ModelMixin,
            ConfigMixin=ConfigMixin)
    d.configuration_utils = mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin,
                                register_to_config=register_to_config)
    d.utils = mod("diffusers.utils", BaseOutput=BaseOutput)
    d.utils.torch_utils = mod("diffusers.utils.torch_utils", randn_tensor=randn_tensor)
    d.schedulers = mod("diffusers.schedulers", KarrasDiffusionSchedulers=KarrasDiffusionSchedulers,
                       SchedulerMixin=SchedulerMixin)
    d.schedulers.scheduling_utils = mod("diffusers.schedulers.scheduling_utils",
                                        KarrasDiffusionSchedulers=KarrasDiffusionSchedulers,
                                        SchedulerMixin=SchedulerMixin)
    d.models = mod("diffusers.models", ModelMixin=ModelMixin)
    d.pipelines = mod("diffusers.pipelines", ImagePipelineOutput=ImagePipelineOutput)

    class _Inert(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Inert(self.__name__ + "." + k)

        def __call__(self, *a, **k):
            return None

    for name in ("open3d", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = _Inert(name)

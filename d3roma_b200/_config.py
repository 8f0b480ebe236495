"""Tiny stand-ins for the ``diffusers`` plumbing the reference's classes inherit from (``ConfigMixin`` /
``register_to_config`` / ``BaseOutput``): attribute-style ``.config``, ``from_config(dict)`` that ignores unknown
keys (the reference relies on that: ``skip_prk_steps`` in config.py:206), and mutable dataclass outputs."""
from __future__ import annotations

import functools
import inspect
from collections import OrderedDict
from dataclasses import fields, is_dataclass


class FrozenConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = list(sig.parameters.values())[1:]
        cfg = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
        cfg.update({p.name: a for p, a in zip(params, args)})
        known = {k: v for k, v in kwargs.items() if k in sig.parameters}
        cfg.update(known)
        self._internal_dict = FrozenConfig(cfg)
        init(self, *args, **known)

    return inner


class ConfigMixin:
    @property
    def config(self) -> FrozenConfig:
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config)
        cfg.update(kwargs)
        sig = inspect.signature(cls.__init__)
        return cls(**{k: v for k, v in cfg.items() if k in sig.parameters})


def save_config(obj, path, filename):
    """Write ``obj.config`` (+ ``_class_name``) as JSON, diffusers style."""
    import json
    import os
    os.makedirs(path, exist_ok=True)
    cfg = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in dict(obj.config).items()}
    cfg["_class_name"] = type(obj).__name__
    with open(os.path.join(path, filename), "w") as f:
        json.dump(cfg, f, indent=2)


def load_config(path, filename):
    import json
    import os
    with open(os.path.join(path, filename)) as f:
        return json.load(f)


class BaseOutput(OrderedDict):
    """Dataclass-style output: attribute access, tuple-style indexing, reassignable fields."""

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

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class PretrainedMixin:
    """``save_pretrained`` / ``from_pretrained`` in the diffusers directory layout (``config.json`` +
    ``diffusion_pytorch_model.safetensors`` or ``.bin``) for the network classes whose ``load_state_dict`` takes diffusers
    key names.  The networks keep their weights in kernel layouts, so writing a checkpoint back needs the fp32
    ``state_dict`` they were loaded from: set ``keep_state_dict = True`` before loading (``from_pretrained(...,
    keep_state_dict=True)``) when ``save_pretrained`` / ``state_dict`` will be called."""

    keep_state_dict = False
    _sd_cpu = None
    _weights_basename = "diffusion_pytorch_model"

    def _remember_state_dict(self, sd):
        self._sd_cpu = {k: v.detach().to("cpu", copy=True) for k, v in sd.items()} if self.keep_state_dict else None

    def state_dict(self):
        if self._sd_cpu is None:
            raise RuntimeError(f"{type(self).__name__}.state_dict() needs keep_state_dict=True at load time")
        return dict(self._sd_cpu)

    def save_pretrained(self, path):
        import json
        import os
        from safetensors.torch import save_file
        sd = self.state_dict()
        os.makedirs(path, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(self.config).items()}
        cfg.update(_class_name=type(self).__name__, compute_dtype=self.compute_dtype)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        save_file({k: v.contiguous() for k, v in sd.items()}, os.path.join(path, self._weights_basename + ".safetensors"))

    @classmethod
    def from_pretrained(cls, path, subfolder=None, compute_dtype=None, device="cuda", keep_state_dict=False, **overrides):
        """``Class.from_pretrained("checkpoint/stable-diffusion/unet")`` / ``(..., subfolder="vae")`` (train.py:750, 830):
        keys of ``config.json`` the class does not model are validated or ignored by its constructor; bookkeeping keys
        (``_class_name``, ``_diffusers_version``, ``_name_or_path``) are dropped."""
        import json
        import os
        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        cfg.update(overrides)
        if compute_dtype is not None:
            cfg["compute_dtype"] = compute_dtype
        net = cls(**cfg)
        if device is not None:
            net.to(device)
        net.keep_state_dict = keep_state_dict
        return net.load_state_dict(load_weights(path, cls._weights_basename))


def load_weights(path, basename):
    """``<basename>.safetensors`` when present, else ``<basename>.bin`` (torch pickle, tensors only)."""
    import os
    st = os.path.join(path, basename + ".safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    import torch
    return torch.load(os.path.join(path, basename + ".bin"), map_location="cpu", weights_only=True)

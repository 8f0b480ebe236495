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

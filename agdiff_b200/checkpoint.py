"""Checkpoint / config front door of the reference's scripts.

``scripts/train.py:218-229`` saves ``{"config": EasyDict, "model": state_dict, optimizers, schedulers, "iteration",
"avg_val_loss"}`` with ``torch.save``; ``scripts/test.py:78-82,111-114`` reads it back with ``torch.load`` and builds
``get_model(ckpt["config"].model)``.  The pickle references ``easydict.EasyDict`` (a ``dict`` subclass with attribute access),
which is not a dependency of this package: ``load_checkpoint`` supplies a stand-in with the same behaviour while unpickling,
so a reference checkpoint loads whether or not ``easydict`` is installed.  ``load_config`` reads ``configs/*.yml`` the way the
scripts do.
"""
from __future__ import annotations

import sys
import types
from typing import Any, Dict, Tuple

import torch


class AttrDict(dict):
    """dict with attribute access, nested dicts converted on the way in (the behaviour of ``easydict.EasyDict`` the reference
    relies on: ``config.model.hidden_dim``, ``config.train.seed``; missing keys raise AttributeError so that ``copy`` / ``pickle``
    protocols probing for dunder attributes keep working)"""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setstate__(self, state):          # EasyDict pickles carry their items both as dict items and in __dict__
        for k, v in (state or {}).items():
            self[k] = v


def load_config(path: str) -> AttrDict:
    """``EasyDict(yaml.safe_load(f))`` of scripts/test.py:81-82 / scripts/train.py:46-47"""
    import yaml
    with open(path, "r") as f:
        return AttrDict(yaml.safe_load(f))


def load_checkpoint(path: str, map_location="cpu") -> Dict[str, Any]:
    """``torch.load`` of a reference checkpoint; ``ckpt["config"]`` comes back as an ``AttrDict`` when ``easydict`` is absent"""
    injected = False
    if "easydict" not in sys.modules:
        try:
            import easydict  # noqa: F401
        except ImportError:
            stub = types.ModuleType("easydict")
            stub.EasyDict = AttrDict
            sys.modules["easydict"] = stub
            injected = True
    try:
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
    finally:
        if injected:
            del sys.modules["easydict"]
    return ckpt


def model_from_checkpoint(path: str, device="cuda:0") -> Tuple[Any, Dict[str, Any]]:
    """scripts/test.py:78,111-114 in one call: load, ``get_model(ckpt["config"].model)``, ``load_state_dict``, ``eval``"""
    from .epsnet import get_model
    ckpt = load_checkpoint(path, "cpu")
    model = get_model(ckpt["config"].model)
    model.load_state_dict(ckpt["model"])
    return model.eval().to(device), ckpt

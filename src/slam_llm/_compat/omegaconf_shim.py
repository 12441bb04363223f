"""Minimal stand-in for the part of OmegaConf the SLAM-LLM recipes use.

Installed into sys.modules as `omegaconf` ONLY when the real package is not importable (this image has
no network).  Covers: DictConfig / ListConfig, OmegaConf.{create, structured, merge, to_container,
set_struct, to_yaml}, attribute + item access, .get, `del cfg[k]`, ** unpacking.
"""
from __future__ import annotations

import copy
import dataclasses
import enum
from typing import Any

import yaml


class ListConfig(list):
    pass


class DictConfig(dict):
    """dict with attribute access.  Unknown attributes raise AttributeError (like OmegaConf's struct mode
    only when the node was built from a dataclass AND struct is on; the shim is lenient and allows new keys)."""

    def __init__(self, content=None, **kw):
        super().__init__()
        object.__setattr__(self, "_struct", False)
        if content:
            for k, v in dict(content).items():
                self[k] = v
        for k, v in kw.items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, _wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(f"Missing key {k}") from None

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def get(self, k, default=None):
        v = super().get(k, default)
        return default if v is None and k not in self else v

    def copy(self):
        return DictConfig(copy.deepcopy(dict(self)))

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, (DictConfig, ListConfig)):
        return v
    if dataclasses.is_dataclass(v):
        return _from_dataclass(v)
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(_wrap(x) for x in v)
    if isinstance(v, enum.Enum):
        return v.name
    return v


def _from_dataclass(obj) -> DictConfig:
    if isinstance(obj, type):
        obj = obj()
    out = DictConfig()
    for f in dataclasses.fields(obj):
        out[f.name] = getattr(obj, f.name)
    return out


def _merge_into(dst: DictConfig, src) -> DictConfig:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_into(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def create(obj=None) -> Any:
        if obj is None:
            return DictConfig()
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        return _wrap(obj)

    @staticmethod
    def structured(obj) -> DictConfig:
        return _from_dataclass(obj)

    @staticmethod
    def merge(*cfgs) -> DictConfig:
        out = DictConfig()
        for c in cfgs:
            if c is None:
                continue
            _merge_into(out, _wrap(c))
        return out

    @staticmethod
    def to_container(cfg, resolve: bool = True, **_):
        return _plain(cfg)

    @staticmethod
    def set_struct(cfg, value: bool) -> None:
        if isinstance(cfg, DictConfig):
            object.__setattr__(cfg, "_struct", bool(value))

    @staticmethod
    def to_yaml(cfg) -> str:
        return yaml.safe_dump(_plain(cfg), sort_keys=False)

    @staticmethod
    def load(path) -> DictConfig:
        with open(path, encoding="utf-8") as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def from_dotlist(items) -> DictConfig:
        out = DictConfig()
        for it in items:
            key, _, val = it.partition("=")
            set_by_path(out, key.lstrip("+"), parse_value(val))
        return out


def parse_value(text: str):
    """Hydra-style scalar / list parsing of a CLI override value."""
    t = text.strip()
    if (t.startswith('"') and t.endswith('"')) or (t.startswith("'") and t.endswith("'")):
        return t[1:-1]
    if t.lower() in ("null", "none", "~"):
        return None
    try:
        v = yaml.safe_load(t)
    except yaml.YAMLError:
        return t
    if isinstance(v, (dict,)):
        return t  # paths with ':' etc. stay strings
    if isinstance(v, str):
        try:
            return float(v) if any(c in v for c in ".eE") and not v.startswith("/") else v   # YAML 1.1 misses "5e-5"
        except ValueError:
            return v
    return v


def set_by_path(cfg: DictConfig, dotted: str, value) -> None:
    parts = dotted.split(".")
    node = cfg
    for p in parts[:-1]:
        if p not in node or not isinstance(node[p], dict):
            node[p] = DictConfig()
        node = node[p]
    node[parts[-1]] = value

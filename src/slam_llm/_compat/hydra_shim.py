"""Minimal stand-in for `hydra.main` as the SLAM-LLM recipes use it:
`@hydra.main(config_name=None, version_base=None)` + CLI `--config-path P --config-name N`, overrides
`a.b=c`, `+a.b=c`, `++a.b=c`; `hydra.*` keys (e.g. hydra.run.dir=...) are accepted and ignored except that
run.dir is created.  Installed as `hydra` ONLY when the real package is not importable."""
from __future__ import annotations

import functools
import inspect
import os
import sys

from .omegaconf_shim import DictConfig, OmegaConf, parse_value, set_by_path


def _parse_argv(argv):
    cfg_path, cfg_name, overrides = None, None, []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("--config-path", "-cp"):
            cfg_path = argv[i + 1]; i += 2; continue
        if a.startswith("--config-path="):
            cfg_path = a.split("=", 1)[1]; i += 1; continue
        if a in ("--config-name", "-cn"):
            cfg_name = argv[i + 1]; i += 2; continue
        if a.startswith("--config-name="):
            cfg_name = a.split("=", 1)[1]; i += 1; continue
        if "=" in a and not a.startswith("--"):
            overrides.append(a)
        i += 1
    return cfg_path, cfg_name, overrides


def main(config_path=None, config_name=None, version_base=None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(cfg_passthrough=None):
            if cfg_passthrough is not None:
                return fn(cfg_passthrough)
            cli_path, cli_name, overrides = _parse_argv(sys.argv[1:])
            path, name = cli_path or config_path, cli_name or config_name
            cfg = DictConfig()
            if name is not None:
                base = os.path.dirname(os.path.abspath(inspect.getsourcefile(fn)))
                full = os.path.join(path if path and os.path.isabs(path) else os.path.join(base, path or ""), name)
                if not os.path.exists(full) and not full.endswith((".yaml", ".yml")):
                    full += ".yaml"
                cfg = OmegaConf.load(full)
            for ov in overrides:
                key, _, val = ov.partition("=")
                key = key.lstrip("+")
                if key.startswith("hydra."):
                    if key == "hydra.run.dir":
                        os.makedirs(val, exist_ok=True)
                    continue
                set_by_path(cfg, key, parse_value(val))
            return fn(cfg)
        return wrapper
    return deco

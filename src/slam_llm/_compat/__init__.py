"""Offline compatibility shims: installed into sys.modules only when the real package is missing."""
from __future__ import annotations

import importlib
import importlib.util
import sys


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install() -> None:
    if _missing("omegaconf"):
        from . import omegaconf_shim
        sys.modules["omegaconf"] = omegaconf_shim
    if _missing("hydra"):
        from . import hydra_shim
        sys.modules["hydra"] = hydra_shim
    if _missing("whisper"):
        from . import whisper_shim
        sys.modules["whisper"] = whisper_shim

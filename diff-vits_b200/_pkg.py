"""Public API of diffvits_b200 (see DESIGN.md / INTEGRATION.md)."""
from __future__ import annotations

import os
import sys

from . import _lib
from ._lib import DVError, LIB_PATH
from .schedule import (NoiseScheduleVP, dpmpp_multistep_table, get_time_steps, interpolate_fn, model_wrapper,
                       unipc_multistep_table)

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")

__all__ = ["DVError", "LIB_PATH", "DROPIN_DIR", "install_dropin", "build", "NoiseScheduleVP", "model_wrapper",
           "interpolate_fn", "get_time_steps", "dpmpp_multistep_table", "unipc_multistep_table", "lazy"]
LAZY = ("NativeDenoiser", "sample_native", "DenoiserEngine")


def install_dropin() -> str:
    """Put the drop-in `unet1d` / `sampler` packages ahead of the reference's on sys.path. Call before importing the
    reference's `model3` / `tts_infer` (they resolve the classes by module path, model3.py:40-41,1128,1162)."""
    for name in ("unet1d", "sampler"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(DROPIN_DIR):
            raise RuntimeError(f"module {name!r} was already imported from {getattr(mod, '__file__', None)}; "
                               "call install_dropin() before importing the reference")
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR


def build(force: bool = False, verbose: bool = False) -> str:
    from .build import build as _build
    return _build(force=force, verbose=verbose)


def lazy(name):
    """Resolve the heavier public names on first use (NativeDenoiser, sample_native, DenoiserEngine)."""
    if name in ("NativeDenoiser", "sample_native"):
        import importlib
        return getattr(importlib.import_module("diffvits_b200.native"), name)
    if name == "DenoiserEngine":
        import importlib
        return importlib.import_module("diffvits_b200.engine").DenoiserEngine
    raise AttributeError(f"module 'diffvits_b200' has no attribute {name!r}")

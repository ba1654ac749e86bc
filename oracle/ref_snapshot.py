"""ORACLE helper (test / measurement infrastructure, never imported by the product path): run the UNMODIFIED reference
from `oracle/_ref/` (snapshot made by oracle/make_ref_snapshot.py; falls back to /root/reference in the build container).

The reference resolves its classes by module name (`unet1d`, `sampler`, `model3`, ...), the same names the drop-in
packages use, so ONE process hosts ONE arm: call `activate("reference")` or `activate("dropin")` before anything imports
those names, then `import_model3()`. Comparisons between the arms run them in separate processes (`run_arm`).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import json
import os
import subprocess
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNAPSHOT = os.path.join(ROOT, "oracle", "_ref")
CHECKOUT = os.environ.get("DIFFVITS_REFERENCE", "/root/reference")
# third-party packages model3.py imports at module level but never touches on the sampling path
STUBBED = ("vocos", "accelerate", "matplotlib", "librosa", "ema_pytorch")


def ref_root() -> str:
    """Directory holding the reference files: the snapshot if present (it is what travels to the GPU box), else the checkout."""
    if os.path.isfile(os.path.join(SNAPSHOT, "model3.py")):
        return SNAPSHOT
    if os.path.isfile(os.path.join(CHECKOUT, "model3.py")):
        return CHECKOUT
    raise FileNotFoundError("no reference available: run `python -m oracle.make_ref_snapshot` in the build container "
                            "(oracle/_ref/ then travels with the working tree)")


def available() -> bool:
    try:
        ref_root()
        return True
    except FileNotFoundError:
        return False


def kind() -> str:
    """What the CPU arm is called in bench.py's cpu_baseline.kind: the reference's own code."""
    return "reference"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        m.__path__ = []
        sys.modules[m.__name__] = m
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return self


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


def install_stubs() -> None:
    """Empty stand-ins for the packages of STUBBED that are really absent (import-time only; none is used by
    Diffusion_Encoder / sample_fun / the samplers)."""
    missing = []
    for name in STUBBED:
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is None:
                missing.append(name)
        except (ImportError, ValueError):
            missing.append(name)
    if missing and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder(missing))


def activate(arm: str) -> str:
    """Make `unet1d` / `sampler` resolve to the reference's packages (arm='reference') or to the drop-in packages
    (arm='dropin', everything else still from the reference). Returns the reference root."""
    assert arm in ("reference", "dropin")
    for name in ("unet1d", "sampler", "model3"):
        if name in sys.modules:
            raise RuntimeError(f"{name!r} is already imported in this process; one process hosts one arm")
    root = ref_root()
    install_stubs()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if arm == "dropin":
        import diffvits_b200
        sys.path.append(root)               # reference files behind the drop-in packages
        diffvits_b200.install_dropin()      # drop-in `unet1d` / `sampler` first on sys.path
    else:
        sys.path.insert(0, root)
    return root


def import_model3():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("model3")


def load_config() -> dict:
    with open(os.path.join(ref_root(), "config.json")) as f:
        return json.load(f)


def run_arm(script: str, arm: str, args=(), timeout: int = 1800, env=None) -> subprocess.CompletedProcess:
    """Run `script` (a file under tests/helpers or tools) in a fresh interpreter with DV_ARM=<arm>."""
    e = dict(os.environ)
    e["DV_ARM"] = arm
    e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
    if env:
        e.update(env)
    return subprocess.run([sys.executable, script, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=e)

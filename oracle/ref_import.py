"""ORACLE helper (test infrastructure): import the *reference* implementation from /root/reference (read-only,
present only in the build container) or from its snapshot oracle/_ref/ under private module names, without disturbing the drop-in `unet1d` /
`sampler` packages. Used by oracle/make_golden.py and by the tests that pin the oracle against the live reference.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
import warnings

def _root() -> str:
    """The reference checkout (build container) or its snapshot oracle/_ref/ (travels to the GPU box)."""
    from oracle import ref_snapshot
    for cand in (os.environ.get("DIFFVITS_REFERENCE", "/root/reference"), ref_snapshot.SNAPSHOT):
        if os.path.isfile(os.path.join(cand, "unet1d", "unet_1d_condition.py")):
            return cand
    return os.environ.get("DIFFVITS_REFERENCE", "/root/reference")


REF_ROOT = _root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "unet1d", "unet_1d_condition.py"))


def _load_file(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def load_samplers():
    """(ref_dpm_solver module, ref_uni_pc module) - both files are self-contained."""
    return (_load_file("_ref_dpm_solver", os.path.join(REF_ROOT, "sampler", "dpm_solver.py")),
            _load_file("_ref_uni_pc", os.path.join(REF_ROOT, "sampler", "uni_pc.py")))


def load_unet_module():
    """The reference `unet1d` package imported as `_ref_unet1d` (it uses relative imports only)."""
    if "_ref_unet1d" in sys.modules:
        return importlib.import_module("_ref_unet1d.unet_1d_condition")
    pkg_dir = os.path.join(REF_ROOT, "unet1d")
    spec = importlib.util.spec_from_file_location("_ref_unet1d", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["_ref_unet1d"] = pkg
    spec.loader.exec_module(pkg)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("_ref_unet1d.unet_1d_condition")


# Diffusion_Encoder's UNet constructor arguments (model3.py:887-896 with config.json:40-46). The oracle owns its copy of
# the constants; tests/test_host.py checks that the product's workload definition (diffvits_b200.synthetic) agrees.
UNET_KW = dict(in_channels=228, out_channels=100, block_out_channels=(128, 256, 384, 512), norm_num_groups=8,
               cross_attention_dim=128, attention_head_dim=8, addition_embed_type="text",
               resnet_time_scale_shift="scale_shift")


def build_ref_unet(seed: int = 1234):
    import torch
    mod = load_unet_module()
    torch.manual_seed(seed)
    return mod.UNet1DConditionModel(**UNET_KW).eval()

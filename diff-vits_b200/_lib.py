"""ctypes binding of libdiffvits_b200.so (C ABI declared in include/diffvits_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DV_DEBUG_LIB=1 loads the measurement build (libdiffvits_b200_dbg.so, include/diffvits_b200_debug.h) instead of the
# product library; it is only ever selected explicitly, by profiling tools.
DEBUG_LIB = os.environ.get("DV_DEBUG_LIB", "0") == "1"
LIB_PATH = os.path.join(_HERE, "libdiffvits_b200_dbg.so" if DEBUG_LIB else "libdiffvits_b200.so")


class DVError(RuntimeError):
    """A libdiffvits_b200 call returned a negative dv_status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"diffvits_b200 error {code}: {msg}")
        self.code = code


class UnetCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int),
        ("out_channels", C.c_int),
        ("n_levels", C.c_int),
        ("block_out_channels", C.c_int * 8),
        ("layers_per_block", C.c_int),
        ("norm_num_groups", C.c_int),
        ("cross_attention_dim", C.c_int),
        ("num_heads", C.c_int),
        ("norm_eps", C.c_float),
        ("x_channels", C.c_int),
        ("precision", C.c_int),
    ]


class PromptCfg(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("hidden", C.c_int), ("n_layers", C.c_int), ("num_heads", C.c_int),
                ("ffn_kernel", C.c_int), ("ffn_mult", C.c_int)]


class GemmSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("L", C.c_int), ("C", C.c_int), ("ld", C.c_int), ("bs", C.c_longlong)]


class GemmSeg(C.Structure):
    _fields_ = [("src", C.c_int), ("row_off", C.c_int), ("c0", C.c_int), ("npanels", C.c_int), ("ntap", C.c_int),
                ("wk0", C.c_int), ("tap_stride", C.c_int), ("xf", C.c_int), ("xf_c0", C.c_int)]

    def __init__(self, src=0, row_off=0, c0=0, npanels=0, ntap=1, wk0=-1, tap_stride=0, xf=0, xf_c0=0):
        super().__init__(src, row_off, c0, npanels, ntap, wk0, tap_stride, xf, xf_c0)


class GemmNorm(C.Structure):
    _fields_ = [("stats", C.c_void_p * 2), ("chunks", C.c_int * 2), ("nblk", C.c_int * 2), ("fold", C.c_int * 2),
                ("cover", C.c_int * 2), ("nst", C.c_int), ("C", C.c_int), ("L", C.c_int), ("G", C.c_int),
                ("eps", C.c_float), ("silu", C.c_int), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("scale_shift", C.c_void_p), ("ss_ld", C.c_int)]


class GemmOpts(C.Structure):
    _fields_ = [("bn", C.c_int), ("stats", C.c_void_p), ("out_bs", C.c_longlong), ("trim_col", C.c_int),
                ("norm", C.POINTER(GemmNorm))]


# name -> (restype, argtypes); mirrors include/diffvits_b200.h one to one
_vp, _i, _f, _sz, _ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
SIGNATURES = {
    "dv_version": (_i, []),
    "dv_last_error": (C.c_char_p, []),
    "dv_plan_create": (_i, [C.POINTER(UnetCfg), _i, _i, _i, C.POINTER(_vp)]),
    "dv_plan_destroy": (None, [_vp]),
    "dv_plan_workspace_bytes": (_sz, [_vp]),
    "dv_plan_param_bytes": (_sz, [_vp]),
    "dv_plan_manifest_json": (C.c_char_p, [_vp]),
    "dv_plan_launches_per_forward": (_i, [_vp]),
    "dv_plan_last_sample_launches": (_i, [_vp]),
    "dv_plan_sample_diag": (_i, [_vp, C.POINTER(_sz), C.POINTER(_i)]),
    "dv_plan_health_flag": (_i, [_vp, C.POINTER(_sz)]),
    "dv_plan_bind": (_i, [_vp, _vp, _vp, _vp]),
    "dv_prompt_prepare": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "dv_set_cond": (_i, [_vp, _vp, _vp]),
    "dv_unet_forward": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "dv_sample": (_i, [_vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _i, _vp]),
    "dv_sample_group": (_i, [C.POINTER(_vp), _i, _i, _i, C.POINTER(_f), C.POINTER(_f), C.POINTER(_vp), _i, _vp]),
    "dv_prompt_plan_create": (_i, [C.POINTER(PromptCfg), _i, _i, C.POINTER(_vp)]),
    "dv_prompt_plan_destroy": (None, [_vp]),
    "dv_prompt_plan_workspace_bytes": (_sz, [_vp]),
    "dv_prompt_plan_param_bytes": (_sz, [_vp]),
    "dv_prompt_plan_manifest_json": (C.c_char_p, [_vp]),
    "dv_prompt_plan_bind": (_i, [_vp, _vp, _vp, _vp]),
    "dv_prompt_encode": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "dv_comm_unique_id": (_i, [C.c_char_p, _vp]),
    "dv_comm_init": (_i, [C.c_char_p, _vp, _i, _i, C.POINTER(_vp)]),
    "dv_comm_all_gather": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "dv_comm_destroy": (None, [_vp]),
    "dv_plan_tensor_count": (_i, [_vp]),
    "dv_plan_tensor_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_sz), C.POINTER(_i), C.POINTER(_i)]),
    "dv_plan_op_count": (_i, [_vp]),
    "dv_plan_op_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dv_plan_profile_forward": (_i, [_vp, _vp, _i, _vp, _i, _vp, C.POINTER(_f), _i, _vp]),
    "dv_gemm": (_i, [C.POINTER(GemmSrc), _i, C.POINTER(GemmSeg), _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i,
                     _i, _i, C.POINTER(GemmOpts), _vp]),
    "dv_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "dv_groupnorm": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _f, _i, _vp, _vp]),
    "dv_groupnorm_scratch_floats": (_sz, [_i, _i, _i]),
    "dv_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "dv_dpmpp_update": (_i, [_vp, _vp, _vp, _vp, _f, _f, _f, _i, _f, _f, _sz, _vp]),
    "dv_unipc_predict": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _sz, _vp]),
    "dv_unipc_correct": (_i, [_vp, _vp, _vp, _vp, _f, _sz, _vp]),
    "dv_nct_to_ntc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
}

# include/diffvits_b200_debug.h: present only in the debug library
DEBUG_SIGNATURES = {
    "dv_gemm_test_trace": (None, [_vp]),
    "dv_gemm_test_flags": (None, [_i]),
    "dv_debug_skip_kinds": (_i, [_i]),
    "dv_debug_trace_op": (_i, [_vp, _i, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (building is the job of __graft_entry__.build / diffvits_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m diffvits_b200.build` (needs nvcc); "
            "diffvits_b200 has no CPU or PyTorch fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    sigs = dict(SIGNATURES)
    if DEBUG_LIB:
        sigs.update(DEBUG_SIGNATURES)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise DVError(rc, load().dv_last_error().decode("utf-8", "replace"))

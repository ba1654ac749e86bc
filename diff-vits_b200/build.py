"""In-tree build of libdiffvits_b200.so (sm_100a only) with plain nvcc; cross-compiles without a GPU.

The .so has no torch / pybind dependency: it exports the C ABI of include/diffvits_b200.h and is loaded with
ctypes by diffvits_b200._lib. Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# DV_DEBUG=1 builds the measurement variant (-DDV_DEBUG_HOOKS: timing-ablation switches, CTA timelines, dv_debug_* entry
# points of include/diffvits_b200_debug.h) as a separate library; the product library never contains them.
DEBUG = os.environ.get("DV_DEBUG", "0") == "1"
BUILD = os.path.join(HERE, "build_dbg" if DEBUG else "build")
LIB = os.path.join(HERE, "libdiffvits_b200_dbg.so" if DEBUG else "libdiffvits_b200.so")
SOURCES = ["gemm_tc.cu", "gemm_tc2.cu", "chain_tc.cu", "attention.cu", "attention_tc.cu", "attention_split.cu", "norm.cu", "elementwise.cu", "engine.cu", "prompt_encoder.cu", "comm.cu", "api.cu"]
KERNEL_SOURCES = ["gemm_tc.cu", "gemm_tc2.cu", "chain_tc.cu", "attention.cu", "attention_tc.cu", "attention_split.cu", "norm.cu",
                  "elementwise.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-diag-suppress", "550",
] + (["-DDV_DEBUG_HOOKS"] if DEBUG else [])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: diffvits_b200 has no prebuilt or CPU fallback path")


def _newest_header() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    hs.append(os.path.join(HERE, "..", "include", "diffvits_b200.h"))
    hs.append(os.path.join(HERE, "..", "include", "diffvits_b200_debug.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _newest_header()
    jobs = []
    objs = []
    # every kernel source twice: BF16 operands (namespace dv) and FP16 operands (-DDV_F16, namespace dv16, mode "fp16s")
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        for suffix, extra in ((".o", []), ("_f16.o", ["-DDV_F16"])):
            if extra and src not in KERNEL_SOURCES:
                continue
            o = os.path.join(BUILD, src.replace(".cu", suffix))
            objs.append(o)
            if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
                jobs.append((s, o, extra))

    def compile_one(job):
        s, o, extra = job
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Pipelines of the tap-GEMM timed separately (debug flags) on big problems so host overhead is negligible."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as U
from diffvits_b200 import _lib
lib = _lib.load()
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, generator=g, device="cuda")).to(torch.bfloat16)
print("B L K N mode bn | tiles/CTA | full | no-epilogue | no-mma | no-mma-no-tma | epilogue-only(no mma, no tma)   [us, cycles/tile]")
for (B, L, K, N, mode, bn) in [(256, 1000, 128, 128, 0, 128), (256, 1000, 384, 128, 0, 128), (256, 1000, 1024, 128, 0, 128),
                               (64, 1000, 128, 1024, 2, 128), (128, 1000, 1024, 512, 0, 128), (128, 1000, 1024, 512, 0, 256),
                               (128, 1000, 128, 512, 0, 256), (128, 1000, 128, 512, 0, 128)]:
    a = rnd(B, L, K); w = rnd(N, K); bias = torch.randn(N, device="cuda")
    res = rnd(B, L, N) if mode == 0 else None
    tiles = B * 8 * (N // bn)
    per = (tiles + 147) // 148
    out = []
    for flags in (0, 1, 2, 6, 7):
        lib.dv_gemm_test_flags(flags)
        us = timeit(lambda: U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res, out_mode=mode, bn=bn))
        out.append(f"{us:7.1f} ({us * 1920 / per:6.0f})")
    lib.dv_gemm_test_flags(0)
    print(B, L, K, N, mode, bn, "|", per, "|", " | ".join(out))

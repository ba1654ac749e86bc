# Needs the debug library: DV_DEBUG=1 python -m diffvits_b200.build ; run with DV_DEBUG_LIB=1 (include/diffvits_b200_debug.h)
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
names = {0: "full", 1: "no-epi", 6: "epi only", 6 + 8: "epi -res", 6 + 16: "epi -stg", 6 + 8 + 16: "epi -res-stg", 6 + 8 + 16 + 32: "epi -res-stg-ldtm", 2: "tma+epi"}
for (B, L, K, N, mode, bn, withres) in [(512, 1000, 128, 128, 0, 128, True), (512, 1000, 128, 128, 0, 128, False), (512, 1000, 384, 128, 0, 128, True)]:
    a = rnd(B, L, K); w = rnd(N, K); bias = torch.randn(N, device="cuda")
    res = rnd(B, L, N) if withres else None
    tiles = B * 8 * (N // bn); per = (tiles + 147) // 148
    print(f"B={B} K={K} N={N} res={withres} tiles/CTA={per}")
    for flags, nm in names.items():
        lib.dv_gemm_test_flags(flags)
        us = timeit(lambda: U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res, out_mode=mode, bn=bn))
        print(f"   {nm:20s} {us:8.1f} us  {us * 1920 / per:7.0f} cyc/tile")
    lib.dv_gemm_test_flags(0)

# Needs the debug library: DV_DEBUG=1 python -m diffvits_b200.build ; run with DV_DEBUG_LIB=1 (include/diffvits_b200_debug.h)
"""Timeline of CTA 0 of one tap-GEMM launch (debug trace built into the kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as U
from diffvits_b200 import _lib
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, generator=g, device="cuda")).to(torch.bfloat16)
for (B, L, K, N, mode, bn) in [(32, 1000, 128, 1024, 2, 128), (32, 1000, 128, 128, 0, 128), (32, 1000, 384, 128, 0, 128)]:
    a = rnd(B, L, K); w = rnd(N, K); bias = torch.randn(N, device="cuda")
    res = rnd(B, L, N) if mode == 0 else None
    U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res, out_mode=mode, bn=bn)   # warm
    buf = torch.zeros(4001, dtype=torch.int64, device="cuda")
    lib.dv_gemm_test_trace(U.vp(buf))
    U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res, out_mode=mode, bn=bn)
    lib.dv_gemm_test_trace(None)
    torch.cuda.synchronize()
    h = buf.cpu().tolist()
    n = min(h[0], 4000)
    recs = []
    for v in h[1:1 + n]:
        v &= (1 << 64) - 1
        recs.append((v & 0xFFFFFFFFFF, (v >> 60) & 0xF, (v >> 44) & 0xFFFF, (v >> 40) & 0xF))
    recs.sort()
    t0 = recs[0][0]
    print(f"== B={B} L={L} K={K} N={N} mode={mode} bn={bn}: {n} records")
    names = {(0, 0): "prod tile start", (1, 0): "mma wait tmem_empty", (1, 1): "mma got tmem", (1, 2): "mma committed",
             (2, 0): "epi wait full", (2, 1): "epi got full", (2, 2): "epi tile done"}
    for t, role, tile, ev in recs[:60]:
        print(f"  {t - t0:8d} cyc  tile {tile:5d}  {names.get((role, ev), (role, ev))}")

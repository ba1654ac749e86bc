# Needs the debug library: DV_DEBUG=1 python -m diffvits_b200.build ; run with DV_DEBUG_LIB=1 (include/diffvits_b200_debug.h)
"""Timeline of CTA (0,0,0) of one tcgen05 attention launch (debug trace built into the kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as U
from diffvits_b200 import _lib
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
B, H = 32, 8
names = {(0, 0): "tma  slot free, issue K/V", (1, 0): "mma  K/V landed", (1, 1): "mma  S issued+committed", (1, 2): "mma  P_FULL seen",
         (1, 3): "mma  PV issued+committed", (2, 0): "smx  tile begin (wait S)", (2, 1): "smx  S_FULL seen", (2, 2): "smx  S in regs",
         (2, 3): "smx  max+exchange done", (2, 4): "smx  exp+P stores done", (2, 5): "smx  fence + P_FULL arrive",
         (2, 6): "smx  O_FULL(prev) seen", (2, 7): "smx  fold done"}
for (D, L) in [(16, 1000), (32, 500)]:
    C = H * D
    qkv = torch.randn(B, L, 3 * C, generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    U.attention(q, k, v, H)
    buf = torch.zeros(4001, dtype=torch.int64, device="cuda")
    lib.dv_gemm_test_trace(U.vp(buf))
    U.attention(q, k, v, H)
    lib.dv_gemm_test_trace(None)
    torch.cuda.synchronize()
    h = buf.cpu().tolist()
    recs = []
    for x in h[1:]:
        x &= (1 << 64) - 1
        if x:
            recs.append((x & 0xFFFFFFFFFF, ((x >> 60) & 0xF) - 1, (x >> 44) & 0xFFFF, (x >> 40) & 0xF))
    n = len(recs)
    recs.sort()
    t0 = recs[0][0]
    print(f"== D={D} L={L}: {n} records")
    for t, role, tile, ev in recs[:150]:
        print(f"  {t - t0:8d} cyc  tile {tile:3d}  {names.get((role, ev), (role, ev))}")

# Needs the debug library: DV_DEBUG=1 python -m diffvits_b200.build ; run with DV_DEBUG_LIB=1 (include/diffvits_b200_debug.h)
"""Timeline of CTA 0 of consecutive tap-GEMM launches (debug trace built into the kernel, no atomics): four identical launches
back to back on one stream (PDL chain as in the product); the third and fourth are printed. SM clock cycles per role,
%globaltimer (ns) for kernel entry / previous-kernel-complete / exit so that the launches share one time axis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as U
from diffvits_b200 import _lib
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
REGION = 160
def rnd(*s): return (torch.randn(*s, generator=g, device="cuda")).to(torch.bfloat16)
NAMES = {(0, 8): "kernel entry", (0, 9): "prologue done (barriers, TMEM, bias)", (0, 10): "griddepcontrol.wait passed",
         (0, 0): "producer: tile start (A loads issued)", (0, 11): "kernel exit",
         (1, 0): "mma: wait tmem_empty", (1, 1): "mma: got accumulator", (1, 3): "mma: first panel landed", (1, 2): "mma: all panels issued + commit",
         (2, 0): "epi: wait accumulator", (2, 1): "epi: got accumulator", (2, 2): "epi: tile done",
         (2, 6): "epi chunk: TMEM load complete", (2, 7): "epi chunk: math done", (2, 8): "epi chunk: stores issued",
         (2, 3): "epi: pass 1 done", (2, 4): "epi: item complete", (2, 5): "epi: statistics folded"}
def main():
    CASES = [  # (B, L, K, N, mode, with_res, bn)
        (32, 250, 384, 384, 0, False, 192), (32, 250, 384, 384, 0, True, 192), (32, 500, 256, 768, 0, False, 256),
        (32, 1000, 128, 128, 0, False, 128), (32, 125, 512, 512, 0, False, 128), (32, 500, 256, 2048, 2, False, 256)]
    import ctypes as C
    from diffvits_b200._lib import GemmOpts, GemmSeg, GemmSrc, check
    for (B, L, K, N, mode, withres, bn) in CASES:
        a = rnd(B, L, K); w = rnd(N, K); bias = torch.randn(N, device="cuda")
        res = rnd(B, L, N) if withres else None
        out = torch.zeros(B, L, N // 2 if mode == 2 else N, dtype=torch.bfloat16, device="cuda")
        S = (GemmSrc * 1)(GemmSrc(a.data_ptr(), L, K, K))
        G = (GemmSeg * 1)(GemmSeg(0, 0, 0, K // 64))
        opts = GemmOpts(bn, None, 0, 0)

        def launch():
            check(lib.dv_gemm(S, 1, G, 1, U.vp(w), K, N, N, U.vp(bias), U.vp(res), 0 if res is None else N, U.vp(out),
                              out.shape[-1], mode, B, L, 0, C.byref(opts), U.stream()))
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        bufs = [torch.zeros(1 + 4 * REGION, dtype=torch.int64, device="cuda") for _ in range(4)]
        torch.cuda.synchronize()
        for b_ in bufs:
            lib.dv_gemm_test_trace(U.vp(b_))
            launch()
        lib.dv_gemm_test_trace(None)
        torch.cuda.synchronize()
        print(f"== B={B} L={L} K={K} N={N} mode={mode} res={withres} bn={bn}")
        gt0 = None
        for li, b_ in enumerate(bufs):
            h = b_.cpu().tolist()
            recs, gts = [], {}
            for role in range(4):
                base = 1 + role * REGION
                n = h[base] & ((1 << 64) - 1)
                for v in h[base + 1: base + 1 + n]:
                    v &= (1 << 64) - 1
                    t, ev, tile = v & 0xFFFFFFFFFF, (v >> 40) & 0xF, (v >> 44) & 0xFFFF
                    if role == 3:
                        gts[ev] = t
                    else:
                        recs.append((t, role, tile, ev))
            if gt0 is None:
                gt0 = gts.get(0, 0)
            if li < 2:
                continue
            recs.sort()
            t0 = recs[0][0]
            print(f"  -- launch {li}: globaltimer entry {gts.get(0, 0) - gt0:7d} ns, wait passed {gts.get(1, 0) - gt0:7d} ns, exit {gts.get(2, 0) - gt0:7d} ns")
            for t, role, tile, ev in recs:
                print(f"     {t - t0:7d} clk  {NAMES.get((role, ev), (role, ev))}" + (f"  (tile {tile})" if ev < 8 and tile else ""))


if __name__ == "__main__":
    main()

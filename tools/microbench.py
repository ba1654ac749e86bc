"""Steady-state device time of every kernel class at the real per-level shapes of a workload (default C2:
B=32, T=1000): each case is launched ITERS times back to back on one stream (PDL on, as in the product) and the
CUDA-event interval is divided by ITERS. Prints one line per case: us, TFLOP/s or GB/s (algorithmic).

    python tools/microbench.py [--B 32] [--T 1000] [--only gemm,attn,gn,ln] [--flags N] [--lib]

--lib adds the "library bar" of SURVEY.md section 8d next to every case: the op the reference would run for the same shape
through PyTorch on the same GPU in BF16 (cuDNN conv1d, cuBLASLt linear (+ the separate GEGLU / residual kernels), SDPA,
native group_norm + silu, layer_norm, the solver update expression), timed the same way (20 launches in one CUDA graph).
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import gpu_util as U  # noqa: E402
from diffvits_b200 import _lib  # noqa: E402
from diffvits_b200._lib import GemmOpts, GemmSeg, GemmSrc, check  # noqa: E402

ITERS = 20


def timeit(fn, iters=ITERS):
    """fn(stream) enqueues one launch. ITERS launches are captured into one CUDA graph (no CPU launch cost, PDL edges
    as in the product) and the replay is timed; best of 5."""
    fn(U.stream())
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters):
            fn(U.stream())
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def time_torch(fn, iters=ITERS):
    """The same measurement for a PyTorch op: `iters` calls captured into one CUDA graph on a side stream, best of 5."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr), torch.no_grad():
        for _ in range(iters):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--Tp", type=int, default=282)
    ap.add_argument("--only", default="gemm,attn,gn,ln,samp")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--taps3", action="store_true", help="conv k3 as three single-tap segments (three A loads per panel)")
    ap.add_argument("--lib", action="store_true", help="also time the PyTorch library op of the same shape (cuDNN / cuBLASLt / SDPA)")
    a = ap.parse_args()
    only = set(a.only.split(","))
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(0)

    def rnd(*s):
        return torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)

    B = a.B
    levels = [(a.T, 128), (a.T // 2, 256), (a.T // 4, 384), (a.T // 8, 512)]
    if "gemm" in only:
        if a.flags:
            lib.dv_gemm_test_flags(a.flags)      # debug library only (DV_DEBUG_LIB=1)
        print(f"# GEMM (dbg flags {a.flags}): tag, us, TFLOP/s, algorithmic MB")
        for li, (L, Cc) in enumerate(levels):
            cases = [
                ("conv3 C->C +res+stats", 3, Cc, Cc, 0, True, True),
                ("lin   C->C +res", 1, Cc, Cc, 0, True, False),
                ("lin   C->C", 1, Cc, Cc, 0, False, False),
                ("qkv   C->3C", 1, Cc, 3 * Cc, 0, False, False),
                ("geglu C->8C", 1, Cc, 8 * Cc, 2, False, False),
                ("ffout 4C->C +res", 1, 4 * Cc, Cc, 0, True, False),
            ]
            if li == 0:
                cases.append(("conv3 2C->C (cat) ", 3, 2 * Cc, Cc, 0, False, True))
            for tag, taps, K1, N, mode, withres, stats in cases:
                src = rnd(B, L, K1)
                ktot = taps * K1
                w = rnd(N, ktot)
                bias = torch.randn(N, device="cuda")
                res = rnd(B, L, N) if withres else None
                out = torch.zeros(B, L, N // 2 if mode == 2 else N, dtype=torch.bfloat16, device="cuda")
                S = (GemmSrc * 1)(GemmSrc(src.data_ptr(), L, K1, K1))
                # conv k3: one staged 136-row box read by three row-shifted descriptors (default), or three loads (--taps3)
                if taps == 3 and not a.taps3:
                    G = (GemmSeg * 1)(GemmSeg(0, -1, 0, K1 // 64, 3))
                    segs = [None]
                else:
                    segs = [(0, t - 1, 0, K1) for t in range(3)] if taps == 3 else [(0, 0, 0, K1)]
                    G = (GemmSeg * len(segs))(*[GemmSeg(s, ro, c0, ch // 64) for (s, ro, c0, ch) in segs])
                stt = torch.zeros(B, (L + 127) // 128 * 4, N // 16, 2, device="cuda") if stats else None
                lib_us = None
                if a.lib:
                    import torch.nn.functional as F
                    if taps == 3:      # cuDNN conv1d k3 p1 on NCL bf16 (+ the residual add the reference does separately)
                        xin = src.transpose(1, 2).contiguous()
                        wc = w.view(N, 3, K1).permute(0, 2, 1).contiguous()
                        rr = res.transpose(1, 2).contiguous() if withres else None
                        bb = bias.to(torch.bfloat16)
                        lib_us = time_torch(lambda: (F.conv1d(xin, wc, bb, padding=1) + rr) if withres else F.conv1d(xin, wc, bb, padding=1))
                    else:              # cuBLASLt linear on [B*L, K] bf16 (+ GEGLU chunk/gelu/mul or the residual add)
                        x2 = src.view(B * L, K1)
                        bb = bias.to(torch.bfloat16)
                        r2 = res.view(B * L, N) if withres else None
                        if mode == 2:
                            def lf():
                                u, gt = F.linear(x2, w, bb).chunk(2, dim=-1)
                                return u * F.gelu(gt)
                            lib_us = time_torch(lf)
                        else:
                            lib_us = time_torch(lambda: (F.linear(x2, w, bb) + r2) if withres else F.linear(x2, w, bb))
                for bn in (64, 128, 192, 256):
                    if N % bn or (bn == 192 and (N != 384 or mode == 2)) or (bn == 64 and (mode == 2 or B * L > 4096)):
                        continue

                    def fn(st):
                        opts = GemmOpts(bn, None if stt is None else stt.data_ptr(), 0, 0)
                        check(lib.dv_gemm(S, 1, G, len(segs), U.vp(w), ktot, N, N, U.vp(bias), U.vp(res),
                                          0 if res is None else N, U.vp(out), out.shape[-1], mode, B, L, 0,
                                          C.byref(opts), st))
                    us = timeit(fn)
                    fl = 2.0 * B * L * N * ktot
                    mb = (B * L * K1 * 2 + N * ktot * 2 + out.numel() * 2 + (B * L * N * 2 if withres else 0)) / 1e6
                    print(f"L{li} {tag:24s} bn={bn} M={B * L:6d} N={N:5d} K={ktot:5d}  {us:8.2f} us  "
                          f"{fl / us / 1e6:7.1f} TF/s  {mb:7.2f} MB" +
                          (f"   | torch bf16 {lib_us:8.2f} us ({lib_us / us:4.2f}x)" if lib_us else ""))
        if a.flags:
            lib.dv_gemm_test_flags(0)
    if "attn" in only:
        print("# attention: tag, us, TFLOP/s, Gexp/s")
        for li, (L, Cc) in enumerate(levels):
            D = Cc // 8
            qkv = rnd(B, L, 3 * Cc)
            q, k, v = qkv[:, :, :Cc], qkv[:, :, Cc:2 * Cc], qkv[:, :, 2 * Cc:]
            o = torch.zeros(B, L, Cc, dtype=torch.bfloat16, device="cuda")
            kv = rnd(B, a.Tp, 2 * Cc)
            kc, vc = kv[:, :, :Cc], kv[:, :, Cc:]
            qd = rnd(B, L, Cc)
            for tag, qq, kk, vv, Lk in (("self ", q, k, v, L), ("cross", qd, kc, vc, a.Tp)):
                def fn(st):
                    check(lib.dv_attention(U.vp(qq), U.vp(kk), U.vp(vv), U.vp(o), qq.stride(1), kk.stride(1), vv.stride(1),
                                           o.stride(1), qq.stride(0), kk.stride(0), vv.stride(0), o.stride(0), B, 8, L, Lk,
                                           D, float(D) ** -0.5, None, st))
                us = timeit(fn)
                fl = 4.0 * B * 8 * L * Lk * D
                extra = ""
                if a.lib:
                    import torch.nn.functional as F
                    q4 = qq.reshape(B, L, 8, D).transpose(1, 2).contiguous()
                    k4 = kk.reshape(B, Lk, 8, D).transpose(1, 2).contiguous()
                    v4 = vv.reshape(B, Lk, 8, D).transpose(1, 2).contiguous()
                    # the reference passes an additive [B, 1, 1, Tp] bias on the cross-attention (attention_processor.py:1003-1007)
                    bias4 = torch.zeros(B, 1, 1, Lk, dtype=torch.bfloat16, device="cuda") if tag.strip() == "cross" else None
                    lus = time_torch(lambda: F.scaled_dot_product_attention(q4, k4, v4, attn_mask=bias4))
                    extra = f"   | torch SDPA bf16 {lus:8.2f} us ({lus / us:4.2f}x)"
                print(f"L{li} attn {tag} D={D:2d} Lq={L:5d} Lk={Lk:5d}  {us:8.2f} us  {fl / us / 1e6:7.1f} TF/s  "
                      f"{B * 8 * L * Lk / us / 1e3:7.1f} Gexp/s" + extra)
    if "gn" in only:
        print("# gn_apply (+stats kernel): tag, us (apply only), GB/s (read+write bf16)")
        for li, (L, Cc) in enumerate(levels):
            for c0, c1 in ((Cc, 0), (Cc, Cc)):
                x0 = rnd(B, L, c0)
                x1 = rnd(B, L, c1) if c1 else None
                Ct = c0 + c1
                gamma, beta = torch.randn(Ct, device="cuda"), torch.randn(Ct, device="cuda")
                out = torch.zeros(B, L, Ct, dtype=torch.bfloat16, device="cuda")
                stats = torch.empty(lib.dv_groupnorm_scratch_floats(B, L, Ct), dtype=torch.float32, device="cuda")

                def fn(st):
                    check(lib.dv_groupnorm(U.vp(x0), c0, U.vp(x1), c1, B, L, 8, U.vp(stats), U.vp(gamma), U.vp(beta), None, 0,
                                           1e-5, 1, U.vp(out), st))
                us = timeit(fn)
                extra = ""
                if a.lib:
                    import torch.nn.functional as F
                    xn = (torch.cat([x0, x1], -1) if c1 else x0).transpose(1, 2).contiguous()
                    gb, bb = gamma.to(torch.bfloat16), beta.to(torch.bfloat16)
                    lus = time_torch(lambda: F.silu(F.group_norm(xn, 8, gb, bb, 1e-5)))
                    extra = f"   | torch group_norm+silu bf16 {lus:8.2f} us ({lus / us:4.2f}x)"
                print(f"L{li} gn stats+apply C={c0}+{c1} L={L:5d}  {us:8.2f} us (two kernels)  "
                      f"{6.0 * B * L * Ct / us / 1e3:7.1f} GB/s (3 passes)" + extra)
    if "ln" in only:
        print("# layernorm: tag, us, GB/s")
        for li, (L, Cc) in enumerate(levels):
            x = rnd(B, L, Cc)
            gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
            out = torch.zeros_like(x)

            def fn(st):
                check(lib.dv_layernorm(U.vp(x), U.vp(out), U.vp(gamma), U.vp(beta), B * L, Cc, 1e-5, st))
            us = timeit(fn)
            extra = ""
            if a.lib:
                import torch.nn.functional as F
                gb, bb = gamma.to(torch.bfloat16), beta.to(torch.bfloat16)
                lus = time_torch(lambda: F.layer_norm(x, (Cc,), gb, bb, 1e-5))
                extra = f"   | torch layer_norm bf16 {lus:8.2f} us ({lus / us:4.2f}x)"
            print(f"L{li} layernorm C={Cc} rows={B * L:6d}  {us:8.2f} us  {4.0 * B * L * Cc / us / 1e3:7.1f} GB/s" + extra)
    if "samp" in only:
        n = B * 100 * a.T
        x = torch.randn(n, device="cuda")
        m1 = torch.randn(n, device="cuda")
        m0 = torch.randn(n, device="cuda")

        def fn(st):
            check(lib.dv_dpmpp_update(U.vp(x), U.vp(x), U.vp(m1), U.vp(m0), 0.9, 0.1, 0.01, 0, 1.0, 1.0, n, st))
        us = timeit(fn)
        extra = ""
        if a.lib:
            y = torch.empty_like(x)
            lus = time_torch(lambda: torch.add(torch.add(x * 0.9, m1, alpha=0.1), m0, alpha=0.01, out=y))
            extra = f"   | torch a*x + b*m1 + c*m0 (fp32, 3 kernels) {lus:8.2f} us ({lus / us:4.2f}x)"
        print(f"dpmpp_update n={n}  {us:8.2f} us  {16.0 * n / us / 1e3:7.1f} GB/s" + extra)


if __name__ == "__main__":
    main()

"""Helpers for the -m gpu parity tests: call the C ABI (include/diffvits_b200.h) on torch CUDA tensors."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from diffvits_b200 import _lib
from diffvits_b200._lib import GemmNorm, GemmOpts, GemmSeg, GemmSrc, check


def vp(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max())


def bf16r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def gemm(srcs: Sequence[torch.Tensor], segs: Sequence[Tuple[int, int, int, int]], w: torch.Tensor, n: int,
         bias: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None, out_mode: int = 0, impl: int = 0,
         l_out: Optional[int] = None, views: Optional[Sequence[Tuple[int, int]]] = None, bn: int = 0,
         stats: bool = False, layout: Optional[Tuple[int, int]] = None, norm: Optional[dict] = None):
    """srcs: bf16 [B, L, C] tensors; segs: (src, row_off, c0, channels[, ntap, wk0, tap_stride, xf, xf_c0]); w: bf16 [n_pad, k_tot].
    views: optional (L_view, C_view) per source (the stride-2 conv reads [L/2, 2C]).
    layout: optional (out_bs, trim_col) of dv_gemm_opts; the result is then the flat [B, out_bs] buffer."""
    lib = _lib.load()
    B = srcs[0].shape[0]
    n_pad, k_tot = w.shape
    S = (GemmSrc * len(srcs))()
    for i, s in enumerate(srcs):
        assert s.dtype == torch.bfloat16 and s.is_contiguous()
        L, Cc = (s.shape[1], s.shape[2]) if views is None else views[i]
        S[i] = GemmSrc(s.data_ptr(), L, Cc, Cc, s.shape[1] * s.shape[2])      # batch stride of the real tensor
    G = (GemmSeg * len(segs))()
    for i, sg in enumerate(segs):
        # (src, row_off, c0, channels[, ntap[, wk0[, tap_stride[, xf[, xf_c0]]]]])
        src, ro, c0, ch = sg[:4]
        extra = list(sg[4:]) + [1, -1, 0, 0, 0][len(sg) - 4:]
        G[i] = GemmSeg(src, ro, c0, ch // 64, extra[0], extra[1], extra[2], extra[3], extra[4])
    L_out = l_out if l_out is not None else srcs[0].shape[1]
    dev = srcs[0].device
    if impl == 1:
        out = torch.empty(B, L_out, n_pad, dtype=torch.float32, device=dev)
        ld_out = n_pad
    elif out_mode == 0:
        out = (torch.zeros(B, L_out, n, dtype=torch.bfloat16, device=dev) if layout is None else
               torch.full((B, layout[0]), 777.0, dtype=torch.bfloat16, device=dev))
        ld_out = n
    elif out_mode == 1:
        out = torch.zeros(B, n, L_out, dtype=torch.float32, device=dev)
        ld_out = 0
    else:
        out = torch.zeros(B, L_out, n // 2, dtype=torch.bfloat16, device=dev)
        ld_out = n // 2
    st = None
    if stats:
        st = torch.full((B, (L_out + 127) // 128 * 4, n // 16, 2), float("nan"), dtype=torch.float32, device=dev)
    opts = GemmOpts(bn, None if st is None else st.data_ptr(), 0 if layout is None else int(layout[0]),
                    0 if layout is None else int(layout[1]))
    if norm is not None:
        # GroupNorm fused into the operand path: norm = dict(stats=[partials per source], cover=[channels per source],
        # C, L, G, eps, silu, gamma, beta, scale_shift (or None))
        nm = GemmNorm()
        for i, (sp, cov) in enumerate(zip(norm["stats"], norm["cover"])):
            nm.stats[i] = sp.data_ptr()
            nm.chunks[i], nm.nblk[i], nm.fold[i], nm.cover[i] = sp.shape[1], sp.shape[2], 0, cov
        nm.nst = len(norm["stats"])
        nm.C, nm.L, nm.G, nm.eps, nm.silu = norm["C"], norm["L"], norm["G"], norm["eps"], 1 if norm["silu"] else 0
        nm.gamma, nm.beta = norm["gamma"].data_ptr(), norm["beta"].data_ptr()
        ss = norm.get("scale_shift")
        nm.scale_shift, nm.ss_ld = (None, 0) if ss is None else (ss.data_ptr(), ss.shape[1])
        opts.norm = C.pointer(nm)
    check(lib.dv_gemm(S, len(srcs), G, len(segs), vp(w), k_tot, n, n_pad, vp(bias), vp(res),
                      0 if res is None else res.shape[-1], vp(out), ld_out, out_mode, B, L_out, impl, C.byref(opts),
                      stream()))
    return (out, st) if stats else out


def attention(q, k, v, heads: int, kv_len: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B, Lq, C], k/v [B, Lk, C] bf16 (views with a last-dim stride of 1 are fine)."""
    lib = _lib.load()
    B, Lq, Cc = q.shape
    Lk = k.shape[1]
    D = Cc // heads
    o = torch.zeros(B, Lq, Cc, dtype=torch.bfloat16, device=q.device)
    check(lib.dv_attention(vp(q), vp(k), vp(v), vp(o), q.stride(1), k.stride(1), v.stride(1), o.stride(1),
                           q.stride(0), k.stride(0), v.stride(0), o.stride(0), B, heads, Lq, Lk, D,
                           float(D) ** -0.5, vp(kv_len), stream()))
    return o


def groupnorm(x0, x1, G, gamma, beta, eps, silu, ss=None) -> torch.Tensor:
    lib = _lib.load()
    B, L, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[2]
    out = torch.zeros(B, L, c0 + c1, dtype=torch.bfloat16, device=x0.device)
    stats = torch.empty(lib.dv_groupnorm_scratch_floats(B, L, c0 + c1), dtype=torch.float32, device=x0.device)
    check(lib.dv_groupnorm(vp(x0), c0, vp(x1), c1, B, L, G, vp(stats), vp(gamma), vp(beta), vp(ss),
                           0 if ss is None else ss.shape[1], eps, 1 if silu else 0, vp(out), stream()))
    return out


def layernorm(x, gamma, beta, eps=1e-5) -> torch.Tensor:
    lib = _lib.load()
    rows, Cc = x.shape[0] * x.shape[1], x.shape[2]
    out = torch.zeros_like(x)
    check(lib.dv_layernorm(vp(x), vp(out), vp(gamma), vp(beta), rows, Cc, eps, stream()))
    return out


def gn_partials(x: torch.Tensor) -> torch.Tensor:
    """GroupNorm partial statistics of a bf16 [B, L, C] tensor in the layout the GEMM epilogue / gn_stats emit:
    float [B][ceil(L/32)][C/16][2] = (sum, sum of squares) per 32-row chunk and 16-channel block."""
    B, L, Cc = x.shape
    rows = (L + 31) // 32 * 32
    y = torch.zeros(B, rows, Cc, dtype=torch.float32, device=x.device)
    y[:, :L] = x.float()
    blk = y.reshape(B, rows // 32, 32, Cc // 16, 16)
    return torch.stack([blk.sum(dim=(2, 4)), (blk * blk).sum(dim=(2, 4))], dim=-1).contiguous()

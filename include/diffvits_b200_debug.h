/*
 * diffvits_b200 - measurement hooks of the DEBUG build only (libdiffvits_b200_dbg.so, compiled with -DDV_DEBUG_HOOKS by
 * `DV_DEBUG=1 python -m diffvits_b200.build`). None of these symbols exists in the product library libdiffvits_b200.so,
 * and the kernel-side switches they drive are compiled out of it. They have no counterpart in the reference: they exist to
 * attribute time inside the captured loop (profiles/README.md) and results are garbage while any of them is active.
 */
#ifndef DIFFVITS_B200_DEBUG_H
#define DIFFVITS_B200_DEBUG_H
#include "diffvits_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Device buffer of >= 4001 uint64; CTA 0 of the next dv_gemm / dv_attention launches of this thread appends
 * (role<<60 | tile<<44 | event<<40 | clock) records, buf[0] = count. roles: 0 TMA producer, 1 MMA issuer, 2 first
 * epilogue warp. NULL = off. */
DV_API void dv_gemm_test_trace(unsigned long long* buf);
/* Next dv_gemm calls of this thread: bit 0 skip the epilogue body, bit 1 skip the MMA issue, bit 2 skip the TMA loads,
 * bit 3 skip the global stores, bit 4 skip the residual loads (times the pipelines of the kernel separately). The
 * engine's launches take the same bits from the environment variables DV_GEMM_DBG / DV_CHAIN_DBG. */
DV_API void dv_gemm_test_flags(int flags);
/* bit k set = the denoiser ops of kernel class k (0 gemm, 2 groupnorm apply, 3 layernorm, 4 attention, 5 chained gemm;
 * dv_plan_op_info's `kind`) are NOT launched by later dv_unet_forward / dv_sample calls, so the difference in step time
 * is that class's cost inside the captured loop. Initial value: DV_SKIP_KINDS (default 0). Returns the previous mask. */
DV_API int dv_debug_skip_kinds(int mask);
/* CTA-0 timeline (csrc/gemm_tc.cu trace_ev) of tap-GEMM op `op` of the plan's per-call program, for every later launch of it;
 * buf: device, 1 + 4 * 160 uint64 (NULL switches it off). tools/trace_plan_op.py reads it. */
DV_API int dv_debug_trace_op(dv_plan* plan, int op, unsigned long long* buf);
#ifdef __cplusplus
}
#endif
#endif /* DIFFVITS_B200_DEBUG_H */

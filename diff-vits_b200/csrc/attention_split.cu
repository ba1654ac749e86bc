// Attention core of the FP32-class precision mode ("fp32x3"): F.scaled_dot_product_attention
// (unet1d/attention_processor.py:1032, additive -10000 key-padding bias :1003-1007) on operands stored as hi + lo bf16
// planes, evaluated entirely in fp32 on the CUDA cores. Q.K^T, the softmax and P.V never round to bf16: the only
// rounding points are the hi + lo operands (16 mantissa bits) and the hi + lo output. This mode exists to prove the
// engine against the fp32 reference at the 1e-4 gate; the BF16 mode's tensor-core kernels (attention.cu /
// attention_tc.cu) are the fast path.
//
// CTA = 64 queries of one (batch, head) x 4 threads per query: thread `part` owns D/4 of the head dimension (its slice
// of q and of the output accumulator); a score is the sum of the four partial dot products (two shuffles). K/V tiles of
// 64 keys are reconstructed to fp32 (hi + lo) in shared memory; the 8 queries of a warp read the same key row (broadcast).
#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {
namespace {

constexpr int SQ = 64;      // queries per CTA
constexpr int SK = 64;      // keys per shared-memory tile
constexpr int SCH = 16;     // keys per online-softmax chunk

__device__ __forceinline__ float ld_hilo(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t i) {
  return h16_to_f32(hi[i]) + h16_to_f32(lo[i]);
}

template <int D>
__global__ void __launch_bounds__(SQ * 4) attention_split_kernel(const AttnParams p) {
  constexpr int DP = D / 4;                         // dims per thread
  pdl_launch_dependents();
  __shared__ __align__(16) float sk[SK][D];
  __shared__ __align__(16) float sv[SK][D];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * SQ + (threadIdx.x >> 2);
  const int part = threadIdx.x & 3;
  const bool qvalid = qi < p.Lq;
  const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(p.q) + static_cast<size_t>(b) * p.bs_q + h * D;
  const __nv_bfloat16* Ql = reinterpret_cast<const __nv_bfloat16*>(p.q_lo) + static_cast<size_t>(b) * p.bs_q + h * D;
  const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(p.k) + static_cast<size_t>(b) * p.bs_k + h * D;
  const __nv_bfloat16* Kl = reinterpret_cast<const __nv_bfloat16*>(p.k_lo) + static_cast<size_t>(b) * p.bs_k + h * D;
  const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(p.v) + static_cast<size_t>(b) * p.bs_v + h * D;
  const __nv_bfloat16* Vl = reinterpret_cast<const __nv_bfloat16*>(p.v_lo) + static_cast<size_t>(b) * p.bs_v + h * D;
  pdl_wait();
  const float LOG2E = 1.4426950408889634f;
  float q[DP], acc[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) {
    const size_t idx = static_cast<size_t>(qvalid ? qi : 0) * p.ld_q + part * DP + i;
    q[i] = ld_hilo(Q, Ql, idx) * (p.scale * LOG2E);   // scores in log2 units: softmax = 2^(s - max) / sum
    acc[i] = 0.f;
  }
  const int kvl = p.kv_len ? p.kv_len[b] : p.Lk;
  const float bias = -10000.0f * LOG2E;               // the reference's additive mask value, in log2 units
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < p.Lk; k0 += SK) {
    const int nk = min(SK, p.Lk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < SK * D; i += SQ * 4) {
      const int j = i / D, d = i % D;
      float kv = 0.f, vv = 0.f;
      if (j < nk) {
        kv = ld_hilo(K, Kl, static_cast<size_t>(k0 + j) * p.ld_k + d);
        vv = ld_hilo(V, Vl, static_cast<size_t>(k0 + j) * p.ld_v + d);
      }
      sk[j][d] = kv;
      sv[j][d] = vv;
    }
    __syncthreads();
    for (int c0 = 0; c0 < nk; c0 += SCH) {
      const int nc = min(SCH, nk - c0);
      float s[SCH];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < SCH; ++j) {
        float d = 0.f;
        if (j < nc) {
#pragma unroll
          for (int i = 0; i < DP; ++i) d = fmaf(q[i], sk[c0 + j][part * DP + i], d);
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if (k0 + c0 + j >= kvl) d += bias;
        s[j] = j < nc ? d : -INFINITY;
        cmax = fmaxf(cmax, s[j]);
      }
      const float mn = fmaxf(m, cmax);
      const float corr = exp2f(m - mn);               // m = -inf on the first chunk: 0
      l *= corr;
#pragma unroll
      for (int i = 0; i < DP; ++i) acc[i] *= corr;
#pragma unroll
      for (int j = 0; j < SCH; ++j) {
        if (j < nc) {
          const float pj = exp2f(s[j] - mn);
          l += pj;
#pragma unroll
          for (int i = 0; i < DP; ++i) acc[i] = fmaf(pj, sv[c0 + j][part * DP + i], acc[i]);
        }
      }
      m = mn;
    }
  }
  if (!qvalid) return;
  const float inv = 1.0f / l;
  __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(p.o) + static_cast<size_t>(b) * p.bs_o + static_cast<size_t>(qi) * p.ld_o + h * D + part * DP;
  __nv_bfloat16* Ol = reinterpret_cast<__nv_bfloat16*>(p.o_lo) + static_cast<size_t>(b) * p.bs_o + static_cast<size_t>(qi) * p.ld_o + h * D + part * DP;
#pragma unroll
  for (int i = 0; i < DP; i += 2) {
    const float y0 = acc[i] * inv, y1 = acc[i + 1] * inv;
    const uint32_t hi = pack_bf16(y0, y1);
    *reinterpret_cast<uint32_t*>(O + i) = hi;
    *reinterpret_cast<uint32_t*>(Ol + i) = pack_bf16(y0 - bf16_lo(hi), y1 - bf16_hi(hi));
  }
}

}  // namespace

cudaError_t launch_attention_split(const AttnParams& p, cudaStream_t st) {
  if (p.Lk < 1 || p.Lq < 1 || !p.q_lo || !p.k_lo || !p.v_lo || !p.o_lo) return cudaErrorInvalidValue;
  dim3 grid((p.Lq + SQ - 1) / SQ, p.H, p.B);
  switch (p.D) {
    case 16: return launch_k(attention_split_kernel<16>, grid, dim3(SQ * 4), 0, st, p);
    case 32: return launch_k(attention_split_kernel<32>, grid, dim3(SQ * 4), 0, st, p);
    case 48: return launch_k(attention_split_kernel<48>, grid, dim3(SQ * 4), 0, st, p);
    case 64: return launch_k(attention_split_kernel<64>, grid, dim3(SQ * 4), 0, st, p);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace DV_NS

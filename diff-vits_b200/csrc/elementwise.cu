// Small / HBM-bound kernels of the denoising loop:
//  * layout change at the API boundary ([B, C, T] fp32 -> channel-last bf16 GEMM operand),
//  * timestep embedding (sinusoid -> MLP -> + text embedding -> SiLU -> all 22 time_emb_proj at once)
//    (reference: unet1d/embeddings.py:24-64,157-201; unet_1d_condition.py:841-918; resnet.py:614-617),
//  * the samplers' per-step multistep updates as single coalesced float4 passes
//    (reference: sampler/dpm_solver.py:547-580, 796-852; sampler/uni_pc.py:471-588).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {
namespace {

// ------------------------------------------------------------------ layout
__global__ void __launch_bounds__(256) nct_to_ntc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         int Csrc, int T, int ld_dst, int col0, int split, int shift,
                                                         __nv_bfloat16* __restrict__ dst_lo) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, t = t0 + tx;
    tile[k][tx] = (c < Csrc && t < T) ? src[(static_cast<size_t>(b) * Csrc + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int t = t0 + k, c = c0 + tx;
    if (t < T && c < Csrc) {
      const int col = col0 + c + (c >= split ? shift : 0);
      const __nv_bfloat16 hi = h16_from_f32(tile[tx][k]);
      dst[(static_cast<size_t>(b) * T + t) * ld_dst + col] = hi;
      if (dst_lo) dst_lo[(static_cast<size_t>(b) * T + t) * ld_dst + col] = h16_from_f32(tile[tx][k] - h16_to_f32(hi));
    }
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                          size_t n, __nv_bfloat16* __restrict__ dst_lo) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i4 = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i4);
    uint2 w;
    w.x = pack_bf16(v.x, v.y);
    w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i4) = w;
    if (dst_lo) {
      uint2 wl;
      wl.x = pack_bf16(v.x - bf16_lo(w.x), v.y - bf16_hi(w.x));
      wl.y = pack_bf16(v.z - bf16_lo(w.y), v.w - bf16_hi(w.y));
      *reinterpret_cast<uint2*>(dst_lo + i4) = wl;
    }
  } else {
    for (size_t i = i4; i < n; ++i) {
      const __nv_bfloat16 hi = h16_from_f32(src[i]);
      dst[i] = hi;
      if (dst_lo) dst_lo[i] = h16_from_f32(src[i] - h16_to_f32(hi));
    }
  }
}

// ------------------------------------------------------------------ time embedding
// temb_mlp: e = [cos(t f), sin(t f)] ; h = SiLU(W1 e + b1) ; emb = W2 h + b2   (embeddings.py:24-64, 192-197)
// grid (16, nb): blockIdx.x owns 32 of the 512 outputs of the second Linear (the first Linear is recomputed by every
// CTA: 64 K MACs), nb = 1 when all batch items share t (the sampling loop) else B. One warp per output row, lanes
// stride the reduction axis so weight rows are read with 128-byte coalesced loads.
__global__ void __launch_bounds__(512) temb_mlp_kernel(const TembParams p) {
  pdl_launch_dependents();
  __shared__ float e[256];
  __shared__ float h[512];
  const int bi = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  const float t = p.t[static_cast<size_t>(bi) * p.t_stride];
  const int half = p.sin_dim / 2;
  if (threadIdx.x < half) {
    const float a = t * p.freqs[threadIdx.x];
    e[threadIdx.x] = cosf(a);          // flip_sin_to_cos=True: cos half first
    e[half + threadIdx.x] = sinf(a);
  }
  __syncthreads();
  for (int o = warp; o < p.emb_dim; o += 16) {
    float acc = 0.f;
    for (int k = lane; k < p.sin_dim; k += 32) acc += __ldg(p.w1 + static_cast<size_t>(o) * p.sin_dim + k) * e[k];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) h[o] = silu_f(acc + p.b1[o]);
  }
  __syncthreads();
  for (int oo = warp; oo < 32; oo += 16) {
    const int o = blockIdx.x * 32 + oo;
    float acc = 0.f;
#pragma unroll 4
    for (int k = lane; k < p.emb_dim; k += 32) acc += __ldg(p.w2 + static_cast<size_t>(o) * p.emb_dim + k) * h[k];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) p.emb_base[static_cast<size_t>(bi) * p.emb_dim + o] = acc + p.b2[o];
  }
}

// ss_all[b, j] = b_all[j] + sum_k w_all_t[k, j] * SiLU(emb_base[b] + aug[b])[k]       (resnet.py:614-617, all 22 at once)
// CTA = 32 output columns j x 16 batch items: thread (j, 2 batch items); the transposed weight matrix makes the
// per-k loads coalesced across the warp's 32 columns; activations are broadcast from shared memory.
constexpr int TP_BCHUNK = 16;
__global__ void __launch_bounds__(256) temb_proj_kernel(const TembParams p) {
  pdl_launch_dependents();
  __shared__ float es[TP_BCHUNK][512];
  const int b0 = blockIdx.y * TP_BCHUNK;
  const int nb = min(TP_BCHUNK, p.B - b0);
  pdl_wait();
  for (int i = threadIdx.x; i < TP_BCHUNK * 512; i += 256) {
    const int bb = i >> 9, k = i & 511;
    float v = 0.f;
    if (bb < nb) {
      v = p.emb_base[static_cast<size_t>(p.t_stride ? (b0 + bb) : 0) * 512 + k];
      if (p.aug) v += p.aug[static_cast<size_t>(b0 + bb) * 512 + k];
      v = silu_f(v);
    }
    es[bb][k] = v;
  }
  __syncthreads();
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int bq = (threadIdx.x >> 5) * 2;
  if (j >= p.n_all) return;
  float a0 = 0.f, a1 = 0.f;
  const float* w = p.w_all + j;
#pragma unroll 8
  for (int k = 0; k < 512; ++k) {
    const float wv = __ldg(w + static_cast<size_t>(k) * p.n_all);
    a0 = fmaf(wv, es[bq][k], a0);
    a1 = fmaf(wv, es[bq + 1][k], a1);
  }
  const float bias = __ldg(p.b_all + j);
  if (bq < nb) p.ss_all[static_cast<size_t>(b0 + bq) * p.n_all + j] = a0 + bias;
  if (bq + 1 < nb) p.ss_all[static_cast<size_t>(b0 + bq + 1) * p.n_all + j] = a1 + bias;
}

// ------------------------------------------------------------------ sampler updates (fp32, float4)
__device__ __forceinline__ float4 ld4(const float* p, size_t i) { return *reinterpret_cast<const float4*>(p + i); }
__device__ __forceinline__ void st4(float* p, size_t i, float4 v) { *reinterpret_cast<float4*>(p + i) = v; }

__device__ __forceinline__ float roundtrip(float x, float m, float alp, float sig) {
  // sampler/dpm_solver.py:290-292 then :437-439 : x0 -> eps -> x0, evaluated in the reference's order
  const float eps = __fdiv_rn(__fsub_rn(x, __fmul_rn(alp, m)), sig);
  return __fdiv_rn(__fsub_rn(x, __fmul_rn(sig, eps)), alp);
}


// ---- sampler diagnostics (replaces the reference's syncing `assert torch.isnan(x).any() == False`, model3.py:903):
// every update kernel counts the non-finite values among the model output it consumes and the state it writes, and
// takes their sums of squares: per-thread accumulators -> warp shuffle tree -> the 8 warps of the CTA through shared memory
// (fixed order, no atomics) -> ONE 16-byte store per CTA into this launch's row of partials; diag_reduce_kernel folds the
// rows once, after the loop. part == nullptr: diagnostics off (the stand-alone C-ABI entry points).
struct DiagAcc { float bad = 0.f, sm = 0.f, sx = 0.f; };
__device__ __forceinline__ float nonfinite(float v) { return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u ? 1.f : 0.f; }
__device__ __forceinline__ void diag_add(DiagAcc& d, float m, float r) {
  d.bad += nonfinite(m) + nonfinite(r);
  d.sm = fmaf(m, m, d.sm);
  d.sx = fmaf(r, r, d.sx);
}
__device__ __forceinline__ void diag_add4(DiagAcc& d, const float4& m, const float4& r) {
  diag_add(d, m.x, r.x); diag_add(d, m.y, r.y); diag_add(d, m.z, r.z); diag_add(d, m.w, r.w);
}
__device__ __forceinline__ void diag_store(float4* part, DiagAcc d, float kind) {   // all 256 threads of the CTA
  __shared__ float sh[8][3];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    d.bad += __shfl_xor_sync(0xffffffffu, d.bad, o);
    d.sm += __shfl_xor_sync(0xffffffffu, d.sm, o);
    d.sx += __shfl_xor_sync(0xffffffffu, d.sx, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = d.bad; sh[threadIdx.x >> 5][1] = d.sm; sh[threadIdx.x >> 5][2] = d.sx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f, m = 0.f, x = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { b += sh[w][0]; m += sh[w][1]; x += sh[w][2]; }
    part[blockIdx.x] = make_float4(b, m, x, kind);
  }
}
// rows[j] = fold of the `per_row` CTA partials of update launch j: {non-finite count, sum m^2, sum x^2, kernel kind}
__global__ void __launch_bounds__(256) diag_reduce_kernel(const float4* __restrict__ part, float4* __restrict__ rows, int per_row) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sh[8][3];
  const float4* p = part + static_cast<size_t>(blockIdx.x) * per_row;
  DiagAcc d;
  for (int i = threadIdx.x; i < per_row; i += 256) { const float4 v = p[i]; d.bad += v.x; d.sm += v.y; d.sx += v.z; }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    d.bad += __shfl_xor_sync(0xffffffffu, d.bad, o);
    d.sm += __shfl_xor_sync(0xffffffffu, d.sm, o);
    d.sx += __shfl_xor_sync(0xffffffffu, d.sx, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = d.bad; sh[threadIdx.x >> 5][1] = d.sm; sh[threadIdx.x >> 5][2] = d.sx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f, m = 0.f, x = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { b += sh[w][0]; m += sh[w][1]; x += sh[w][2]; }
    rows[blockIdx.x] = make_float4(b, m, x, p[0].w);
  }
}

__global__ void __launch_bounds__(256) axpbypcz_kernel(float* xo, const float* x, float* m1,
                                                       const float* __restrict__ m0, float a, float b, float c,
                                                       int rt, float alp, float sig, size_t n, float4* diag) {
  pdl_launch_dependents();
  pdl_wait();
  DiagAcc dg;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 xv = ld4(x, i);
      float4 mv = ld4(m1, i);
      if (rt) {
        mv.x = roundtrip(xv.x, mv.x, alp, sig); mv.y = roundtrip(xv.y, mv.y, alp, sig);
        mv.z = roundtrip(xv.z, mv.z, alp, sig); mv.w = roundtrip(xv.w, mv.w, alp, sig);
        st4(m1, i, mv);
      }
      float4 r;
      // reference order: (sigma_t/sigma_s) * x - alpha_t*phi * m1 - 0.5*alpha_t*phi/r0 * (m1 - m0)
      r.x = a * xv.x + b * mv.x; r.y = a * xv.y + b * mv.y; r.z = a * xv.z + b * mv.z; r.w = a * xv.w + b * mv.w;
      if (m0) {
        const float4 pv = ld4(m0, i);
        r.x += c * pv.x; r.y += c * pv.y; r.z += c * pv.z; r.w += c * pv.w;
      }
      st4(xo, i, r);
      diag_add4(dg, mv, r);
    } else {
      for (size_t k = i; k < n; ++k) {
        float mv = m1[k];
        if (rt) { mv = roundtrip(x[k], mv, alp, sig); m1[k] = mv; }
        float r = a * x[k] + b * mv;
        if (m0) r += c * m0[k];
        xo[k] = r;
        diag_add(dg, mv, r);
      }
    }
  }
  if (diag) diag_store(diag, dg, 0.f);
}

__global__ void __launch_bounds__(256) unipc_pred_kernel(float* xp, float* __restrict__ base,
                                                         const float* x, const float* __restrict__ m1,
                                                         const float* __restrict__ m0, float a, float b, float cp,
                                                         float cc, int has_m0, size_t n, float4* diag) {
  pdl_launch_dependents();
  pdl_wait();
  DiagAcc dg;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 xv = ld4(x, i), mv = ld4(m1, i);
      float4 xb;
      xb.x = a * xv.x + b * mv.x; xb.y = a * xv.y + b * mv.y; xb.z = a * xv.z + b * mv.z; xb.w = a * xv.w + b * mv.w;
      float4 p = xb, q = xb;
      if (has_m0) {
        const float4 pv = ld4(m0, i);
        const float dx = pv.x - mv.x, dy = pv.y - mv.y, dz = pv.z - mv.z, dw = pv.w - mv.w;
        p.x += cp * dx; p.y += cp * dy; p.z += cp * dz; p.w += cp * dw;
        q.x += cc * dx; q.y += cc * dy; q.z += cc * dz; q.w += cc * dw;
      }
      st4(xp, i, p);
      if (base) st4(base, i, q);
      diag_add4(dg, mv, p);
    } else {
      for (size_t k = i; k < n; ++k) {
        const float xb = a * x[k] + b * m1[k];
        const float d = has_m0 ? (m0[k] - m1[k]) : 0.f;
        xp[k] = xb + cp * d;
        if (base) base[k] = xb + cc * d;
        diag_add(dg, m1[k], xb + cp * d);
      }
    }
  }
  if (diag) diag_store(diag, dg, 1.f);
}

__global__ void __launch_bounds__(256) unipc_corr_kernel(float* __restrict__ x, const float* __restrict__ base,
                                                         const float* __restrict__ mt, const float* __restrict__ m1,
                                                         float d, size_t n, float4* diag) {
  pdl_launch_dependents();
  pdl_wait();
  DiagAcc dg;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 bv = ld4(base, i), tv = ld4(mt, i), mv = ld4(m1, i);
      float4 r;
      r.x = bv.x + d * (tv.x - mv.x); r.y = bv.y + d * (tv.y - mv.y);
      r.z = bv.z + d * (tv.z - mv.z); r.w = bv.w + d * (tv.w - mv.w);
      st4(x, i, r);
      diag_add4(dg, tv, r);
    } else {
      for (size_t k = i; k < n; ++k) {
        const float r = base[k] + d * (mt[k] - m1[k]);
        x[k] = r;
        diag_add(dg, mt[k], r);
      }
    }
  }
  if (diag) diag_store(diag, dg, 2.f);
}

inline int ew_grid(size_t n) {
  const size_t blocks = (n / 4 + 255) / 256;
  const size_t cap = 148 * 8;  // 8 resident CTAs of 256 threads per SM
  return static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DV_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

cudaError_t launch_nct_to_ntc(const float* src, void* dst, int B, int Csrc, int T, int ld_dst, int col0, int split,
                              int shift, cudaStream_t st, void* dst_lo) {
  dim3 grid((T + 31) / 32, (Csrc + 31) / 32, B);
  return launch_k(nct_to_ntc_kernel, grid, dim3(256), 0, st, src, reinterpret_cast<__nv_bfloat16*>(dst), Csrc, T, ld_dst, col0, split, shift,
                  reinterpret_cast<__nv_bfloat16*>(dst_lo));
}
cudaError_t launch_f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t st, void* dst_lo) {
  const size_t blocks = (n / 4 + 256) / 256;
  return launch_k(f32_to_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, src, reinterpret_cast<__nv_bfloat16*>(dst), n,
                  reinterpret_cast<__nv_bfloat16*>(dst_lo));
}
cudaError_t launch_temb(const TembParams& p, cudaStream_t st) {
  if (p.emb_dim != 512 || p.sin_dim > 256 || p.sin_dim % 2) return cudaErrorInvalidValue;
  cudaError_t e = launch_k(temb_mlp_kernel, dim3(16, p.t_stride ? p.B : 1), dim3(512), 0, st, p);
  if (e != cudaSuccess) return e;
  dim3 grid((p.n_all + 31) / 32, (p.B + TP_BCHUNK - 1) / TP_BCHUNK);
  return launch_k(temb_proj_kernel, grid, dim3(256), 0, st, p);
}

cudaError_t launch_axpbypcz(float* x_out, const float* x, float* m1, const float* m0, float a, float b, float c,
                            int roundtrip_flag, float alp, float sig, size_t n, cudaStream_t st, float* diag) {
  return launch_k(axpbypcz_kernel, dim3(ew_grid(n)), dim3(256), 0, st, x_out, x, m1, m0, a, b, c, roundtrip_flag, alp, sig, n,
                  reinterpret_cast<float4*>(diag));
}
cudaError_t launch_unipc_pred(float* xpred, float* base, const float* x, const float* m1, const float* m0, float a,
                              float b, float cp, float cc, int has_m0, size_t n, cudaStream_t st, float* diag) {
  return launch_k(unipc_pred_kernel, dim3(ew_grid(n)), dim3(256), 0, st, xpred, base, x, m1, m0, a, b, cp, cc, has_m0, n,
                  reinterpret_cast<float4*>(diag));
}
cudaError_t launch_unipc_corr(float* x, const float* base, const float* mt, const float* m1, float d, size_t n,
                              cudaStream_t st, float* diag) {
  return launch_k(unipc_corr_kernel, dim3(ew_grid(n)), dim3(256), 0, st, x, base, mt, m1, d, n, reinterpret_cast<float4*>(diag));
}
int sampler_diag_ctas(size_t n) { return ew_grid(n); }
cudaError_t launch_diag_reduce(const float* part, float* rows, int nrows, int per_row, cudaStream_t st) {
  return launch_k(diag_reduce_kernel, dim3(nrows), dim3(256), 0, st, reinterpret_cast<const float4*>(part), reinterpret_cast<float4*>(rows), per_row);
}

}  // namespace DV_NS

// GroupNorm (+SiLU, + time scale/shift) and LayerNorm on channel-last bf16 activations. HBM/L2-bound kernels:
// 16-byte vector loads along the contiguous channel axis, fp32 math, statistics combined in fp64.
//
// GroupNorm statistics arrive as partial sums per (32-row chunk, 16-channel block), normally written by the epilogue
// of the GEMM that produced the tensor (csrc/gemm_tc.cu), so the tensor is not re-read; gn_stats below produces the
// same layout for tensors that do not come from a GEMM (tests, stand-alone use). A group is any union of blocks,
// also across the two sources of torch.cat([h, skip], 1) in the up blocks (never materialised; groups may straddle
// the seam, e.g. 896 = 512 + 384 channels in 8 groups of 112). All reductions are fixed-order: bit reproducible and
// independent of the batch size.
//   gn_apply : y = SiLU?( ((x-mean)*rstd*gamma+beta) * (1+scale[b,c]) + shift[b,c] ) -> bf16 GEMM operand.
// Reference: unet1d/resnet.py:591-631 (norm1/norm2 + scale_shift), transformer_1d.py:256 (eps 1e-6),
// unet_1d_condition.py:1029-1031 (conv_norm_out + SiLU).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ const uint4* gn_src(const GnParams& p, int b, int l, int c, bool lo = false) {
  if (c < p.c0)
    return reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(lo ? p.x0_lo : p.x0) +
                                          (static_cast<size_t>(b) * p.L + l) * p.c0 + c);
  return reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(lo ? p.x1_lo : p.x1) +
                                        (static_cast<size_t>(b) * p.L + l) * p.c1 + (c - p.c0));
}

// One CTA per (32-row chunk, batch item): partial sums per 16-channel block.
__global__ void __launch_bounds__(NT) gn_stats_kernel(const GnParams p, float* __restrict__ scratch) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float ps[NT][2];
  const int C = p.c0 + p.c1;
  const int octs = C >> 3;
  const int R = NT / octs;
  const int b = blockIdx.y;
  const int lbeg = blockIdx.x * 32;
  const int lend = min(lbeg + 32, p.L);
  const int r = threadIdx.x / octs, o = threadIdx.x % octs;
  float s = 0.f, ss = 0.f;
  if (r < R) {
    for (int l = lbeg + r; l < lend; l += R) {
      const uint4 v = __ldg(gn_src(p, b, l, o * 8));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(w[j]), c = bf16_hi(w[j]);
        s += a + c;
        ss += a * a + c * c;
      }
    }
  }
  ps[threadIdx.x][0] = s;
  ps[threadIdx.x][1] = ss;
  __syncthreads();
  if (threadIdx.x < (octs >> 1)) {   // fold the R row slots and the two octets of a 16-channel block, fixed order
    float a = 0.f, c = 0.f;
    for (int rr = 0; rr < R; ++rr)
      for (int h = 0; h < 2; ++h) {
        a += ps[rr * octs + threadIdx.x * 2 + h][0];
        c += ps[rr * octs + threadIdx.x * 2 + h][1];
      }
    float* dst = scratch + ((static_cast<size_t>(b) * gridDim.x + blockIdx.x) * (octs >> 1) + threadIdx.x) * 2;
    dst[0] = a;
    dst[1] = c;
  }
}

// SPLIT: FP32-class mode - the input is hi + lo, the result is written as hi and lo planes.
template <bool SPLIT>
__global__ void __launch_bounds__(NT) gn_apply_kernel(const GnParams p, int rows_per_cta) {
  pdl_launch_dependents();
  __shared__ float s_mean[32], s_rstd[32];
  const int C = p.c0 + p.c1;
  const int cpg = C / p.G;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int octs = C >> 3;
  const int R = NT / octs;
  const int r = threadIdx.x / octs, o = threadIdx.x % octs;
  const int cbase = o * 8;
  // affine parameters are static data: fetch them before waiting on the producer kernel
  float ga[8], be[8];
  if (r < R) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + cbase + j));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + cbase + j));
      ga[j] = g4.x; ga[j + 1] = g4.y; ga[j + 2] = g4.z; ga[j + 3] = g4.w;
      be[j] = b4.x; be[j + 1] = b4.y; be[j + 2] = b4.z; be[j + 3] = b4.w;
    }
  }
  pdl_wait();
  // first batch of rows: issued before the statistics are reduced, so the two global round trips overlap
  constexpr int UNR = 4;                                     // rows in flight per thread
  const int lbeg = blockIdx.x * rows_per_cta;
  const int lend = min(lbeg + rows_per_cta, p.L);
  uint4 v[UNR];
  uint4 vl[SPLIT ? UNR : 1];
  if (r < R) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int l = lbeg + r + u * R;
      v[u] = l < lend ? __ldg(gn_src(p, b, l, cbase)) : make_uint4(0, 0, 0, 0);
      if (SPLIT) vl[u] = l < lend ? __ldg(gn_src(p, b, l, cbase, true)) : make_uint4(0, 0, 0, 0);
    }
  }
  // ---- group statistics: one warp per group; the (block, chunk, fold) index space is flattened over the lanes so
  // every load of the warp is independent; fixed lane assignment + fixed shuffle tree: deterministic.
  for (int g = warp; g < p.G; g += NT / 32) {
    double sum = 0.0, sumsq = 0.0;
    const int kb0 = (g * cpg) >> 4, nkb = cpg >> 4;
    // all blocks of a group come from at most two sources; walk them source by source
    int cov0 = 0;
    for (int si = 0; si < p.nst; ++si) {
      const GnStatSrc st = p.st[si];
      const int lo = max(kb0, cov0 >> 4), hi = min(kb0 + nkb, (cov0 + st.cover) >> 4);
      if (hi > lo) {
        const int nf = st.fold ? 2 : 1;
        const int per_blk = st.chunks * nf;
        const int items = (hi - lo) * per_blk;
        const float* base = st.ptr + static_cast<size_t>(b) * st.chunks * st.nblk * 2;
        for (int i0 = 0; i0 < items; i0 += 128) {           // 4 independent loads per lane in flight
          float2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            v[u] = make_float2(0.f, 0.f);
            if (i < items) {
              const int kl = lo - (cov0 >> 4) + i / per_blk;
              const int rem = i % per_blk;
              const int ch = rem / nf, f = rem % nf;
              v[u] = __ldg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(ch) * st.nblk + kl + f * (st.nblk >> 1)) * 2));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) { sum += static_cast<double>(v[u].x); sumsq += static_cast<double>(v[u].y); }
        }
      }
      cov0 += st.cover;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, m);
      sumsq += __shfl_xor_sync(0xffffffffu, sumsq, m);
    }
    if (lane == 0) {
      const double n = static_cast<double>(p.L) * cpg;
      const double mean = sum / n;
      double var = sumsq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = static_cast<float>(mean);
      s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    }
  }
  __syncthreads();
  if (r >= R) return;
  const int grp = cbase / cpg;
  const float mean = s_mean[grp], rstd = s_rstd[grp];
  float ca[8], cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = rstd * ga[j];
    float bb = be[j] - mean * a;
    ca[j] = a;
    cb[j] = bb;
  }
  if (p.scale_shift) {
    const float* ssp = p.scale_shift + static_cast<size_t>(b) * p.ss_ld + cbase;
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(ssp + j));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(ssp + C + j));
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float s1 = 1.0f + scv[k];
        ca[j + k] *= s1;
        cb[j + k] = cb[j + k] * s1 + shv[k];
      }
    }
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  for (int l0 = lbeg + r; l0 < lend; l0 += UNR * R) {
    if (l0 != lbeg + r) {                                    // later batches (rows_per_cta > UNR * R only)
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int l = l0 + u * R;
        v[u] = l < lend ? __ldg(gn_src(p, b, l, cbase)) : make_uint4(0, 0, 0, 0);
        if (SPLIT) vl[u] = l < lend ? __ldg(gn_src(p, b, l, cbase, true)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int l = l0 + u * R;
      if (l >= lend) continue;
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      uint32_t y[4], yl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float xa = bf16_lo(w[j]), xc = bf16_hi(w[j]);
        if (SPLIT) {
          const uint32_t wl = j == 0 ? vl[u].x : (j == 1 ? vl[u].y : (j == 2 ? vl[u].z : vl[u].w));
          xa += bf16_lo(wl);
          xc += bf16_hi(wl);
        }
        float a = fmaf(xa, ca[2 * j], cb[2 * j]);
        float c = fmaf(xc, ca[2 * j + 1], cb[2 * j + 1]);
        if (p.silu) { a = silu_f(a); c = silu_f(c); }
        y[j] = pack_bf16(a, c);
        if (SPLIT) yl[j] = pack_bf16(a - bf16_lo(y[j]), c - bf16_hi(y[j]));
      }
      *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * p.L + l) * C + cbase) = make_uint4(y[0], y[1], y[2], y[3]);
      if (SPLIT && p.out_lo != nullptr)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_lo) + (static_cast<size_t>(b) * p.L + l) * C + cbase) =
            make_uint4(yl[0], yl[1], yl[2], yl[3]);
    }
  }
}

// One warp per row, row kept in registers (C <= 1024, multiple of 128): exact two-pass mean / variance. Each warp
// works on LN_ROWS rows at a time with all of their loads issued up front (memory-level parallelism: at C = 128 a
// row is only 8 bytes per lane).
constexpr int LN_WARPS = 8;
template <int CHUNKS, bool SPLIT>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int rows, float eps, const __nv_bfloat16* __restrict__ x_lo,
                                                        __nv_bfloat16* __restrict__ out_lo) {
  constexpr int C = CHUNKS * 128;
  constexpr int LN_ROWS = CHUNKS <= 2 ? 4 : 2;
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  float4 gm[CHUNKS], bt[CHUNKS];
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma + (lane + 32 * i) * 4));
    bt[i] = __ldg(reinterpret_cast<const float4*>(beta + (lane + 32 * i) * 4));
  }
  pdl_wait();
  const int row0 = (blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * LN_ROWS;
  if (row0 >= rows) return;
  uint2 w[LN_ROWS][CHUNKS];
  uint2 wl[SPLIT ? LN_ROWS : 1][SPLIT ? CHUNKS : 1];
#pragma unroll
  for (int rr = 0; rr < LN_ROWS; ++rr) {
    const int row = min(row0 + rr, rows - 1);
    const uint2* xr = reinterpret_cast<const uint2*>(x + static_cast<size_t>(row) * C);
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) w[rr][i] = __ldg(xr + lane + 32 * i);
    if (SPLIT) {
      const uint2* xl = reinterpret_cast<const uint2*>(x_lo + static_cast<size_t>(row) * C);
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) wl[rr][i] = __ldg(xl + lane + 32 * i);
    }
  }
#pragma unroll
  for (int rr = 0; rr < LN_ROWS; ++rr) {
    const int row = row0 + rr;
    if (row >= rows) continue;
    float v[CHUNKS][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
      v[i][0] = bf16_lo(w[rr][i].x); v[i][1] = bf16_hi(w[rr][i].x); v[i][2] = bf16_lo(w[rr][i].y); v[i][3] = bf16_hi(w[rr][i].y);
      if (SPLIT) {
        v[i][0] += bf16_lo(wl[rr][i].x); v[i][1] += bf16_hi(wl[rr][i].x);
        v[i][2] += bf16_lo(wl[rr][i].y); v[i][3] += bf16_hi(wl[rr][i].y);
      }
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = v[i][j] - mean; q += d * d; }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) q += __shfl_xor_sync(0xffffffffu, q, m);
    const float rstd = rsqrtf(q * (1.0f / C) + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * C);
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
      const float y0 = (v[i][0] - mean) * rstd * gm[i].x + bt[i].x, y1 = (v[i][1] - mean) * rstd * gm[i].y + bt[i].y;
      const float y2 = (v[i][2] - mean) * rstd * gm[i].z + bt[i].z, y3 = (v[i][3] - mean) * rstd * gm[i].w + bt[i].w;
      uint2 o;
      o.x = pack_bf16(y0, y1);
      o.y = pack_bf16(y2, y3);
      orow[lane + 32 * i] = o;
      if (SPLIT && out_lo != nullptr) {
        uint2 ol;
        ol.x = pack_bf16(y0 - bf16_lo(o.x), y1 - bf16_hi(o.x));
        ol.y = pack_bf16(y2 - bf16_lo(o.y), y3 - bf16_hi(o.y));
        reinterpret_cast<uint2*>(out_lo + static_cast<size_t>(row) * C)[lane + 32 * i] = ol;
      }
    }
  }
}

// LayerNorm over the first c_real of 128 stored columns with row masks (prompt encoder: LN(100) in front of `pre`, the
// final LayerNorm with zeroed padding rows). One warp per row.
__global__ void __launch_bounds__(256) layernorm_masked_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               int rows, float eps, LnMaskParams m) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  float g[4], bt[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane * 4 + j;
    g[j] = c < m.c_real ? __ldg(gamma + c) : 0.f;
    bt[j] = c < m.c_real ? __ldg(beta + c) : 0.f;
  }
  pdl_wait();
  if (row >= rows) return;
  const bool padded = m.row_len != nullptr && (row % m.rows_per_batch) >= __ldg(m.row_len + row / m.rows_per_batch);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (!(padded && m.mask_in)) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(x + static_cast<size_t>(row) * 128) + lane);
    v[0] = bf16_lo(w.x); v[1] = bf16_hi(w.x); v[2] = bf16_lo(w.y); v[3] = bf16_hi(w.y);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { if (lane * 4 + j >= m.c_real) v[j] = 0.f; s += v[j]; }
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  const float mean = s / m.c_real;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float d = lane * 4 + j < m.c_real ? v[j] - mean : 0.f; q += d * d; }
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) q += __shfl_xor_sync(0xffffffffu, q, k);
  const float rstd = rsqrtf(q / m.c_real + eps);
  float y[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    y[j] = lane * 4 + j < m.c_real ? (v[j] - mean) * rstd * g[j] + bt[j] : 0.f;
    if (padded && m.mask_out) y[j] = 0.f;
  }
  uint2 o;
  o.x = pack_bf16(y[0], y[1]);
  o.y = pack_bf16(y[2], y[3]);
  reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * 128)[lane] = o;
  if (m.out_f32) reinterpret_cast<float4*>(m.out_f32 + static_cast<size_t>(row) * 128)[lane] = make_float4(y[0], y[1], y[2], y[3]);
}

}  // namespace

cudaError_t launch_layernorm_masked(const void* x, void* out, const float* gamma, const float* beta, int rows, float eps,
                                    const LnMaskParams& m, cudaStream_t st) {
  if (m.c_real < 1 || m.c_real > 128 || rows < 1 || (m.row_len && m.rows_per_batch < 1)) return cudaErrorInvalidValue;
  return launch_k(layernorm_masked_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x),
                  reinterpret_cast<__nv_bfloat16*>(out), gamma, beta, rows, eps, m);
}

static bool gn_ok(const GnParams& p) {
  const int C = p.c0 + p.c1;
  return p.G > 0 && p.G <= 32 && C % (16 * p.G) == 0 && p.c0 % 16 == 0 && p.c1 % 16 == 0 && C / 8 <= NT;
}
size_t gn_scratch_floats(int B, int L, int C) {
  return static_cast<size_t>(B) * ((L + 31) / 32) * (C / 16) * 2;
}
cudaError_t launch_gn_stats(const GnParams& p, float* scratch, cudaStream_t st) {
  if (!gn_ok(p)) return cudaErrorInvalidValue;
  dim3 grid((p.L + 31) / 32, p.B);
  return launch_k(gn_stats_kernel, grid, dim3(NT), 0, st, p, scratch);
}
cudaError_t launch_gn_apply(const GnParams& p, cudaStream_t st) {
  if (!gn_ok(p) || p.nst < 1 || p.nst > 2) return cudaErrorInvalidValue;
  // four rows in flight per thread: R = 256 / (C/8) rows per pass -> 4R rows per CTA (64 at C = 128, 8 at C = 1024)
  const int C = p.c0 + p.c1;
  // (measured on the C2 step, ms per call: 2 rows in flight 4.36, 3 -> 4.26, 4 -> 4.21, 8 -> 4.24)
  int rows = 4 * (NT / (C >> 3));
  if (rows < 4) rows = 4;
  dim3 grid((p.L + rows - 1) / rows, p.B);
  // SPLIT kernel: the input is hi + lo (FP32X3: every tensor; FP16S: residual-stream tensors); the lo plane of the output is
  // written only where the consumer is not the tensor core alone (out_lo != nullptr)
  if (p.x0_lo != nullptr) {
    if (p.c1 && !p.x1_lo) return cudaErrorInvalidValue;
    return launch_k(gn_apply_kernel<true>, grid, dim3(NT), 0, st, p, rows);
  }
  if (p.out_lo != nullptr) return cudaErrorInvalidValue;
  return launch_k(gn_apply_kernel<false>, grid, dim3(NT), 0, st, p, rows);
}
cudaError_t launch_layernorm(const void* x, void* out, const float* gamma, const float* beta, int rows, int C,
                             float eps, cudaStream_t st, const void* x_lo, void* out_lo) {
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* oo = reinterpret_cast<__nv_bfloat16*>(out);
  const __nv_bfloat16* xl = reinterpret_cast<const __nv_bfloat16*>(x_lo);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  const int rpc = LN_WARPS * (C <= 256 ? 4 : 2);   // rows per CTA, matches LN_ROWS in the kernel
  dim3 grid((rows + rpc - 1) / rpc);
  if (x_lo == nullptr && out_lo != nullptr) return cudaErrorInvalidValue;
#define DV_LN_CASE(CH)                                                                                                   \
  case CH * 128:                                                                                                         \
    return x_lo ? launch_k(layernorm_kernel<CH, true>, grid, dim3(256), 0, st, xi, oo, gamma, beta, rows, eps, xl, ol) \
                  : launch_k(layernorm_kernel<CH, false>, grid, dim3(256), 0, st, xi, oo, gamma, beta, rows, eps, xl, ol);
  switch (C) {
    DV_LN_CASE(1)
    DV_LN_CASE(2)
    DV_LN_CASE(3)
    DV_LN_CASE(4)
    DV_LN_CASE(5)
    DV_LN_CASE(6)
    DV_LN_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef DV_LN_CASE
}

}  // namespace DV_NS

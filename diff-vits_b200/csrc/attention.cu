// Fused flash-style attention for the UNet's 8-head self- and prompt cross-attention
// (reference: unet1d/attention_processor.py:980-1052, F.scaled_dot_product_attention at :1032).
//
// Head dims on this path are 16/32/48/64 (C/8 per level). At d=16 one S tile costs 64 FLOP per exp, so the
// kernel is bound by the MUFU/FMA pipes (exp2, max, scale), not by the tensor pipe; S and P therefore stay in
// registers (mma.sync m16n8k16 bf16, fp32 accumulate) instead of a TMEM round trip, and the softmax is done on
// the accumulator fragments directly. K/V tiles are staged with cp.async (16-byte, zero-fill for tail rows),
// double buffered; rows are padded by 16 bytes so the 32-bit fragment loads and ldmatrix.trans are conflict free.
//
// Masking: self-attention has no mask (attention_mask=None in the reference, padded frames participate).
// Cross-attention uses the reference's additive bias: keys >= kv_len[b] get -10000 added to the scaled score
// (unet_1d_condition.py:816-818), reproduced literally so a fully masked row degrades the same way.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {
namespace {

constexpr int BQ = 64;    // queries per CTA (4 warps x 16 rows)
constexpr int BKV = 64;   // keys per tile
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." DV_MMA_T "." DV_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_ptr)));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnParams p) {
  constexpr int PITCH = D + 8;          // elements; +16 bytes
  constexpr int CHUNKS = D / 8;         // 16-byte chunks per row
  __shared__ __align__(16) __nv_bfloat16 sQ[BQ * PITCH];
  __shared__ __align__(16) __nv_bfloat16 sK[2][BKV * PITCH];
  __shared__ __align__(16) __nv_bfloat16 sV[2][BKV * PITCH];

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  pdl_wait();
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;

  const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(p.q) + b * p.bs_q + h * D;
  const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(p.k) + b * p.bs_k + h * D;
  const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(p.v) + b * p.bs_v + h * D;

  // ---- stage Q (group 0 together with the first K/V tile)
  for (int i = tid; i < BQ * CHUNKS; i += ATT_THREADS) {
    const int r = i / CHUNKS, c = i % CHUNKS;
    const int row = q0 + r;
    const bool ok = row < p.Lq;
    cp_async16(&sQ[r * PITCH + c * 8], Q + static_cast<size_t>(ok ? row : 0) * p.ld_q + c * 8, ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * BKV;
    for (int i = tid; i < BKV * CHUNKS; i += ATT_THREADS) {
      const int r = i / CHUNKS, c = i % CHUNKS;
      const int row = k0 + r;
      const bool ok = row < p.Lk;
      const size_t rr = static_cast<size_t>(ok ? row : 0);
      cp_async16(&sK[buf][r * PITCH + c * 8], K + rr * p.ld_k + c * 8, ok ? 16 : 0);
      cp_async16(&sV[buf][r * PITCH + c * 8], V + rr * p.ld_v + c * 8, ok ? 16 : 0);
    }
  };
  const int ntiles = (p.Lk + BKV - 1) / BKV;
  load_kv(0, 0);
  cp_async_commit();

  constexpr float LOG2E = 1.4426950408889634f;
  const float sl2 = p.scale * LOG2E;
  const int kvlen = p.kv_len ? p.kv_len[b] : p.Lk;
  const float mask_raw = -10000.0f / p.scale;          // the reference's additive -10000 bias, in raw-score units
  const int kvalid = min(kvlen, p.Lk);                 // keys below this index need no masking

  uint32_t qf[D / 16][4];
  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) {
      load_kv(tile + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (tile == 0) {
      const __nv_bfloat16* qb = &sQ[(warp * 16) * PITCH];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        qf[kk][0] = *reinterpret_cast<const uint32_t*>(&qb[g * PITCH + kk * 16 + 2 * t]);
        qf[kk][1] = *reinterpret_cast<const uint32_t*>(&qb[(g + 8) * PITCH + kk * 16 + 2 * t]);
        qf[kk][2] = *reinterpret_cast<const uint32_t*>(&qb[g * PITCH + kk * 16 + 2 * t + 8]);
        qf[kk][3] = *reinterpret_cast<const uint32_t*>(&qb[(g + 8) * PITCH + kk * 16 + 2 * t + 8]);
      }
    }

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[BKV / 8][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const __nv_bfloat16* kb = &sK[buf][(j * 8 + g) * PITCH + 2 * t];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kb + kk * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kb + kk * 16 + 8);
        mma_bf16_16816(s[j], qf[kk], b0, b1);
      }
    }

    // ---- online softmax on the raw scores (rows g and g+8 of this warp's 16). The softmax scale is folded into
    // the exponent: p = 2^(s*c - m*c), c = scale*log2(e): one FFMA + one MUFU.EX2 per element. Masking work is only
    // done for tiles that contain masked / out-of-range keys (warp-uniform test).
    if (tile * BKV + BKV > kvalid) {
      const int kbase = tile * BKV + 2 * t;
#pragma unroll
      for (int j = 0; j < BKV / 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kbase + j * 8 + e;
          float add = 0.f;
          if (key >= p.Lk) add = -INFINITY;
          else if (key >= kvlen) add = mask_raw;
          s[j][e] += add;
          s[j][2 + e] += add;
        }
      }
    }
    float mx0 = fmaxf(s[0][0], s[0][1]), mx1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int j = 1; j < BKV / 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 key < Lk
    const float c0 = ex2_approx((m0 - mn0) * sl2), c1 = ex2_approx((m1 - mn1) * sl2);
    m0 = mn0; m1 = mn1;
    const float ms0 = mn0 * sl2, ms1 = mn1 * sl2;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < D / 8; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }

    uint32_t pf[BKV / 16][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      const float p0 = ex2_approx(fmaf(s[j][0], sl2, -ms0)), p1 = ex2_approx(fmaf(s[j][1], sl2, -ms0));
      const float p2 = ex2_approx(fmaf(s[j][2], sl2, -ms1)), p3 = ex2_approx(fmaf(s[j][3], sl2, -ms1));
      l0 += p0 + p1;
      l1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }

    // ---- O += P V   (V fragments through ldmatrix.trans)
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int jd = 0; jd < D / 8; jd += 2) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, &sV[buf][krow * PITCH + (jd + (lane >> 4)) * 8]);
        mma_bf16_16816(o[jd], pf[kk], vf[0], vf[1]);
        mma_bf16_16816(o[jd + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(p.o) + b * p.bs_o + h * D;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < D / 8; ++j) {
    if (r0 < p.Lq)
      *reinterpret_cast<uint32_t*>(&O[static_cast<size_t>(r0) * p.ld_o + j * 8 + 2 * t]) =
          pack_bf16(o[j][0] * i0, o[j][1] * i0);
    if (r1 < p.Lq)
      *reinterpret_cast<uint32_t*>(&O[static_cast<size_t>(r1) * p.ld_o + j * 8 + 2 * t]) =
          pack_bf16(o[j][2] * i1, o[j][3] * i1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Second-generation softmax inner loop (same tiling, same contract). The first kernel spends 9 instructions per
// score pair (2 FFMA, 2 MUFU, 2 FADD row sum, 1 max, 1 pack, ...) and is bound by instruction issue and the MUFU
// pipe together (ncu: XU 66 %, issue 64 %; MUFU.EX2 measured 16.4 / clk / SM, tools/ubench/mufu.cu; the packed
// bf16x2 / f16x2 ex2 forms are two MUFU instructions each and buy nothing). Here:
//   * the exponent argument of a pair is one packed FFMA2 (fma.rn.f32x2);
//   * the row sum is not accumulated with FADDs: P (already in A-fragment form) is multiplied by an all-ones
//     B fragment, one extra HMMA per 16 keys, so l = sum of the bf16-rounded probabilities the PV product uses;
//   * POLYJ of every 8 accumulator columns take their 2^t from the FMA pipe instead of the MUFU pipe:
//     t clamped to >= -126, n = round(t) by the 1.5*2^23 trick, 2^(t-n) by a degree-3 minimax polynomial on
//     [-0.5, 0.5] (relative error 7.5e-5, 50x below the bf16 rounding of P), exponent patched in with one integer
//     multiply-add. All packed f32x2. This moves POLYJ/8 of the exponentials off the 16/clk/SM pipe.
template <int POLYJ>
__device__ __forceinline__ constexpr bool attn_use_poly(int j) { return ((j + 1) * POLYJ) / 8 != (j * POLYJ) / 8; }

__device__ __forceinline__ float2 mul_f32x2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

// 2^t for a pair, t <= 0, on the FMA / ALU pipes
__device__ __forceinline__ uint32_t exp2_poly_pack(float2 t) {
  constexpr float MAGIC = 12582912.0f;                  // 1.5 * 2^23: t + MAGIC holds round(t) in its low mantissa bits
  t.x = fmaxf(t.x, -126.0f);
  t.y = fmaxf(t.y, -126.0f);
  const float2 t2 = add_f32x2(t, make_float2(MAGIC, MAGIC));
  const float2 n = add_f32x2(t2, make_float2(-MAGIC, -MAGIC));
  const float2 f = add_f32x2(t, make_float2(-n.x, -n.y));
  float2 q = fma_f32x2(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.242611140f, 0.242611140f));
  q = fma_f32x2(q, f, make_float2(0.693260968f, 0.693260968f));
  q = fma_f32x2(q, f, make_float2(0.999928057f, 0.999928057f));
  const uint32_t r0 = __float_as_uint(q.x) + (__float_as_uint(t2.x) << 23);
  const uint32_t r1 = __float_as_uint(q.y) + (__float_as_uint(t2.y) << 23);
  return pack_bf16(__uint_as_float(r0), __uint_as_float(r1));
}

template <int D, int POLYJ, int NW, int TK>
__global__ void __launch_bounds__(32 * NW) attention2_kernel(const AttnParams p) {
  constexpr int BKV = TK;               // keys per tile (shadows the file-level 64): 48 when that pads Lk less (Lk = 282)
  constexpr int BQ = 16 * NW;           // queries per CTA (shadows the file-level 64)
  constexpr int ATT_THREADS = 32 * NW;
  constexpr int PITCH = D + 8;          // elements; +16 bytes
  constexpr int CHUNKS = D / 8;         // 16-byte chunks per row
  __shared__ __align__(16) __nv_bfloat16 sQ[BQ * PITCH];
  __shared__ __align__(16) __nv_bfloat16 sK[2][BKV * PITCH];
  __shared__ __align__(16) __nv_bfloat16 sV[2][BKV * PITCH];

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  pdl_wait();
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;

  const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(p.q) + b * p.bs_q + h * D;
  const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(p.k) + b * p.bs_k + h * D;
  const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(p.v) + b * p.bs_v + h * D;

  for (int i = tid; i < BQ * CHUNKS; i += ATT_THREADS) {
    const int r = i / CHUNKS, c = i % CHUNKS;
    const int row = q0 + r;
    const bool ok = row < p.Lq;
    cp_async16(&sQ[r * PITCH + c * 8], Q + static_cast<size_t>(ok ? row : 0) * p.ld_q + c * 8, ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * BKV;
    for (int i = tid; i < BKV * CHUNKS; i += ATT_THREADS) {
      const int r = i / CHUNKS, c = i % CHUNKS;
      const int row = k0 + r;
      const bool ok = row < p.Lk;
      const size_t rr = static_cast<size_t>(ok ? row : 0);
      cp_async16(&sK[buf][r * PITCH + c * 8], K + rr * p.ld_k + c * 8, ok ? 16 : 0);
      cp_async16(&sV[buf][r * PITCH + c * 8], V + rr * p.ld_v + c * 8, ok ? 16 : 0);
    }
  };
  const int ntiles = (p.Lk + BKV - 1) / BKV;
  load_kv(0, 0);
  cp_async_commit();

  constexpr float LOG2E = 1.4426950408889634f;
  const float sl2 = p.scale * LOG2E;
  const float2 sl2v = make_float2(sl2, sl2);
  const int kvlen = p.kv_len ? p.kv_len[b] : p.Lk;
  const float mask_raw = -10000.0f / p.scale;          // the reference's additive -10000 bias, in raw-score units
  const int kvalid = min(kvlen, p.Lk);                 // keys below this index need no masking
#ifdef DV_F16
  constexpr uint32_t ONES = 0x3C003C00u;               // f16x2 (1, 1)
#else
  constexpr uint32_t ONES = 0x3F803F80u;               // bf16x2 (1, 1)
#endif

  uint32_t qf[D / 16][4];
  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float lacc[4] = {0.f, 0.f, 0.f, 0.f};                // row sums: columns of P . ones (every column is the sum)
  float m0 = -INFINITY, m1 = -INFINITY;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) {
      load_kv(tile + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (tile == 0) {
      const __nv_bfloat16* qb = &sQ[(warp * 16) * PITCH];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        qf[kk][0] = *reinterpret_cast<const uint32_t*>(&qb[g * PITCH + kk * 16 + 2 * t]);
        qf[kk][1] = *reinterpret_cast<const uint32_t*>(&qb[(g + 8) * PITCH + kk * 16 + 2 * t]);
        qf[kk][2] = *reinterpret_cast<const uint32_t*>(&qb[g * PITCH + kk * 16 + 2 * t + 8]);
        qf[kk][3] = *reinterpret_cast<const uint32_t*>(&qb[(g + 8) * PITCH + kk * 16 + 2 * t + 8]);
      }
    }

    float s[BKV / 8][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const __nv_bfloat16* kb = &sK[buf][(j * 8 + g) * PITCH + 2 * t];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kb + kk * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kb + kk * 16 + 8);
        mma_bf16_16816(s[j], qf[kk], b0, b1);
      }
    }

    if (tile * BKV + BKV > kvalid) {
      const int kbase = tile * BKV + 2 * t;
#pragma unroll
      for (int j = 0; j < BKV / 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kbase + j * 8 + e;
          float add = 0.f;
          if (key >= p.Lk) add = -INFINITY;
          else if (key >= kvlen) add = mask_raw;
          s[j][e] += add;
          s[j][2 + e] += add;
        }
      }
    }
    float mx0 = fmaxf(s[0][0], s[0][1]), mx1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int j = 1; j < BKV / 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 key < Lk
    const float c0 = ex2_approx((m0 - mn0) * sl2), c1 = ex2_approx((m1 - mn1) * sl2);
    m0 = mn0; m1 = mn1;
    const float2 nms0 = make_float2(-mn0 * sl2, -mn0 * sl2), nms1 = make_float2(-mn1 * sl2, -mn1 * sl2);
    lacc[0] *= c0; lacc[2] *= c1;
    const float2 c0v = make_float2(c0, c0), c1v = make_float2(c1, c1);
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      const float2 a = mul_f32x2(make_float2(o[j][0], o[j][1]), c0v);
      const float2 bb = mul_f32x2(make_float2(o[j][2], o[j][3]), c1v);
      o[j][0] = a.x; o[j][1] = a.y; o[j][2] = bb.x; o[j][3] = bb.y;
    }

    uint32_t pf[BKV / 16][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      const float2 ta = fma_f32x2(make_float2(s[j][0], s[j][1]), sl2v, nms0);
      const float2 tb = fma_f32x2(make_float2(s[j][2], s[j][3]), sl2v, nms1);
      if (attn_use_poly<POLYJ>(j)) {
        pf[j >> 1][(j & 1) * 2 + 0] = exp2_poly_pack(ta);
        pf[j >> 1][(j & 1) * 2 + 1] = exp2_poly_pack(tb);
      } else {
        pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(ex2_approx(ta.x), ex2_approx(ta.y));
        pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(ex2_approx(tb.x), ex2_approx(tb.y));
      }
    }

#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      mma_bf16_16816(lacc, pf[kk], ONES, ONES);
#pragma unroll
      for (int jd = 0; jd < D / 8; jd += 2) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, &sV[buf][krow * PITCH + (jd + (lane >> 4)) * 8]);
        mma_bf16_16816(o[jd], pf[kk], vf[0], vf[1]);
        mma_bf16_16816(o[jd + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();
  }

  const float i0 = 1.f / lacc[0], i1 = 1.f / lacc[2];
  __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(p.o) + b * p.bs_o + h * D;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < D / 8; ++j) {
    if (r0 < p.Lq)
      *reinterpret_cast<uint32_t*>(&O[static_cast<size_t>(r0) * p.ld_o + j * 8 + 2 * t]) =
          pack_bf16(o[j][0] * i0, o[j][1] * i0);
    if (r1 < p.Lq)
      *reinterpret_cast<uint32_t*>(&O[static_cast<size_t>(r1) * p.ld_o + j * 8 + 2 * t]) =
          pack_bf16(o[j][2] * i1, o[j][3] * i1);
  }
}


constexpr int kDefaultPolyJ = 0;   // measured (profiles/r1f_attention_v2.txt): every polynomial share is slower

template <int POLYJ>
cudaError_t launch_attention2(const AttnParams& p, cudaStream_t st) {
  dim3 grid((p.Lq + BQ - 1) / BQ, p.H, p.B);
  // d = 32 / 48: 8 warps / 128 queries per CTA share every K/V tile load between twice the rows: 38.6 vs 40.2 us (self) and 26.3
  // vs 27.3 us (cross) at the C2 level-1 shapes, 18.3 vs 18.8 and 21.3 vs 22.0 us at level 2; at d = 16 the same is slower
  // (104.6 vs 100.4 us) and at d = 64 the tiles do not fit the static 48 KB. DV_ATTN_NW4=1 turns it off.
  // 48-key tiles when they pad the key count less than 64-key tiles do (prompt cross-attention: Lk = 282 -> 288 instead of 320)
  static const bool tk48 = getenv("DV_ATTN_NO48") == nullptr;
  // (C2 shapes, us: d = 16 31.7 vs 35.4, d = 32 24.4 vs 26.3, d = 48 20.8 vs 21.3; d = 64 is slower with them: 15.5 vs 14.1)
  const bool use48 = tk48 && POLYJ == 0 && p.D <= 48 && ((p.Lk + 47) / 48) * 48 * 100 < ((p.Lk + 63) / 64) * 64 * 92;
  static const bool wide = getenv("DV_ATTN_NW4") == nullptr;
  if (POLYJ == 0 && wide && p.Lq > 64) {
    dim3 grid8((p.Lq + 127) / 128, p.H, p.B);
    if (p.D == 32) return use48 ? launch_k(attention2_kernel<32, 0, 8, 48>, grid8, dim3(256), 0, st, p)
                                : launch_k(attention2_kernel<32, 0, 8, 64>, grid8, dim3(256), 0, st, p);
    if (p.D == 48) return use48 ? launch_k(attention2_kernel<48, 0, 8, 48>, grid8, dim3(256), 0, st, p)
                                : launch_k(attention2_kernel<48, 0, 8, 64>, grid8, dim3(256), 0, st, p);
  }
  if (use48) {
    if (p.D == 16) return launch_k(attention2_kernel<16, 0, 4, 48>, grid, dim3(ATT_THREADS), 0, st, p);
    if (p.D == 32) return launch_k(attention2_kernel<32, 0, 4, 48>, grid, dim3(ATT_THREADS), 0, st, p);
    if (p.D == 48) return launch_k(attention2_kernel<48, 0, 4, 48>, grid, dim3(ATT_THREADS), 0, st, p);
  }
  switch (p.D) {
    case 16: return launch_k(attention2_kernel<16, POLYJ, 4, 64>, grid, dim3(ATT_THREADS), 0, st, p);
    case 32: return launch_k(attention2_kernel<32, POLYJ, 4, 64>, grid, dim3(ATT_THREADS), 0, st, p);
    case 48: return launch_k(attention2_kernel<48, POLYJ, 4, 64>, grid, dim3(ATT_THREADS), 0, st, p);
    case 64: return launch_k(attention2_kernel<64, POLYJ, 4, 64>, grid, dim3(ATT_THREADS), 0, st, p);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

// mma.sync kernel of round 1; kept selectable (DV_ATTN_LEGACY=1) as the A/B bar for the tcgen05 kernel
cudaError_t launch_attention_legacy(const AttnParams& p, cudaStream_t st) {
  if (p.Lk < 1 || p.Lq < 1) return cudaErrorInvalidValue;
  dim3 grid((p.Lq + BQ - 1) / BQ, p.H, p.B);
  switch (p.D) {
    case 16: return launch_k(attention_kernel<16>, grid, dim3(ATT_THREADS), 0, st, p);
    case 32: return launch_k(attention_kernel<32>, grid, dim3(ATT_THREADS), 0, st, p);
    case 48: return launch_k(attention_kernel<48>, grid, dim3(ATT_THREADS), 0, st, p);
    case 64: return launch_k(attention_kernel<64>, grid, dim3(ATT_THREADS), 0, st, p);
    default: return cudaErrorInvalidValue;
  }
}

// Dispatch. The kernels implement the same contract and all are parity-tested at every head dim. Which one runs is a
// measured choice (B200, profiles/r2e_attention_dispatch.txt): the tcgen05 kernel (TMEM accumulators, TMA-fed operands,
// csrc/attention_tc.cu) is the default for head dims >= 32 (levels 1-3: 22 of the 32 attention launches of a call; stand-alone
// 37 vs 41, 15.7 vs 18.7, 6.5 vs 8.7 us; C2 step 201.6 vs 202.4 ms, C5 step 150.6 vs 156.3 ms); at d = 16 (level 0) a score
// costs 64 FLOP per exponential, both kernels are bound by the MUFU pipe and per-tile latency, and the register-resident
// mma.sync kernel with 8+ CTAs of TLP per SM is faster (100 vs 117 us; C2 step 210.9 ms with tcgen05 everywhere).
// DV_ATTN=tc|legacy forces one kernel, DV_ATTN_TC_MIN_D moves the threshold.
constexpr int kTcMinD = 32;        // measured choice, see profiles/r2e_attention_dispatch.txt
cudaError_t launch_attention(const AttnParams& p, cudaStream_t st) {
  if (p.q_lo != nullptr) return DV_NS::launch_attention_split(p, st);     // FP32-class precision mode (hi + lo planes)
  // read per launch (launches are captured into a graph on the hot path): tests flip it between calls
  const char* e = getenv("DV_ATTN");
  if (e && e[0] == 't') return DV_NS::launch_attention_tc(p, st);   // tcgen05 kernel
  {
    // head dims from which the tcgen05 kernel is the default (it wins stand-alone for d >= 32; DV_ATTN_TC_MIN_D overrides)
    static const int tc_min_d = [] { const char* m = getenv("DV_ATTN_TC_MIN_D"); return m ? atoi(m) : kTcMinD; }();
    if (!(e && e[0] == 'l') && p.D >= tc_min_d) return DV_NS::launch_attention_tc(p, st);
  }
  if (e && e[0] == 'l') return DV_NS::launch_attention_legacy(p, st);      // first mma.sync kernel
  if (p.Lk < 1 || p.Lq < 1) return cudaErrorInvalidValue;
  const char* pj = getenv("DV_ATTN_POLY");                          // accumulator columns (of 8) on the FMA-pipe exp2
  const int polyj = pj ? atoi(pj) : kDefaultPolyJ;
  switch (polyj) {
    case 0: return launch_attention2<0>(p, st);
    case 2: return launch_attention2<2>(p, st);
    case 3: return launch_attention2<3>(p, st);
    case 4: return launch_attention2<4>(p, st);
    case 5: return launch_attention2<5>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace DV_NS

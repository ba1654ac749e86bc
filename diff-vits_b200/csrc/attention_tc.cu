// Flash-style attention on the 5th-gen tensor cores (tcgen05 / TMEM) for the UNet's 8-head self-attention and the
// cross-attention to the speech prompt (reference: unet1d/attention_processor.py:980-1052,
// F.scaled_dot_product_attention at :1032; key-padding bias unet_1d_condition.py:816-818).
//
// One CTA = 128 queries of one (batch item, head); two CTAs are resident per SM. Key tiles are 64 wide.
//   warps 0-3 : softmax. TMEM lane = query row, one thread owns one row: no shuffles in the row max / row sum.
//               S tile (fp32, 64 columns) tcgen05.ld -> registers -> online softmax (scale folded into ex2) ->
//               P as bf16 straight into the shared-memory operand layout of the P.V MMA (K-major, 128-byte swizzle)
//               -> fence.proxy.async + mbarrier. The P.V product of every key tile lands in a fresh TMEM
//               accumulator and is folded into the register-resident output row one tile later
//               (O = O * corr + PV), so the tensor pipe never waits for a rescale of TMEM contents.
//   warp 4    : MMA issuer (one lane). S(j+1) = Q.K^T is issued before P(j).V, so the tensor core fills the other
//               S buffer while the softmax warps work on tile j. M = 128, N = 64 (S) / D (PV), K = 16 per instruction.
//   warp 5    : loader. Q / K rows are copied into 128-byte-pitch rows with the UMMA 128-byte swizzle (only the
//               first D columns of a row are used: the MMA K loop covers D/16 steps); V is transposed on the way in
//               (V^T [D x 64 keys], K-major for the P.V product). Out-of-range keys are zero filled. 3-stage ring.
// At d = 16 a score costs 64 FLOP per exp: the kernel is bound by the MUFU pipe (16 ex2 / clk / SM), the tensor
// pipe only has to stay out of the way - which is the point of moving both GEMMs off the issue slots.
//
// Masking: self-attention has no mask (padded frames participate, as in the reference). Cross-attention adds the
// reference's -10000 to the scaled score of keys >= kv_len[b], literally, so a fully masked row degrades the same.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace dv {
namespace {

constexpr int TQ = 128;          // queries per CTA (= TMEM lanes)
constexpr int TK = 64;           // keys per tile
constexpr int NS = 3;            // K/V ring stages
constexpr int AT_THREADS = 192;

template <int D>
struct AttSmem {
  static constexpr int Q_BYTES = TQ * 128;
  static constexpr int K_BYTES = TK * 128;
  static constexpr int V_BYTES = TK * 128;    // same row layout as K; consumed as an MN-major B operand
  static constexpr int P_BYTES = TQ * 128;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_OFF + Q_BYTES;
  static constexpr int V_OFF = K_OFF + NS * K_BYTES;
  static constexpr int P_OFF = V_OFF + NS * V_BYTES;
  static constexpr int BAR_OFF = P_OFF + 2 * P_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;   // + slack for the 1024-byte alignment of the operand boxes
  static_assert(V_BYTES % 1024 == 0, "operand boxes must stay 1024-byte aligned");
  static_assert(2 * TOTAL <= 227 * 1024, "two CTAs per SM");
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(static_cast<unsigned short>(v)) : "memory");
}

// ROWS rows of D bf16 (row stride ld) -> 128-byte-pitch rows, 16-byte pieces XOR-swizzled by (row & 7): the K-major
// SWIZZLE_128B operand layout. Rows >= row_end are zero filled. One warp; consecutive lanes read consecutive pieces.
template <int D, int ROWS>
__device__ __forceinline__ void load_rows_sw128(uint32_t box, const __nv_bfloat16* src, int ld, int row0, int row_end,
                                                int lane) {
  constexpr int CH = D / 8;
  constexpr int N = ROWS * CH / 32;
  constexpr int UB = N % 8 == 0 ? 8 : (N % 6 == 0 ? 6 : (N % 4 == 0 ? 4 : N));   // loads in flight per lane
  static_assert(N % UB == 0, "piece count per lane");
#pragma unroll 1
  for (int i0 = 0; i0 < N; i0 += UB) {
    uint4 v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int t = (i0 + u) * 32 + lane;
      const int r = t / CH, c = t % CH;
      const int row = row0 + r;
      v[u] = row < row_end ? ldg128(src + static_cast<size_t>(row) * ld + c * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int t = (i0 + u) * 32 + lane;
      const int r = t / CH, c = t % CH;
      sts128(box + r * 128 + ((c ^ (r & 7)) << 4), v[u]);
    }
  }
}

// 64 keys x D of V -> V^T [D rows][64 keys] in the same swizzled layout (row d, 16-byte piece key/8, element key%8)
template <int D>
__device__ __forceinline__ void load_v_transposed(uint32_t box, const __nv_bfloat16* src, int ld, int key0, int key_end,
                                                  int lane) {
  constexpr int CH = D / 8;
  constexpr int N = TK * CH / 32;     // 2 pieces per lane per 8 channels
  constexpr int UB = N % 8 == 0 ? 8 : (N % 6 == 0 ? 6 : (N % 4 == 0 ? 4 : N));   // loads in flight per lane
  static_assert(N % UB == 0, "piece count per lane");
#pragma unroll 1
  for (int i0 = 0; i0 < N; i0 += UB) {
    uint4 v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int key = ((i0 + u) & 1) * 32 + lane, c = (i0 + u) >> 1;
      const int row = key0 + key;
      v[u] = row < key_end ? ldg128(src + static_cast<size_t>(row) * ld + c * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int key = ((i0 + u) & 1) * 32 + lane, c = (i0 + u) >> 1;
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      const uint32_t kb = box + (c * 8) * 128 + (key & 7) * 2;
      const int ks = key >> 3;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t val = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xFFFFu);
        sts16(kb + e * 128 + ((ks ^ e) << 4), val);
      }
    }
  }
}

// o[0..D) = o * c + (this thread's row of the P.V accumulator at taddr)
template <int D>
__device__ __forceinline__ void fold_pv(float (&o)[D], float c, uint32_t taddr) {
  if constexpr (D % 32 == 0) {
#pragma unroll
    for (int i = 0; i < D / 32; ++i) {
      uint32_t r[32];
      tmem_ld32(taddr + 32 * i, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) o[32 * i + k] = fmaf(o[32 * i + k], c, __uint_as_float(r[k]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < D / 16; ++i) {
      uint32_t r[16];
      tmem_ld16(taddr + 16 * i, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; ++k) o[16 * i + k] = fmaf(o[16 * i + k], c, __uint_as_float(r[k]));
    }
  }
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 2) attention_tc_kernel(const AttnParams p) {
  using SM = AttSmem<D>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem + SM::Q_OFF, sK = smem + SM::K_OFF, sV = smem + SM::V_OFF, sP = smem + SM::P_OFF;
  const uint32_t bars = smem + SM::BAR_OFF;
  constexpr uint32_t Q_FULL = 0, KV_FULL = 8, KV_EMPTY = KV_FULL + 8 * NS, S_FULL = KV_EMPTY + 8 * NS,
                     P_FULL = S_FULL + 16, O_FULL = P_FULL + 16, TMEM_SLOT = O_FULL + 16;
  static_assert(TMEM_SLOT + 4 <= 256, "barrier block too small");
  constexpr uint32_t S_COL = 0, O_COL = 128;    // TMEM columns: S buffers at 0 / 64, P.V buffers at 128 / 192

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Lk + TK - 1) / TK;

  if (warp == 5 && lane == 0) {
    mbar_init_a(bars + Q_FULL, 1);
    for (int s = 0; s < NS; ++s) { mbar_init_a(bars + KV_FULL + 8 * s, 1); mbar_init_a(bars + KV_EMPTY + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init_a(bars + S_FULL + 8 * a, 1);
      mbar_init_a(bars + P_FULL + 8 * a, 4);
      mbar_init_a(bars + O_FULL + 8 * a, 1);
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + TMEM_SLOT), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bars + TMEM_SLOT));
  pdl_wait();   // Q/K/V come from the previous kernel; the output buffer may still be read by it

  if (warp == 5) {
    // ------------------------------------------------------------------ loader
    const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(p.q) + b * p.bs_q + h * D;
    const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(p.k) + b * p.bs_k + h * D;
    const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(p.v) + b * p.bs_v + h * D;
    load_rows_sw128<D, TQ>(sQ, Q, p.ld_q, q0, p.Lq, lane);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_a(bars + Q_FULL);
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait_a(bars + KV_EMPTY + 8 * st, ph ^ 1);
      load_rows_sw128<D, TK>(sK + st * SM::K_BYTES, K, p.ld_k, j * TK, p.Lk, lane);
      load_rows_sw128<D, TK>(sV + st * SM::V_BYTES, V, p.ld_v, j * TK, p.Lk, lane);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bars + KV_FULL + 8 * st);
      if (++st == NS) { st = 0; ph ^= 1; }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, D) | (1u << 16);   // B (= V tile, [keys][D]) is MN-major
      const uint64_t qdesc = umma_desc_sw128(sQ);
      mbar_wait_a(bars + Q_FULL, 0);
      int st_s = 0;                      // ring stage / phase of the next S tile
      uint32_t ph_s = 0;
      auto issue_s = [&](int j) {
        mbar_wait_a(bars + KV_FULL + 8 * st_s, ph_s);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(sK + st_s * SM::K_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tmem_base + S_COL + (j & 1) * 64, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit_a(bars + S_FULL + 8 * (j & 1));
        if (++st_s == NS) { st_s = 0; ph_s ^= 1; }
      };
      issue_s(0);
      int st = 0;
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        mbar_wait_a(bars + P_FULL + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t pdesc = umma_desc_sw128(sP + (j & 1) * SM::P_BYTES);
        const uint64_t vdesc = umma_desc_sw128(sV + st * SM::V_BYTES);
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)
          umma_bf16(tmem_base + O_COL + (j & 1) * 64, pdesc + 2 * k, vdesc + k * (16 * 128 >> 4), idesc_o, k != 0);   // 16 keys = 2048 B
        umma_commit_a(bars + KV_EMPTY + 8 * st);
        umma_commit_a(bars + O_FULL + 8 * (j & 1));
        if (++st == NS) st = 0;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax: thread = query row
    const int row = warp * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    const int kvlen = p.kv_len ? __ldg(p.kv_len + b) : p.Lk;
    const int kvalid = min(kvlen, p.Lk);                 // keys below this index need no masking
    const float mask_raw = -10000.0f / p.scale;          // the reference's additive -10000 bias, in raw-score units
    float m = -INFINITY, l = 0.f, corr_pend = 0.f;
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    const uint32_t prow = row * 128;
    const int rsw = row & 7;

    for (int j = 0; j < ntiles; ++j) {
      const int bs = j & 1;
      mbar_wait_a(bars + S_FULL + 8 * bs, (j >> 1) & 1);
      tc_fence_after();
      uint32_t sa[32], sb[32];
      tmem_ld32(t_lane + S_COL + bs * 64, sa);
      tmem_ld32(t_lane + S_COL + bs * 64 + 32, sb);
      tmem_ld_wait();
      if (j * TK + TK > kvalid) {                        // warp-uniform: this tile has masked / out-of-range keys
        const int k0 = j * TK;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int ka = k0 + i, kb = k0 + 32 + i;
          float va = __uint_as_float(sa[i]), vb = __uint_as_float(sb[i]);
          va = ka >= p.Lk ? -INFINITY : (ka >= kvlen ? va + mask_raw : va);
          vb = kb >= p.Lk ? -INFINITY : (kb >= kvlen ? vb + mask_raw : vb);
          sa[i] = __float_as_uint(va);
          sb[i] = __float_as_uint(vb);
        }
      }
      float mx0 = __uint_as_float(sa[0]), mx1 = __uint_as_float(sa[1]), mx2 = __uint_as_float(sb[0]), mx3 = __uint_as_float(sb[1]);
#pragma unroll
      for (int i = 2; i < 32; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
        mx2 = fmax3(mx2, __uint_as_float(sb[i]), __uint_as_float(sb[i + 1]));
        if (i + 3 < 32) {
          mx1 = fmax3(mx1, __uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3]));
          mx3 = fmax3(mx3, __uint_as_float(sb[i + 2]), __uint_as_float(sb[i + 3]));
        }
      }
      const float m_new = fmaxf(fmax3(m, mx0, mx1), fmaxf(mx2, mx3));   // finite: every tile has >= 1 key < Lk
      const float corr = ex2f((m - m_new) * sl2);
      const float ms = m_new * sl2;
      m = m_new;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      const uint32_t pbox = sP + bs * SM::P_BYTES + prow;
#pragma unroll
      for (int c = 0; c < 8; ++c) {                      // 16-byte piece c = keys [8c, 8c+8) of this row
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (c & 3) * 8 + 2 * e;
          const float x0 = __uint_as_float(c < 4 ? sa[i] : sb[i]), x1 = __uint_as_float(c < 4 ? sa[i + 1] : sb[i + 1]);
          const float p0 = ex2f(fmaf(x0, sl2, -ms)), p1 = ex2f(fmaf(x1, sl2, -ms));
          if (e & 1) { sum2 += p0; sum3 += p1; } else { sum0 += p0; sum1 += p1; }
          pk[e] = pack_bf16(p0, p1);
        }
        sts128(pbox + ((c ^ rsw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
      fence_proxy_async();      // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
      tc_fence_before();        // the S loads above are complete: the issuer may overwrite this S buffer
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bars + P_FULL + 8 * bs);
      l = fmaf(l, corr, (sum0 + sum1) + (sum2 + sum3));
      if (j > 0) {              // fold the previous tile's P.V (issued one tile ago) into the output row
        mbar_wait_a(bars + O_FULL + 8 * (bs ^ 1), ((j - 1) >> 1) & 1);
        tc_fence_after();
        fold_pv<D>(o, corr_pend, t_lane + O_COL + (bs ^ 1) * 64);
      }
      corr_pend = corr;
    }
    {
      const int jl = ntiles - 1;
      mbar_wait_a(bars + O_FULL + 8 * (jl & 1), (jl >> 1) & 1);
      tc_fence_after();
      fold_pv<D>(o, corr_pend, t_lane + O_COL + (jl & 1) * 64);
    }
    if (q0 + row < p.Lq) {
      const float inv = 1.0f / l;
      __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(p.o) + b * p.bs_o + static_cast<size_t>(q0 + row) * p.ld_o + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c)
        *reinterpret_cast<uint4*>(O + c * 8) =
            make_uint4(pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                       pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

bool g_att_init = false;

template <int D>
cudaError_t launch_d(const AttnParams& p, cudaStream_t st) {
  dim3 grid((p.Lq + TQ - 1) / TQ, p.H, p.B);
  return launch_k(attention_tc_kernel<D>, grid, dim3(AT_THREADS), static_cast<size_t>(AttSmem<D>::TOTAL), st, p);
}

}  // namespace

cudaError_t attention_tc_init() {
  if (g_att_init) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem<16>::TOTAL);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attention_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem<32>::TOTAL);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attention_tc_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem<48>::TOTAL);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attention_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem<64>::TOTAL);
  if (e != cudaSuccess) return e;
  g_att_init = true;
  return cudaSuccess;
}

cudaError_t launch_attention_tc(const AttnParams& p, cudaStream_t st) {
  if (p.Lk < 1 || p.Lq < 1) return cudaErrorInvalidValue;
  if (!g_att_init) {
    cudaError_t e = attention_tc_init();
    if (e != cudaSuccess) return e;
  }
  // 16-byte vector accesses on every operand
  if ((p.ld_q | p.ld_k | p.ld_v | p.ld_o) % 8 != 0 || (p.bs_q | p.bs_k | p.bs_v | p.bs_o) % 8 != 0) return cudaErrorInvalidValue;
  switch (p.D) {
    case 16: return launch_d<16>(p, st);
    case 32: return launch_d<32>(p, st);
    case 48: return launch_d<48>(p, st);
    case 64: return launch_d<64>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dv

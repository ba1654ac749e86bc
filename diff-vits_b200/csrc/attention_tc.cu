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
//   warp 5    : TMA producer (one lane). Q / K / V tiles are [rows x D] boxes of the [B, L, H*D] operands (3-D maps,
//               rows outside [0, L) zero filled), D*2-byte rows with the matching 32/64/128-byte swizzle (d = 48 uses a
//               64-wide box; the MMA only walks the first 48 columns). K is consumed K-major, V as an MN-major B
//               operand of the P.V product - no transpose anywhere. 4-stage ring.
// At d = 16 a score costs 64 FLOP per exp: the kernel is bound by the MUFU pipe (16 ex2 / clk / SM), the tensor
// pipe only has to stay out of the way - which is the point of moving both GEMMs off the issue slots.
//
// Masking: self-attention has no mask (padded frames participate, as in the reference). Cross-attention adds the
// reference's -10000 to the scaled score of keys >= kv_len[b], literally, so a fully masked row degrades the same.
#include <cstdlib>

#include <cudaTypedefs.h>

#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {
namespace {

constexpr int TQ = 128;          // queries per CTA (= TMEM lanes)
constexpr int TK = 64;           // keys per tile
constexpr int AT_THREADS = 320;   // warps 0-7 softmax, warp 8 MMA issuer, warp 9 TMA producer

template <int D>
struct AttSmem {
  static constexpr int DB = D == 48 ? 64 : D;            // box width in elements (power of two: a swizzle span)
  static constexpr int ROW = DB * 2;                     // bytes per operand row = swizzle span (32 / 64 / 128)
  static constexpr int NS = ROW == 128 ? 3 : 4;          // K/V ring stages
  static constexpr int Q_BYTES = TQ * ROW;
  static constexpr int K_BYTES = TK * ROW;
  static constexpr int V_BYTES = TK * ROW;
  static constexpr int P_BYTES = TQ * 128;               // P tile: 64 keys bf16 = 128-byte rows, SWIZZLE_128B
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_OFF + Q_BYTES;
  static constexpr int V_OFF = K_OFF + NS * K_BYTES;
  static constexpr int P_OFF = V_OFF + NS * V_BYTES;
  static constexpr int BAR_OFF = P_OFF + 2 * P_BYTES;
  static constexpr int RED_OFF = BAR_OFF + 256;         // row-max exchange [2 buffers][128 rows][2 halves] + row sums
  static constexpr int TOTAL = RED_OFF + 3072 + 1024;   // + slack for the 1024-byte alignment of the operand boxes
  static constexpr uint32_t LAYOUT = ROW == 32 ? 6u : (ROW == 64 ? 4u : 2u);   // UMMA layout type of the Q/K/V boxes
  static_assert(K_BYTES % 1024 == 0 && Q_BYTES % 1024 == 0, "operand boxes must stay 1024-byte aligned");
  static_assert(2 * TOTAL <= 227 * 1024, "two CTAs per SM");
};

// Shared-memory matrix descriptor for a box of ROW-byte rows with the ROW-byte swizzle: 8-row groups are 8*ROW bytes
// apart (SBO); LBO is unused (K-major: one swizzle span along K; MN-major: the tile is one span wide).
__device__ __forceinline__ uint64_t umma_desc_rows(uint32_t smem_addr, uint32_t row_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * row_bytes) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

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
// barrier of one warp pair (64 threads); ids are immediates so that only 5 hardware barriers are reserved per CTA
__device__ __forceinline__ void pair_bar_sync(int quad) {
  switch (quad) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// waits of the single-lane MMA / TMA roles: back off between polls so the spin does not take issue slots from the
// softmax warps of the same scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_a(bar, parity)) {
    __nanosleep(40);
    if (++spins > 40000000u) mbar_timeout_trap();
  }
}

// o[0..N) = o * c + (this thread's row, N columns at taddr, of the P.V accumulator)
template <int N>
__device__ __forceinline__ void fold_pv(float (&o)[N], float c, uint32_t taddr) {
  static_assert(N == 8 || N == 16 || N == 24 || N == 32, "columns per thread");
  if constexpr (N == 32) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 32; ++k) o[k] = fmaf(o[k], c, __uint_as_float(r[k]));
  } else {
    if constexpr (N >= 16) {
      uint32_t r[16];
      tmem_ld16(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; ++k) o[k] = fmaf(o[k], c, __uint_as_float(r[k]));
    }
    if constexpr (N % 16 == 8) {
      uint32_t r[8];
      tmem_ld8(taddr + (N - 8), r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 8; ++k) o[N - 8 + k] = fmaf(o[N - 8 + k], c, __uint_as_float(r[k]));
    }
  }
}

// debug timeline (enabled only when p.trace != nullptr): CTA (0,0,0), one lane per role; every role appends to its own
// region of the buffer with plain stores (no atomics: the probe must not add a round trip to what it measures)
constexpr int TRACE_CAP = 1300;
__device__ __forceinline__ void trace_ev(const AttnParams& p, int role, int& cnt, int tile, int ev) {
#ifdef DV_DEBUG_HOOKS
  if (p.trace && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && cnt < TRACE_CAP) {
    p.trace[1 + role * TRACE_CAP + cnt] = (static_cast<unsigned long long>(role + 1) << 60) | (static_cast<unsigned long long>(tile & 0xFFFF) << 44) |
                                          (static_cast<unsigned long long>(ev & 0xF) << 40) | (clock64() & 0xFFFFFFFFFFULL);
    ++cnt;
  }
#endif
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 2) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  using SM = AttSmem<D>;
  constexpr int NS = SM::NS;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem + SM::Q_OFF, sK = smem + SM::K_OFF, sV = smem + SM::V_OFF, sP = smem + SM::P_OFF;
  const uint32_t bars = smem + SM::BAR_OFF, red = smem + SM::RED_OFF;
  constexpr uint32_t Q_FULL = 0, KV_FULL = 8, KV_EMPTY = KV_FULL + 8 * NS, S_FULL = KV_EMPTY + 8 * NS,
                     P_FULL = S_FULL + 16, O_FULL = P_FULL + 16, TMEM_SLOT = O_FULL + 16;
  static_assert(TMEM_SLOT + 4 <= 256, "barrier block too small");
  constexpr uint32_t S_COL = 0, O_COL = 128;    // TMEM columns: S buffers at 0 / 64, P.V buffers at 128 / 192

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Lk + TK - 1) / TK;
  int tcnt = 0;   // debug trace cursor

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init_a(bars + Q_FULL, 1);
    for (int s = 0; s < NS; ++s) { mbar_init_a(bars + KV_FULL + 8 * s, 1); mbar_init_a(bars + KV_EMPTY + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init_a(bars + S_FULL + 8 * a, 1);
      mbar_init_a(bars + P_FULL + 8 * a, 8);
      mbar_init_a(bars + O_FULL + 8 * a, 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + TMEM_SLOT), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bars + TMEM_SLOT));
  pdl_wait();   // Q/K/V come from the previous kernel; the output buffer may still be read by it

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx_a(bars + Q_FULL, SM::Q_BYTES);
      tma_load_3d_a(sQ, &p.mapQ, bars + Q_FULL, h * D, q0, b);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait_sleep(bars + KV_EMPTY + 8 * st, ph ^ 1);
        trace_ev(p, 0, tcnt, j, 0);
        mbar_arrive_expect_tx_a(bars + KV_FULL + 8 * st, SM::K_BYTES + SM::V_BYTES);
        tma_load_3d_a(sK + st * SM::K_BYTES, &p.mapK, bars + KV_FULL + 8 * st, h * D, j * TK, b);
        tma_load_3d_a(sV + st * SM::V_BYTES, &p.mapV, bars + KV_FULL + 8 * st, h * D, j * TK, b);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, D) | (1u << 16);   // B (= V tile, [keys][D]) is MN-major
      const uint64_t qdesc = umma_desc_rows(sQ, SM::ROW, SM::LAYOUT);
      mbar_wait_sleep(bars + Q_FULL, 0);
      int st_s = 0;                      // ring stage / phase of the next S tile
      uint32_t ph_s = 0;
      auto issue_s = [&](int j) {
        mbar_wait_sleep(bars + KV_FULL + 8 * st_s, ph_s);
        tc_fence_after();
        trace_ev(p, 1, tcnt, j, 0);
        const uint64_t kdesc = umma_desc_rows(sK + st_s * SM::K_BYTES, SM::ROW, SM::LAYOUT);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tmem_base + S_COL + (j & 1) * 64, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit_a(bars + S_FULL + 8 * (j & 1));
        trace_ev(p, 1, tcnt, j, 1);
        if (++st_s == NS) { st_s = 0; ph_s ^= 1; }
      };
      issue_s(0);
      int st = 0;
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        mbar_wait_sleep(bars + P_FULL + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        trace_ev(p, 1, tcnt, j, 2);
        const uint64_t pdesc = umma_desc_sw128(sP + (j & 1) * SM::P_BYTES);
        const uint64_t vdesc = umma_desc_rows(sV + st * SM::V_BYTES, SM::ROW, SM::LAYOUT);
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)
          umma_bf16(tmem_base + O_COL + (j & 1) * 64, pdesc + 2 * k, vdesc + k * (16 * SM::ROW >> 4), idesc_o, k != 0);   // 16 keys further down
        umma_commit_a(bars + KV_EMPTY + 8 * st);
        umma_commit_a(bars + O_FULL + 8 * (j & 1));
        trace_ev(p, 1, tcnt, j, 3);
        if (++st == NS) st = 0;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax: two threads per query row
    // warps w and w + 4 share TMEM lane quarter w (hardware: quarter = warp id % 4); thread (w, lane) and thread
    // (w + 4, lane) own row 32 w + lane and split every key tile (32 keys each) and the output columns (D/2 each).
    // They exchange their partial row max through shared memory once per tile (named barrier per warp pair).
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr int DH = D / 2;
    const float sl2 = p.scale * LOG2E;
    const int kvlen = p.kv_len ? __ldg(p.kv_len + b) : p.Lk;
    const int kvalid = min(kvlen, p.Lk);                 // keys below this index need no masking
    const float mask_raw = -10000.0f / p.scale;          // the reference's additive -10000 bias, in raw-score units
    float m = -INFINITY, l = 0.f, corr_pend = 0.f;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    const uint32_t prow = row * 128;
    const int rsw = row & 7;
    const uint32_t red_mine = red + (row * 2 + half) * 4, red_other = red + (row * 2 + (half ^ 1)) * 4;

    for (int j = 0; j < ntiles; ++j) {
      const int bs = j & 1;
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 0);
      mbar_wait_a(bars + S_FULL + 8 * bs, (j >> 1) & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 1);
      uint32_t sa[32];
      tmem_ld32(t_lane + S_COL + bs * 64 + half * 32, sa);
      tmem_ld_wait();
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 2);
      if (j * TK + TK > kvalid) {                        // warp-uniform: this tile has masked / out-of-range keys
        const int k0 = j * TK + half * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int ka = k0 + i;
          float va = __uint_as_float(sa[i]);
          va = ka >= p.Lk ? -INFINITY : (ka >= kvlen ? va + mask_raw : va);
          sa[i] = __float_as_uint(va);
        }
      }
      float mx0 = __uint_as_float(sa[0]), mx1 = __uint_as_float(sa[1]);
      float m_new;
      {
#pragma unroll
      for (int i = 2; i < 32; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
        if (i + 3 < 32) mx1 = fmax3(mx1, __uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3]));
      }
      const float mine = fmaxf(mx0, mx1);                // -inf if all 32 keys of this half are beyond Lk
      sts_f32(red_mine + bs * 1024, mine);
      pair_bar_sync(quad);
      m_new = fmax3(m, mine, lds_f32(red_other + bs * 1024));   // finite: every tile has >= 1 key < Lk
      }
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 3);
      const float corr = ex2f((m - m_new) * sl2);
      const float ms = m_new * sl2;
      m = m_new;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      const uint32_t pbox = sP + bs * SM::P_BYTES + prow;
#pragma unroll
      for (int c = 0; c < 4; ++c) {                      // 16-byte piece: keys [32 half + 8c, +8) of this row
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = c * 8 + 2 * e;
          float p0 = fmaf(__uint_as_float(sa[i]), sl2, -ms), p1 = fmaf(__uint_as_float(sa[i + 1]), sl2, -ms);
          p0 = ex2f(p0); p1 = ex2f(p1);
          if (e & 1) { sum2 += p0; sum3 += p1; } else { sum0 += p0; sum1 += p1; }
          pk[e] = pack_bf16(p0, p1);
        }
        sts128(pbox + (((half * 4 + c) ^ rsw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 4);
      fence_proxy_async();      // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
      tc_fence_before();        // the S loads above are complete: the issuer may overwrite this S buffer
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bars + P_FULL + 8 * bs);
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 5);
      l = fmaf(l, corr, (sum0 + sum1) + (sum2 + sum3));
      if (j > 0) {              // fold the previous tile's P.V (issued one tile ago) into the output row
        mbar_wait_a(bars + O_FULL + 8 * (bs ^ 1), ((j - 1) >> 1) & 1);
        tc_fence_after();
        if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 6);
        fold_pv<DH>(o, corr_pend, t_lane + O_COL + (bs ^ 1) * 64 + half * DH);
      }
      corr_pend = corr;
      if (warp == 0 && lane == 0) trace_ev(p, 2, tcnt, j, 7);
    }
    {
      const int jl = ntiles - 1;
      mbar_wait_a(bars + O_FULL + 8 * (jl & 1), (jl >> 1) & 1);
      tc_fence_after();
      fold_pv<DH>(o, corr_pend, t_lane + O_COL + (jl & 1) * 64 + half * DH);
    }
    // row sum = the two halves' partial sums (both are relative to the same running max)
    sts_f32(red_mine + 2048, l);
    pair_bar_sync(quad);
    l += lds_f32(red_other + 2048);
    if (q0 + row < p.Lq) {
      const float inv = 1.0f / l;
      __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(p.o) + b * p.bs_o + static_cast<size_t>(q0 + row) * p.ld_o + h * D + half * DH;
#pragma unroll
      for (int c = 0; c < DH / 8; ++c)
        *reinterpret_cast<uint4*>(O + c * 8) =
            make_uint4(pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                       pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

bool g_att_init = false;

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  return fn;
}

// [rows x D] boxes of a bf16 operand laid out [B, L, ld] whose head h occupies columns [h*D, h*D + D)
bool encode_rows(CUtensorMap* m, const void* base, int cols, int ld, int L, int B, long long bs, int D, int box_rows) {
  auto fn = encode_fn();
  if (!fn) return false;
  const int db = D == 48 ? 64 : D;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(bs) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(db), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = db == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (db == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

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

bool attention_tc_prepare(AttnParams& p, const char** err) {
  static const char* const msg_shape = "attention: unsupported shape / alignment for the TMA path";
  static const char* const msg_enc = "attention: cuTensorMapEncodeTiled failed";
  if ((p.D != 16 && p.D != 32 && p.D != 48 && p.D != 64) || p.Lq < 1 || p.Lk < 1 || p.B < 1 || p.H < 1 ||
      (p.ld_q | p.ld_k | p.ld_v | p.ld_o) % 8 != 0 || (p.bs_q | p.bs_k | p.bs_v | p.bs_o) % 8 != 0 ||
      (reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.k) | reinterpret_cast<uintptr_t>(p.v) |
       reinterpret_cast<uintptr_t>(p.o)) % 16 != 0) {
    if (err) *err = msg_shape;
    return false;
  }
  const int cols = p.H * p.D;
  if (!encode_rows(&p.mapQ, p.q, cols, p.ld_q, p.Lq, p.B, p.bs_q, p.D, TQ) ||
      !encode_rows(&p.mapK, p.k, cols, p.ld_k, p.Lk, p.B, p.bs_k, p.D, TK) ||
      !encode_rows(&p.mapV, p.v, cols, p.ld_v, p.Lk, p.B, p.bs_v, p.D, TK)) {
    if (err) *err = msg_enc;
    return false;
  }
  return true;
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

}  // namespace DV_NS

// Chained tcgen05 GEMMs: the row-local part of a transformer block as ONE persistent kernel per section
// (reference: unet1d/transformer_1d.py:256-300, unet1d/attention.py:130-203, GEGLU :299-301).
//
// Everything between two attention cores is row-local: out-projection + residual, LayerNorm, the next projection,
// the GEGLU feed-forward, proj_out + residual. Launched as separate GEMM / norm kernels each of these is a few
// microseconds of work wrapped in launch, pipeline-fill and drain latency, and every intermediate makes an HBM/L2
// round trip (the [L, 8C] GEGLU pre-activation is the largest tensor of the network). Here a CTA owns a tile of 128
// rows and walks a small stage program (kernels.h: CH_ROW / CH_STREAM / CH_FF); the activation between stages lives
// in shared memory as a ready-made UMMA operand (K-major, 128-byte swizzle, exactly the layout TMA would have
// produced), accumulators live in TMEM, and only weights stream in.
//
//   warp 0      : TMA producer. Streams 16 KB boxes ([128 rows x 64 k] of W, or of the global A operand of the first
//                 stage) through a ring of shared-memory slots, in exactly the order the MMA warp consumes them. It
//                 does not depend on the epilogue, so weights of the next stage / next tile are prefetched while the
//                 current epilogue runs.
//   warp 1      : MMA issuer (one lane): tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 128 per instruction.
//                 TMEM: "main" accumulator [128 x C] at columns [0, C) (row stages, FF output) and one or two 128-column
//                 "streaming" accumulators behind it (GEGLU chunks, streamed output columns).
//   warps 2..17 : epilogue. TMEM lane quarter = warp % 4 (hardware), the four warps of a quarter split the columns.
//                 A thread owns one row: bias / residual / GEGLU / LayerNorm in registers (row statistics exchanged
//                 through shared memory between the four warps of a quarter), results written as bf16 either to
//                 global (coalesced through a per-warp staging slab) or straight into the swizzled operand buffer
//                 of the next stage (fence.proxy.async + mbarrier hand the buffer to the tensor core).
//
// All reductions are fixed-order; numerics equal the unfused kernels (bf16 rounding at the same points, fp32
// accumulation in the same K order), except for the summation order of the LayerNorm statistics.
#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {
namespace {

constexpr int BOX_BYTES = 16384;            // [128 rows x 64 bf16], 128-byte swizzle
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_THREADS = 64 + CH_EPI_WARPS * 32;
constexpr int CH_MAX_SLOTS = 8;
constexpr int RED_BYTES = 2 * 128 * 4 * 4;  // row-statistic exchange: [2][128 rows][4 parts] floats
constexpr int BAR_BYTES = 512;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// Per-warp staging slab (32 rows x 64 bytes, XOR-swizzled 16-byte pieces instead of padding): a thread owns row
// `lane` in the compute phase; in the row-contiguous phase lane -> (row = i*8 + lane/4, piece = lane%4), so that
// global accesses are 64-byte row segments.
struct Slab {
  uint32_t base;
  int lane, srow, spiece;
  __device__ __forceinline__ uint32_t own(int j) const { return base + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4); }
  __device__ __forceinline__ uint32_t coal(int i) const {
    const int row = i * 8 + srow;
    return base + row * 64 + ((spiece ^ ((row >> 1) & 3)) << 4);
  }
};

// residual fetch: every thread reads the 64 bytes of its own row (4 x 16 B; the four loads of a warp cover whole
// 32-byte sectors between them)
__device__ __forceinline__ void res_load(const __nv_bfloat16* res, int ld, size_t grow, bool valid, int gcol,
                                         uint32_t (&rr)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 t = valid ? *reinterpret_cast<const uint4*>(res + grow * ld + gcol + j * 8) : make_uint4(0, 0, 0, 0);
    rr[4 * j] = t.x; rr[4 * j + 1] = t.y; rr[4 * j + 2] = t.z; rr[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void store_chunk(const Slab& s, __nv_bfloat16* out, int ld, size_t row0, int rows_left, int gcol,
                                            const uint32_t (&packed)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(s.own(j), make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + s.srow;
    if (r < rows_left) *reinterpret_cast<uint4*>(out + (row0 + r) * ld + gcol + s.spiece * 8) = lds128(s.coal(i));
  }
  __syncwarp();
}

// acc (+bias) (+res) -> packed bf16; optionally the fp32 sums needed for GroupNorm partials
template <bool STATS>
__device__ __forceinline__ void finish_chunk(const uint32_t (&r)[32], const float* bias, int gcol, bool has_res,
                                             const uint32_t (&rr)[16], uint32_t (&packed)[16], float (&st)[4]) {
  float2 s0 = make_float2(0.f, 0.f), q0 = s0, s1 = s0, q1 = s0;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bq = __ldg(reinterpret_cast<const float4*>(bias + gcol + 4 * j4));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 4 * j4 + 2 * h;
      float2 v = make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
      v = add_f32x2(v, h ? make_float2(bq.z, bq.w) : make_float2(bq.x, bq.y));
      if (has_res) v = add_f32x2(v, make_float2(bf16_lo(rr[j >> 1]), bf16_hi(rr[j >> 1])));
      packed[j >> 1] = pack_bf16(v.x, v.y);
      if (STATS) {
        if (j < 16) { s0 = add_f32x2(s0, v); q0 = fma_f32x2(v, v, q0); }
        else        { s1 = add_f32x2(s1, v); q1 = fma_f32x2(v, v, q1); }
      }
    }
  }
  if (STATS) { st[0] = s0.x + s0.y; st[1] = q0.x + q0.y; st[2] = s1.x + s1.y; st[3] = q1.x + q1.y; }
}

// 16-byte piece `c16` (0..7) of row `r` inside a [128 x 64] bf16 operand box with the 128-byte swizzle
__device__ __forceinline__ uint32_t box_addr(uint32_t box, int r, int c16) {
  return box + r * 128 + ((c16 ^ (r & 7)) << 4);
}

template <int U>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const __grid_constant__ ChainParams p, int nslot) {
  constexpr int KP = 2 * U;                     // 64-wide K panels of a C-wide operand
  constexpr int C = 128 * U;
  constexpr int NPP = U <= 2 ? 2 : 1;           // streaming accumulators (128 columns each)
  constexpr int PP_COL = U <= 2 ? 256 : 384;    // first TMEM column of the streaming accumulators
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // shared memory, as 32-bit shared-window addresses (one base register; everything else is an immediate offset)
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = smem;                               // KP boxes: the operand handed from stage to stage
  const uint32_t sH = sX + KP * BOX_BYTES;                // 2 boxes: GEGLU output chunks; doubles as the 16 staging slabs
  const uint32_t sRing = sH + 2 * BOX_BYTES;              // nslot boxes
  const uint32_t sRed = sRing + nslot * BOX_BYTES;        // row-statistic exchange
  const uint32_t sBias = sRed + RED_BYTES;                // GEGLU bias [8 C] (interleaved u/g packing)
  const uint32_t bars = sBias + 8 * C * 4;
  // barrier k lives at bars + 8 k
  constexpr uint32_t FULL = 0, EMPTY = 8 * CH_MAX_SLOTS, ACC_FULL = EMPTY + 8 * CH_MAX_SLOTS, ACC_EMPTY = ACC_FULL + 16,
                     MAIN_FULL = ACC_EMPTY + 16, MAIN_EMPTY = MAIN_FULL + 8, X_READY = MAIN_EMPTY + 8, X_FREE = X_READY + 8,
                     H_READY = X_FREE + 8, H_FREE = H_READY + 16, TMEM_SLOT = H_FREE + 16;
  static_assert(TMEM_SLOT + 4 <= BAR_BYTES, "barrier block too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_per_batch * p.B;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.mapA[i]);
    for (int i = 0; i < CHAIN_MAX_W; ++i) tma_prefetch_desc(&p.mapW[i]);
    for (int s = 0; s < CH_MAX_SLOTS; ++s) { mbar_init_a(bars + FULL + 8 * s, 1); mbar_init_a(bars + EMPTY + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init_a(bars + ACC_FULL + 8 * a, 1); mbar_init_a(bars + ACC_EMPTY + 8 * a, CH_EPI_WARPS);
      mbar_init_a(bars + H_READY + 8 * a, CH_EPI_WARPS); mbar_init_a(bars + H_FREE + 8 * a, 1);
    }
    mbar_init_a(bars + MAIN_FULL, 1); mbar_init_a(bars + MAIN_EMPTY, CH_EPI_WARPS);
    mbar_init_a(bars + X_READY, CH_EPI_WARPS); mbar_init_a(bars + X_FREE, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + TMEM_SLOT), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {                               // static data: may be fetched before the producer kernel has finished
    for (int s = 0; s < p.nstages; ++s)
      if (p.st[s].kind == CH_FF && p.st[s].bias)
        for (int i = threadIdx.x - 64; i < 2 * p.st[s].n; i += CH_EPI_WARPS * 32) sts_f32(sBias + 4 * i, __ldg(p.st[s].bias + i));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bars + TMEM_SLOT));
  pdl_wait();   // everything below reads activations written by earlier kernels

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      auto push_w = [&](const CUtensorMap* m, int k0, int r0) {
        mbar_wait_a(bars + EMPTY + 8 * slot, ph ^ 1);
        mbar_arrive_expect_tx_a(bars + FULL + 8 * slot, BOX_BYTES);
        tma_load_2d_a(sRing + slot * BOX_BYTES, m, bars + FULL + 8 * slot, k0, r0);
        if (++slot == nslot) { slot = 0; ph ^= 1; }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_batch;
        const int l0 = (tile % p.tiles_per_batch) * 128;
        for (int s = 0; s < p.nstages; ++s) {
          const int kind = p.st[s].kind;
          const CUtensorMap* mw = &p.mapW[p.st[s].w_map];
          if (kind == CH_ROW) {
            const int a_map = p.st[s].a_map;
            for (int kp = 0; kp < KP; ++kp) {
              if (a_map >= 0) {
                mbar_wait_a(bars + EMPTY + 8 * slot, ph ^ 1);
                mbar_arrive_expect_tx_a(bars + FULL + 8 * slot, BOX_BYTES);
                tma_load_3d_a(sRing + slot * BOX_BYTES, &p.mapA[a_map], bars + FULL + 8 * slot, kp * 64, l0, b);
                if (++slot == nslot) { slot = 0; ph ^= 1; }
              }
              for (int u = 0; u < U; ++u) push_w(mw, kp * 64, u * 128);
            }
          } else if (kind == CH_STREAM) {
            const int units = p.st[s].n >> 7;
            for (int un = 0; un < units; ++un)
              for (int kp = 0; kp < KP; ++kp) push_w(mw, kp * 64, un * 128);
          } else {
            const CUtensorMap* mw2 = &p.mapW[p.st[s].w2_map];
            const int nch = p.st[s].n >> 6;
            for (int kp = 0; kp < KP; ++kp) push_w(mw, kp * 64, 0);                       // G(0)
            for (int i = 0; i < nch; ++i) {
              if (i + 1 < nch)
                for (int kp = 0; kp < KP; ++kp) push_w(mw, kp * 64, (i + 1) * 128);       // G(i+1)
              for (int u = 0; u < U; ++u) push_w(mw2, i * 64, u * 128);                   // F(i)
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      int slot = 0;
      uint32_t ph = 0;
      uint32_t main_use = 0, acc_cnt = 0, h_cnt = 0, x_ver = 0;
      // wait for the next ring slot; returns its index (released later with umma_commit_a on its EMPTY barrier)
      auto take = [&]() -> int {
        mbar_wait_a(bars + FULL + 8 * slot, ph);
        tc_fence_after();
        const int idx = slot;
        if (++slot == nslot) { slot = 0; ph ^= 1; }
        return idx;
      };
      // D[128 x 128] (+)= A_box[128 x 64] . B_box[128 x 64]^T
      auto mma_box = [&](uint32_t d_tmem, uint32_t a_box, uint32_t b_box, bool fresh) {
        const uint64_t adesc = umma_desc_sw128(a_box);
        const uint64_t bdesc = umma_desc_sw128(b_box);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
      };
      // one accumulator's worth of X . W_rows^T: KP boxes of W against the KP boxes of X
      auto mma_x_panels = [&](uint32_t d_tmem) {
        for (int kp = 0; kp < KP; ++kp) {
          const int wi = take();
          mma_box(d_tmem, sX + kp * BOX_BYTES, sRing + wi * BOX_BYTES, kp == 0);
          umma_commit_a(bars + EMPTY + 8 * wi);
        }
      };

      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < p.nstages; ++s) {
          const int kind = p.st[s].kind;
          if (kind == CH_ROW) {
            const bool a_glob = p.st[s].a_map >= 0;
            mbar_wait_a(bars + MAIN_EMPTY, (main_use & 1) ^ 1);
            tc_fence_after();
            if (!a_glob) { mbar_wait_a(bars + X_READY, x_ver & 1); tc_fence_after(); }
            for (int kp = 0; kp < KP; ++kp) {
              uint32_t a_box = sX + kp * BOX_BYTES;
              int ai = -1;
              if (a_glob) { ai = take(); a_box = sRing + ai * BOX_BYTES; }
              for (int u = 0; u < U; ++u) {
                const int wi = take();
                mma_box(tmem_base + u * 128, a_box, sRing + wi * BOX_BYTES, kp == 0);
                umma_commit_a(bars + EMPTY + 8 * wi);
              }
              if (a_glob) umma_commit_a(bars + EMPTY + 8 * ai);
            }
            if (!a_glob) { umma_commit_a(bars + X_FREE); ++x_ver; }
            umma_commit_a(bars + MAIN_FULL);
            ++main_use;
          } else if (kind == CH_STREAM) {
            mbar_wait_a(bars + X_READY, x_ver & 1);
            tc_fence_after();
            const int units = p.st[s].n >> 7;
            for (int un = 0; un < units; ++un) {
              const uint32_t j = acc_cnt % NPP, par = (acc_cnt / NPP) & 1;
              mbar_wait_a(bars + ACC_EMPTY + 8 * j, par ^ 1);
              tc_fence_after();
              mma_x_panels(tmem_base + PP_COL + j * 128);
              umma_commit_a(bars + ACC_FULL + 8 * j);
              ++acc_cnt;
            }
            umma_commit_a(bars + X_FREE);
            ++x_ver;
          } else {
            mbar_wait_a(bars + X_READY, x_ver & 1);
            tc_fence_after();
            const int nch = p.st[s].n >> 6;
            // order: G(0); then for every chunk i: G(i+1) (so the tensor core works during the GELU of chunk i), F(i)
            for (int i = -1; i < nch; ++i) {
              if (i + 1 < nch) {   // G(i+1): [u | g] pre-activations of hidden units [64 (i+1), +64)
                const uint32_t j = acc_cnt % NPP, par = (acc_cnt / NPP) & 1;
                mbar_wait_a(bars + ACC_EMPTY + 8 * j, par ^ 1);
                tc_fence_after();
                mma_x_panels(tmem_base + PP_COL + j * 128);
                umma_commit_a(bars + ACC_FULL + 8 * j);
                ++acc_cnt;
                if (i + 1 == nch - 1) { umma_commit_a(bars + X_FREE); ++x_ver; }
              }
              if (i >= 0) {        // F(i): main += h_chunk[128 x 64] . W2[:, chunk]^T
                const uint32_t hj = h_cnt & 1, hp = (h_cnt >> 1) & 1;
                mbar_wait_a(bars + H_READY + 8 * hj, hp);
                tc_fence_after();
                if (i == 0) {
                  mbar_wait_a(bars + MAIN_EMPTY, (main_use & 1) ^ 1);
                  tc_fence_after();
                }
                for (int u = 0; u < U; ++u) {
                  const int wi = take();
                  mma_box(tmem_base + u * 128, sH + hj * BOX_BYTES, sRing + wi * BOX_BYTES, i == 0);
                  umma_commit_a(bars + EMPTY + 8 * wi);
                }
                umma_commit_a(bars + H_FREE + 8 * hj);
                ++h_cnt;
                if (i == nch - 1) { umma_commit_a(bars + MAIN_FULL); ++main_use; }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: warps 2..17
    const int quad = warp & 3;                    // TMEM lane quarter (fixed by hardware)
    const int part = (warp - 2) >> 2;             // 0..3: which columns of the quarter's rows
    const int row = quad * 32 + lane;             // this thread's row inside the tile
    Slab sl;
    sl.base = sH + (warp - 2) * 2048;
    sl.lane = lane; sl.srow = lane >> 2; sl.spiece = lane & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t red = sRed + (row * 4) * 4;    // this row's 4 partial slots (second array 2 KB further)
    uint32_t main_use = 0, acc_cnt = 0, h_cnt = 0, x_ver = 0;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int mt = tile % p.tiles_per_batch;
      const int lq = mt * 128 + quad * 32;                 // first row (inside the batch item) of this warp's slab
      const int rows_left = p.L - lq;                       // rows of the slab that exist
      const bool valid = lane < rows_left;
      const size_t row0 = static_cast<size_t>(b) * p.L + lq;

      for (int s = 0; s < p.nstages; ++s) {
        const ChainStage& st = p.st[s];
        const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(st.res);
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(st.out);
        if (st.kind == CH_ROW) {
          mbar_wait_a(bars + MAIN_FULL, main_use & 1);
          tc_fence_after();
          // pass 1: acc (+bias) (+res) -> bf16 -> global; the rounded values go back to TMEM for the LayerNorm passes
          float sum = 0.f;
#pragma unroll 1
          for (int i = 0; i < U; ++i) {
            const int gcol = (part + 4 * i) * 32;
            uint32_t rr[16];
            if (res) res_load(res, st.ld_res, row0 + lane, valid, gcol, rr);
            uint32_t r[32];
            tmem_ld32(t_lane + gcol, r);
            tmem_ld_wait();
            uint32_t packed[16];
            float dummy[4];
            finish_chunk<false>(r, st.bias, gcol, res != nullptr, rr, packed, dummy);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float lo = bf16_lo(packed[j]), hi = bf16_hi(packed[j]);
              sum += lo + hi;
              r[2 * j] = __float_as_uint(lo);
              r[2 * j + 1] = __float_as_uint(hi);
            }
            tmem_st32(t_lane + gcol, r);
            if (out) store_chunk(sl, out, st.ld_out, row0, rows_left, gcol, packed);
          }
          tmem_st_wait();
          float mean = 0.f, rstd = 1.f;
          if (st.gamma) {       // LayerNorm over the C columns of the row (exact two-pass on the bf16-rounded values)
            sts_f32(red + 4 * part, sum);
            named_bar_sync(1 + quad, 128);
            mean = ((lds_f32(red) + lds_f32(red + 4)) + (lds_f32(red + 8) + lds_f32(red + 12))) * (1.0f / C);
            float sq = 0.f;
#pragma unroll 1
            for (int i = 0; i < U; ++i) {
              uint32_t r[32];
              tmem_ld32(t_lane + (part + 4 * i) * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(r[j]) - mean; sq = fmaf(d, d, sq); }
            }
            sts_f32(red + 2048 + 4 * part, sq);
            named_bar_sync(1 + quad, 128);
            const float var = ((lds_f32(red + 2048) + lds_f32(red + 2052)) + (lds_f32(red + 2056) + lds_f32(red + 2060))) * (1.0f / C);
            rstd = rsqrtf(var + st.eps);
          }
          __nv_bfloat16* out2 = reinterpret_cast<__nv_bfloat16*>(st.out2);
          if (!out2) mbar_wait_a(bars + X_FREE, (x_ver & 1) ^ 1);     // the tensor core is done reading the previous X
#pragma unroll 1
          for (int i = 0; i < U; ++i) {
            const int ch = part + 4 * i;
            const uint32_t box = sX + (ch >> 1) * BOX_BYTES;
            uint32_t r[32];
            tmem_ld32(t_lane + ch * 32, r);
            tmem_ld_wait();
            if (i == U - 1) {                              // last TMEM read of the main accumulator by this warp
              tc_fence_before();
              if (lane == 0) mbar_arrive_a(bars + MAIN_EMPTY);
            }
            uint32_t packed[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float v[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[8 * jj + k]);
              if (st.gamma) {
                const int c = ch * 32 + jj * 8;
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(st.gamma + c));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(st.gamma + c + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(st.beta + c));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(st.beta + c + 4));
                v[0] = (v[0] - mean) * rstd * g0.x + b0.x; v[1] = (v[1] - mean) * rstd * g0.y + b0.y;
                v[2] = (v[2] - mean) * rstd * g0.z + b0.z; v[3] = (v[3] - mean) * rstd * g0.w + b0.w;
                v[4] = (v[4] - mean) * rstd * g1.x + b1.x; v[5] = (v[5] - mean) * rstd * g1.y + b1.y;
                v[6] = (v[6] - mean) * rstd * g1.z + b1.z; v[7] = (v[7] - mean) * rstd * g1.w + b1.w;
              }
              packed[4 * jj] = pack_bf16(v[0], v[1]); packed[4 * jj + 1] = pack_bf16(v[2], v[3]);
              packed[4 * jj + 2] = pack_bf16(v[4], v[5]); packed[4 * jj + 3] = pack_bf16(v[6], v[7]);
              if (!out2)
                sts128(box_addr(box, row, (ch & 1) * 4 + jj),
                       make_uint4(packed[4 * jj], packed[4 * jj + 1], packed[4 * jj + 2], packed[4 * jj + 3]));
            }
            if (out2) store_chunk(sl, out2, st.ld_out2, row0, rows_left, ch * 32, packed);
          }
          ++main_use;
          if (!out2) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bars + X_READY);
            ++x_ver;
          }
        } else if (st.kind == CH_STREAM) {
          const int units = st.n >> 7;
          const size_t stat_chunk = (static_cast<size_t>(b) * p.tiles_per_batch + mt) * 4 + quad;
          for (int un = 0; un < units; ++un) {
            const uint32_t j = acc_cnt % NPP, par = (acc_cnt / NPP) & 1;
            const int gcol = un * 128 + part * 32;
            uint32_t rr[16];
            if (res) res_load(res, st.ld_res, row0 + lane, valid, gcol, rr);
            mbar_wait_a(bars + ACC_FULL + 8 * j, par);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(t_lane + PP_COL + j * 128 + part * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            if (lane == 0) mbar_arrive_a(bars + ACC_EMPTY + 8 * j);
            ++acc_cnt;
            uint32_t packed[16];
            float sv[4];
            if (st.stats) {
              finish_chunk<true>(r, st.bias, gcol, res != nullptr, rr, packed, sv);
              if (!valid) { sv[0] = sv[1] = sv[2] = sv[3] = 0.f; }
              sv[0] = warp_sum_f(sv[0]); sv[1] = warp_sum_f(sv[1]); sv[2] = warp_sum_f(sv[2]); sv[3] = warp_sum_f(sv[3]);
              if (lane == 0)
                *reinterpret_cast<float4*>(st.stats + (stat_chunk * st.stats_nblk + (gcol >> 4)) * 2) =
                    make_float4(sv[0], sv[1], sv[2], sv[3]);
            } else {
              finish_chunk<false>(r, st.bias, gcol, res != nullptr, rr, packed, sv);
            }
            store_chunk(sl, out, st.ld_out, row0, rows_left, gcol, packed);
          }
        } else {
          // ---- GEGLU chunks: accumulator columns [0,64) = u, [64,128) = gate of hidden units [64 i, 64 i + 64)
          const int nch = st.n >> 6;
          const bool has_bias = st.bias != nullptr;
          for (int i = 0; i < nch; ++i) {
            const uint32_t j = acc_cnt % NPP, par = (acc_cnt / NPP) & 1;
            mbar_wait_a(bars + ACC_FULL + 8 * j, par);
            tc_fence_after();
            uint32_t u[16], g[16];
            tmem_ld16(t_lane + PP_COL + j * 128 + part * 16, u);
            tmem_ld16(t_lane + PP_COL + j * 128 + 64 + part * 16, g);
            tmem_ld_wait();
            tc_fence_before();
            if (lane == 0) mbar_arrive_a(bars + ACC_EMPTY + 8 * j);
            ++acc_cnt;
            uint32_t packed[8];
            const uint32_t sb = sBias + (i * 128 + part * 16) * 4;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float4 bu = make_float4(0.f, 0.f, 0.f, 0.f), bg = bu;
              if (has_bias) {
                bu = lds128f(sb + 16 * j4);
                bg = lds128f(sb + 256 + 16 * j4);
              }
              const int jb = 4 * j4;
              const float2 ua = add_f32x2(make_float2(__uint_as_float(u[jb]), __uint_as_float(u[jb + 1])), make_float2(bu.x, bu.y));
              const float2 ub = add_f32x2(make_float2(__uint_as_float(u[jb + 2]), __uint_as_float(u[jb + 3])), make_float2(bu.z, bu.w));
              const float2 ga = add_f32x2(make_float2(__uint_as_float(g[jb]), __uint_as_float(g[jb + 1])), make_float2(bg.x, bg.y));
              const float2 gb = add_f32x2(make_float2(__uint_as_float(g[jb + 2]), __uint_as_float(g[jb + 3])), make_float2(bg.z, bg.w));
              if (DV_DBG(p, 1)) {
                packed[2 * j4] = pack_bf16(ua.x * ga.x, ua.y * ga.y);
                packed[2 * j4 + 1] = pack_bf16(ub.x * gb.x, ub.y * gb.y);
              } else {
                packed[2 * j4] = pack_bf16(ua.x * gelu_fast_f(ga.x), ua.y * gelu_fast_f(ga.y));
                packed[2 * j4 + 1] = pack_bf16(ub.x * gelu_fast_f(gb.x), ub.y * gelu_fast_f(gb.y));
              }
            }
            const uint32_t hj = h_cnt & 1, hp = (h_cnt >> 1) & 1;
            mbar_wait_a(bars + H_FREE + 8 * hj, hp ^ 1);   // the tensor core is done with this buffer's previous chunk
            const uint32_t box = sH + hj * BOX_BYTES;
            sts128(box_addr(box, row, 2 * part), make_uint4(packed[0], packed[1], packed[2], packed[3]));
            sts128(box_addr(box, row, 2 * part + 1), make_uint4(packed[4], packed[5], packed[6], packed[7]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bars + H_READY + 8 * hj);
            ++h_cnt;
          }
          // ---- feed-forward output: main (+bias2) (+res) -> X (bf16 operand of the next stage)
          mbar_wait_a(bars + MAIN_FULL, main_use & 1);
          tc_fence_after();
          mbar_wait_a(bars + X_FREE, (x_ver & 1) ^ 1);
#pragma unroll 1
          for (int i = 0; i < U; ++i) {
            const int ch = part + 4 * i;
            const int gcol = ch * 32;
            uint32_t rr[16];
            if (res) res_load(res, st.ld_res, row0 + lane, valid, gcol, rr);
            uint32_t r[32];
            tmem_ld32(t_lane + gcol, r);
            tmem_ld_wait();
            if (i == U - 1) {
              tc_fence_before();
              if (lane == 0) mbar_arrive_a(bars + MAIN_EMPTY);
            }
            uint32_t packed[16];
            float dummy[4];
            finish_chunk<false>(r, st.bias2, gcol, res != nullptr, rr, packed, dummy);
            if (out) store_chunk(sl, out, st.ld_out, row0, rows_left, gcol, packed);
            const uint32_t box = sX + (ch >> 1) * BOX_BYTES;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              sts128(box_addr(box, row, (ch & 1) * 4 + jj),
                     make_uint4(packed[4 * jj], packed[4 * jj + 1], packed[4 * jj + 2], packed[4 * jj + 3]));
          }
          ++main_use;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(bars + X_READY);
          ++x_ver;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_chain_sms = 0;

inline int chain_fixed_bytes(int U) { return (2 * U + 2) * BOX_BYTES + RED_BYTES + 8 * 128 * U * 4 + BAR_BYTES + 1024; }
inline int chain_slots(int U) {
  int n = (232448 - chain_fixed_bytes(U)) / BOX_BYTES;
  return n > CH_MAX_SLOTS ? CH_MAX_SLOTS : n;
}

template <int U>
cudaError_t launch_u(const ChainParams& p, cudaStream_t st) {
  const int nslot = chain_slots(U);
  const int smem = chain_fixed_bytes(U) + nslot * BOX_BYTES;
  const int tiles = p.tiles_per_batch * p.B;
  const int grid = tiles < g_chain_sms ? tiles : g_chain_sms;
  return launch_k(chain_kernel<U>, dim3(grid), dim3(CH_THREADS), static_cast<size_t>(smem), st, p, nslot);
}

}  // namespace

cudaError_t chain_tc_init() {
  if (g_chain_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_chain_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       chain_fixed_bytes(1) + chain_slots(1) * BOX_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           chain_fixed_bytes(2) + chain_slots(2) * BOX_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(chain_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              chain_fixed_bytes(3) + chain_slots(3) * BOX_BYTES);
}

cudaError_t launch_chain_tc(const ChainParams& p, cudaStream_t st) {
  if (g_chain_sms == 0) {
    cudaError_t e = chain_tc_init();
    if (e != cudaSuccess) return e;
  }
  if (p.nstages < 1 || p.nstages > CHAIN_MAX_STAGES || p.C % 128 != 0 || p.C < 128 || p.C > 384) return cudaErrorInvalidValue;
  // X must be written before it is read, and every version of X is read by exactly one stage
  bool x_live = false;
  for (int s = 0; s < p.nstages; ++s) {
    const ChainStage& cs = p.st[s];
    const bool reads_x = cs.kind != CH_ROW || cs.a_map < 0;
    const bool writes_x = cs.kind == CH_FF || (cs.kind == CH_ROW && !cs.out2);
    if (cs.out2 && cs.kind != CH_ROW) return cudaErrorInvalidValue;
    if (reads_x && !x_live) return cudaErrorInvalidValue;
    if (!reads_x && x_live) return cudaErrorInvalidValue;
    x_live = writes_x;
    if (cs.kind == CH_STREAM && (cs.n % 128 != 0 || cs.n < 128 || !cs.out)) return cudaErrorInvalidValue;
    if (cs.kind == CH_FF && (cs.n % 64 != 0 || cs.n < 64)) return cudaErrorInvalidValue;
  }
  if (x_live) return cudaErrorInvalidValue;
  switch (p.C / 128) {
    case 1: return launch_u<1>(p, st);
    case 2: return launch_u<2>(p, st);
    default: return launch_u<3>(p, st);
  }
}

}  // namespace DV_NS

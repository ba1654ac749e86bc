// Tap-GEMM, second pipeline shape (opt-in: DV_GEMM2=1 / DV_FUSE_GN=1, or any segment with shared taps / a fused norm).
// Same contract, same epilogue and same tcgen05 / TMEM / TMA machinery as csrc/gemm_tc.cu, but:
//   * TWO rings instead of one (A slots of 136 rows, W slots), each with its own producer warp: a k=3 conv stages ONE
//     136-row activation box per channel panel and its three taps read it through row-shifted UMMA descriptors
//     (profiles/r2b_umma_row_shift_probe.txt): a third of the activation traffic;
//   * optional operand-transform warps: GroupNorm (+time scale/shift, +SiLU) of the resnet / transformer input applied to
//     the staged box IN SHARED MEMORY between the TMA load and the MMA, so the normalised tensor never exists in HBM.
// Measured on C2 (B200, profiles/r2d_*): convs run as fast as in the one-ring kernel, but 1x1 / Linear GEMMs lose 20-40 %
// (the single MMA-issuing thread pays two barrier waits per 64-wide panel instead of one), and the fused norm is redone
// N / BN times (every column tile of a row block transforms the same rows) on 4 warps that are MUFU / issue bound:
// the whole step is 4 % (two rings) / 13 % (fused GroupNorm) slower, so the engine's default program does not use it.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace DV_NS {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
// Epilogue warps per CTA (EW): 16 (4 per TMEM lane quarter, 1 CTA per SM, deep TMA ring) or 8 (2 per quarter, small
// enough in registers / shared memory / TMEM for TWO CTAs per SM: one CTA's epilogue and pipeline fill overlap the
// other's main loop, and with PDL the next kernel's CTAs become resident while this kernel's last tiles drain).
constexpr int EPI_PITCH = 80;                       // bytes per staged row: 32 bf16 (64 B) + 16 B pad
constexpr int EPI_SLAB = 32 * EPI_PITCH;            // per epilogue warp
constexpr int MAX_NPAD = 4096;                      // bias vector kept in shared memory for the whole kernel

constexpr int A_ROWS_MAX = 136;                     // conv k3 with shared taps: 128 + 2 halo rows, padded to whole 8-row atoms
constexpr int XF_MAXC = 1024;                       // widest normalised tensor (level-3 concat: 512 + 512 channels)
constexpr int XF_WARPS = 4;

template <int BN, int EW>
struct SmemLayout {
  static constexpr int A_SLOT = A_ROWS_MAX * BK * 2;              // 17408 B = 17 swizzle atoms
  static constexpr int W_SLOT = BN * BK * 2;
  static constexpr int NA = EW == 8 ? 2 : (BN <= 128 ? 4 : 3);
  static constexpr int NW = EW == 8 ? 3 : (BN == 64 ? 8 : (BN == 128 ? 5 : (BN == 192 ? 4 : 3)));
  static constexpr int W_OFF = NA * A_SLOT;
  static constexpr int EPI_OFF = W_OFF + NW * W_SLOT;
  static constexpr int XF_OFF = EPI_OFF + EW * EPI_SLAB;          // coefficient table [2][XF_MAXC] floats + group stats
  static constexpr int XF_BYTES = 2 * XF_MAXC * 4 + 256;
  static constexpr int BAR_OFF = XF_OFF + (EW == 8 ? 0 : XF_BYTES);
  static constexpr int BIAS_OFF = BAR_OFF + 512;       // barriers + tmem slot, then the bias vector (n_pad floats)
  static constexpr int dyn_bytes(int n_pad) { return BIAS_OFF + n_pad * 4 + 1024; }   // + slack for 1024-byte alignment
  static constexpr int max_bytes = EW == 8 ? 116224 : 232448;
  static_assert(dyn_bytes(EW == 8 ? 1024 : MAX_NPAD) <= max_bytes, "exceeds the shared memory budget");
};

// debug timeline (enabled only when p.trace != nullptr): CTA 0, one lane per role
__device__ __forceinline__ void trace_ev(const GemmParams& p, int role, int tile, int ev) {
#ifdef DV_DEBUG_HOOKS
  if (p.trace && blockIdx.x == 0) {
    const unsigned long long i = atomicAdd(p.trace, 1ULL);
    if (i < 4000) p.trace[1 + i] = (static_cast<unsigned long long>(role) << 60) | (static_cast<unsigned long long>(tile & 0xFFFF) << 44) |
                                   (static_cast<unsigned long long>(ev & 0xF) << 40) | (clock64() & 0xFFFFFFFFFFULL);
  }
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

struct EpiCtx {
  uint32_t mine;    // smem address of this thread's staged row
  uint32_t coal;    // smem address for the row-contiguous phase (row = lane/4, 16-byte piece = lane%4)
  uint32_t sbias;   // smem address of the bias vector
  int lane, srow, spiece;
};

__device__ __forceinline__ void epi_load_res(const GemmParams& p, const EpiCtx& e, int gcol, int lq, size_t rowbase, uint4 (&rv)[4],
                                             const void* plane = nullptr) {
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(plane ? plane : p.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + e.srow;
    rv[i] = (lq + r < p.L && !DV_DBG(p, 16)) ? *reinterpret_cast<const uint4*>(res + (rowbase + lq + r) * p.ld_res + gcol + e.spiece * 8)
                           : make_uint4(0, 0, 0, 0);
  }
}

// One 32-row x 32-column chunk of a bf16 output tile: (+bias) (+residual) -> bf16, GroupNorm partials, coalesced store.
template <bool HAS_RES, bool HAS_STATS, bool TRIM, bool SPLIT>
__device__ __forceinline__ void epi_chunk_bf16(const GemmParams& p, const EpiCtx& e, uint32_t taddr_c, int gcol, int lq,
                                              size_t rowbase, size_t obase, bool valid, size_t stat_chunk, bool last,
                                              uint64_t* empty_bar, uint4 (&rv)[4], bool preloaded) {
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  // row-contiguous residual fetch (64-byte segments), in flight during the TMEM load; the first chunk of a tile
  // arrives preloaded (issued before the wait on the accumulator, so its L2 latency hides behind the main loop)
  if (HAS_RES && !preloaded) epi_load_res(p, e, gcol, lq, rowbase, rv);
  uint32_t r[32];
  tmem_ld32(taddr_c, r);
  if (HAS_RES) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sts128(e.coal + i * 8 * EPI_PITCH, rv[i]);
  }
  tmem_ld_wait();
  if (last) {                                           // last TMEM read of this tile by this warp
    tc_fence_before();
    if (e.lane == 0) mbar_arrive(empty_bar);
  }
  uint32_t rr[16];
  uint32_t rrl[SPLIT ? 16 : 1];
  if (HAS_RES) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 t = lds128(e.mine + j * 16);
      rr[4 * j] = t.x; rr[4 * j + 1] = t.y; rr[4 * j + 2] = t.z; rr[4 * j + 3] = t.w;
    }
    if (SPLIT) {                                          // lo plane of the residual, staged the same way
      uint4 rl[4];
      epi_load_res(p, e, gcol, lq, rowbase, rl, p.res_lo);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) sts128(e.coal + i * 8 * EPI_PITCH, rl[i]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 t = lds128(e.mine + j * 16);
        rrl[4 * j] = t.x; rrl[4 * j + 1] = t.y; rrl[4 * j + 2] = t.z; rrl[4 * j + 3] = t.w;
      }
      __syncwarp();
    }
  }
  uint32_t packed[16];
  uint32_t packed_lo[SPLIT ? 16 : 1];
  float2 s0 = make_float2(0.f, 0.f), q0 = s0, s1 = s0, q1 = s0;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {                      // 4 columns per iteration: one broadcast LDS.128 of bias
    float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bq = lds128f(e.sbias + (gcol + 4 * j4) * 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 4 * j4 + 2 * h;
      float2 v = make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
      v = add_f32x2(v, h ? make_float2(bq.z, bq.w) : make_float2(bq.x, bq.y));
      if (HAS_RES) v = add_f32x2(v, make_float2(bf16_lo(rr[j >> 1]), bf16_hi(rr[j >> 1])));
      if (HAS_RES && SPLIT) v = add_f32x2(v, make_float2(bf16_lo(rrl[j >> 1]), bf16_hi(rrl[j >> 1])));
      packed[j >> 1] = pack_bf16(v.x, v.y);
      if (SPLIT) packed_lo[j >> 1] = pack_bf16(v.x - bf16_lo(packed[j >> 1]), v.y - bf16_hi(packed[j >> 1]));
      if (HAS_STATS) {
        if (j < 16) { s0 = add_f32x2(s0, v); q0 = fma_f32x2(v, v, q0); }
        else        { s1 = add_f32x2(s1, v); q1 = fma_f32x2(v, v, q1); }
      }
    }
  }
  if (HAS_STATS) {
    // GroupNorm partials per 16-channel block over this warp's 32 rows (fp32 accumulator values; the stored bf16
    // differs by rounding noise only). Fixed shuffle tree: deterministic.
    float a0 = s0.x + s0.y, b0 = q0.x + q0.y, a1 = s1.x + s1.y, b1 = q1.x + q1.y;
    const bool trimmed = TRIM && gcol >= p.trim_col && lq + e.lane == p.L - 1;   // half row that does not exist
    if (!valid || trimmed) { a0 = b0 = a1 = b1 = 0.f; }
    a0 = warp_sum(a0); b0 = warp_sum(b0); a1 = warp_sum(a1); b1 = warp_sum(b1);
    if (e.lane == 0)
      *reinterpret_cast<float4*>(p.stats + (stat_chunk * p.stats_nblk + (gcol >> 4)) * 2) = make_float4(a0, b0, a1, b1);
  }
#pragma unroll
  for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {
    const uint32_t* pk = (SPLIT && pl) ? packed_lo : packed;
    __nv_bfloat16* o = (SPLIT && pl) ? reinterpret_cast<__nv_bfloat16*>(p.out_lo) : out;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(e.mine + j * 16, make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rrow = i * 8 + e.srow;
      const bool keep = !(TRIM && lq + rrow == p.L - 1 && gcol + e.spiece * 8 >= p.trim_col);
      if (lq + rrow < p.L && keep && !DV_DBG(p, 8))
        *reinterpret_cast<uint4*>(o + obase + static_cast<size_t>(lq + rrow) * p.ld_out + gcol + e.spiece * 8) =
            lds128(e.coal + i * 8 * EPI_PITCH);
    }
    __syncwarp();
  }
}

// GroupNorm coefficients of one batch item for the operand transform: table[c] = a_c, table[XF_MAXC + c] = b_c with
// y = SiLU?(a_c x + b_c). Statistics: the same partial sums, in the same fixed order, as gn_apply (csrc/norm.cu): one warp
// per group, (block, chunk, fold) flattened over the lanes, fp64 accumulation, fixed shuffle tree.
__device__ __forceinline__ void xf_coefficients(const GemmXf& x, int b, float* table, float* gstat, int xw, int lane, int xtid) {
  const int cpg = x.C / x.G;
  for (int g = xw; g < x.G; g += XF_WARPS) {
    double sum = 0.0, sumsq = 0.0;
    const int kb0 = (g * cpg) >> 4, nkb = cpg >> 4;
    int cov0 = 0;
    for (int si = 0; si < x.nst; ++si) {
      const int lo = max(kb0, cov0 >> 4), hi = min(kb0 + nkb, (cov0 + x.st_cover[si]) >> 4);
      if (hi > lo) {
        const int nf = x.st_fold[si] ? 2 : 1;
        const int per_blk = x.st_chunks[si] * nf;
        const int items = (hi - lo) * per_blk;
        const int nblk = x.st_nblk[si];
        const float* base = x.st_ptr[si] + static_cast<size_t>(b) * x.st_chunks[si] * nblk * 2;
        for (int i0 = 0; i0 < items; i0 += 128) {
          float2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            v[u] = make_float2(0.f, 0.f);
            if (i < items) {
              const int kl = lo - (cov0 >> 4) + i / per_blk;
              const int rem = i % per_blk;
              const int ch = rem / nf, f = rem % nf;
              v[u] = __ldg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(ch) * nblk + kl + f * (nblk >> 1)) * 2));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) { sum += static_cast<double>(v[u].x); sumsq += static_cast<double>(v[u].y); }
        }
      }
      cov0 += x.st_cover[si];
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, m);
      sumsq += __shfl_xor_sync(0xffffffffu, sumsq, m);
    }
    if (lane == 0) {
      const double n = static_cast<double>(x.L) * cpg;
      const double mean = sum / n;
      double var = sumsq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      gstat[2 * g] = static_cast<float>(mean);
      gstat[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(x.eps)));
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(XF_WARPS * 32) : "memory");
  for (int c = xtid; c < x.C; c += XF_WARPS * 32) {
    const int g = c / cpg;
    float a = gstat[2 * g + 1] * __ldg(x.gamma + c);
    float bb = __ldg(x.beta + c) - gstat[2 * g] * a;
    if (x.scale_shift) {
      const float s1 = 1.0f + __ldg(x.scale_shift + static_cast<size_t>(b) * x.ss_ld + c);
      a *= s1;
      bb = bb * s1 + __ldg(x.scale_shift + static_cast<size_t>(b) * x.ss_ld + x.C + c);
    }
    table[c] = a;
    table[XF_MAXC + c] = bb;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(XF_WARPS * 32) : "memory");
}

// TRIM: instantiation for the odd-length folded upsample conv (output batch stride != L * ld_out, last half row absent)
// SPLIT: FP32-class precision (hi + lo planes); XF: operand-transform warps present (GroupNorm fused into the consumer)
template <int BN, int EW, bool TRIM, bool SPLIT = false, bool XF = false>
__global__ void __launch_bounds__(96 + EW * 32 + (XF ? XF_WARPS * 32 : 0), EW == 8 ? 2 : 1)
gemm_tc2_kernel(const __grid_constant__ GemmParams p) {
  using SL = SmemLayout<BN, EW>;
  constexpr int PARTS = EW / 4;                       // epilogue warps per TMEM lane quarter
  constexpr int NA = SL::NA, NW = SL::NW;
  static_assert(!(XF && EW == 8), "the transform warps exist only in the one-CTA-per-SM shape");
  constexpr int WPROD = 2 + EW + (XF ? XF_WARPS : 0);   // last warp: weight producer
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // A-slot barriers. A raw panel: TMA -> a_full -> MMA. A normalised (xf) panel: TMA -> x_full -> transform warps -> a_ready
  // -> MMA. The two kinds complete DIFFERENT barriers of a slot, so each barrier's phases count only its own kind of use and
  // every waiter (which keeps one phase bit per slot and kind) stays exactly in step with it - the transform warps never
  // have to observe, or skip over, the raw panels.
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* x_full = a_full + NA;
  uint64_t* a_ready = x_full + NA;
  uint64_t* a_empty = a_ready + NA;
  uint64_t* w_full = a_empty + NA;
  uint64_t* w_empty = w_full + NW;
  uint64_t* tmem_full_bar = w_empty + NW;         // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  static_assert((4 * NA + 2 * NW + 4) * 8 + 4 <= 512, "barrier region");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles_n = p.n_pad / BN;
  const int num_tiles = p.tiles_per_batch * p.B * n_tiles_n;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GEMM_MAX_SRC; ++s) tma_prefetch_desc(&p.mapA[s]);
    tma_prefetch_desc(&p.mapW);
    tma_prefetch_desc(&p.mapW2);
    for (int s = 0; s < NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&x_full[s], 1);
      mbar_init(&a_ready[s], XF_WARPS * 32);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NW; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EW);
    }
    fence_barrier_init();
  }
  constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;   // power of two >= the two accumulators
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (warp >= 2 && warp < 2 + EW && p.bias) {        // bias vector -> shared memory, once per CTA
    float* sb = reinterpret_cast<float*>(smem + SL::BIAS_OFF);
    for (int i = threadIdx.x - 64; i < p.n_pad; i += EW * 32) sb[i] = __ldg(p.bias + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only static data (weights, descriptors); activations of earlier kernels are read only after
  // griddepcontrol.wait. The producer first puts the W panels of its first ring of slots in flight (weights are static):
  // with PDL that load latency overlaps the previous kernel's tail instead of following it.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer, activations (A ring)
    pdl_wait();
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mi = tile / n_tiles_n;
        const int b = mi / p.tiles_per_batch;
        const int l0 = (mi % p.tiles_per_batch) * BM;
        trace_ev(p, 0, tile, 0);
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          const CUtensorMap* ma = &p.mapA[sg.src];
          const uint32_t a_bytes = static_cast<uint32_t>(p.map_rows[sg.src]) * (BK * 2);
          for (int q = 0; q < sg.npanels; ++q) {
            mbar_wait(&a_empty[as], aph ^ 1);
            uint64_t* landed = (XF && sg.xf) ? &x_full[as] : &a_full[as];
            if (DV_DBG(p, 4)) {
              mbar_arrive(landed);
            } else {
              mbar_arrive_expect_tx(landed, a_bytes);
              tma_load_3d(smem + as * SL::A_SLOT, ma, landed, sg.c0 + q * BK, l0 + sg.row_off, b);
            }
            if (++as == NA) { as = 0; aph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == WPROD) {
    // ------------------------------------------------------------------ TMA producer, weights (W ring). Weights are static
    // data: this warp never waits on the previous kernel (no griddepcontrol.wait), so with PDL its first ring of panels is in
    // flight while the previous kernel drains, and it is never held up by the A ring (its own warp, its own barriers).
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles_n) * BN;
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          const CUtensorMap* mw = sg.wsel ? &p.mapW2 : &p.mapW;
          for (int q = 0; q < sg.npanels; ++q)
            for (int t = 0; t < sg.ntap; ++t) {
              mbar_wait(&w_empty[ws], wph ^ 1);
              if (DV_DBG(p, 4)) {
                mbar_arrive(&w_full[ws]);
              } else {
                uint8_t* sb = smem + SL::W_OFF + ws * SL::W_SLOT;
                mbar_arrive_expect_tx(&w_full[ws], SL::W_SLOT);
                // the W map's box is 64 x 64 rows: a BN-wide tile takes BN / 64 boxes, stacked contiguously
#pragma unroll
                for (int h = 0; h < BN / 64; ++h)
                  tma_load_2d(sb + h * (64 * BK * 2), mw, &w_full[ws], (sg.wk0 + t * sg.tap_stride + q) * BK, n0 + h * 64);
              }
              if (++ws == NW) { ws = 0; wph ^= 1; }
            }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    pdl_wait();
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int as = 0, ws = 0;
      uint32_t full_ph = 0, ready_ph = 0, wph = 0;        // per-slot phase bits of a_full / a_ready (bit = slot)
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        trace_ev(p, 1, tile, 0);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        trace_ev(p, 1, tile, 1);
        const uint32_t tmem_d = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          for (int q = 0; q < sg.npanels; ++q) {
            if (XF && sg.xf) {
              mbar_wait(&a_ready[as], (ready_ph >> as) & 1);
              ready_ph ^= 1u << as;
            } else {
              mbar_wait(&a_full[as], (full_ph >> as) & 1);
              full_ph ^= 1u << as;
            }
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + as * SL::A_SLOT);
            for (int t = 0; t < sg.ntap; ++t) {
              mbar_wait(&w_full[ws], wph);
              tc_fence_after();
              // tap t reads rows t .. t + 127 of the slot: start address + t * 128 B (the 128-byte swizzle is a function of
              // the absolute shared-memory address, so a row-shifted start needs nothing else: r2b_umma_row_shift_probe)
              const uint64_t adesc = umma_desc_sw128(sa + t * (BK * 2));
              const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + SL::W_OFF + ws * SL::W_SLOT));
              if (!DV_DBG(p, 2)) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
                  umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit(&w_empty[ws]);
              if (++ws == NW) { ws = 0; wph ^= 1; }
            }
            umma_commit(&a_empty[as]);
            if (++as == NA) as = 0;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        trace_ev(p, 1, tile, 2);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (XF && warp >= 2 + EW && warp < WPROD) {
    // ------------------------------------------------------------------ operand transform: GroupNorm (+SiLU) in shared memory
    pdl_wait();                                          // the statistics partials come from the producing kernels
    const int xw = warp - (2 + EW), xtid = threadIdx.x - (2 + EW) * 32;
    float* table = reinterpret_cast<float*>(smem + SL::XF_OFF);
    float* gstat = table + 2 * XF_MAXC;
    int as = 0;
    uint32_t xph = 0;                                    // per-slot phase bits of x_full
    int cur_b = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mi = tile / n_tiles_n;
      const int b = mi / p.tiles_per_batch;
      const int l0 = (mi % p.tiles_per_batch) * BM;
      if (b != cur_b && !(DV_DBG(p, 256) && cur_b >= 0)) {   // dbg 256: coefficients once per CTA (timing only)
        xf_coefficients(p.xf, b, table, gstat, xw, lane, xtid);
        cur_b = b;
      }
      for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg sg = p.seg[s];
        const int rows = p.map_rows[sg.src];
        for (int q = 0; q < sg.npanels; ++q) {
          if (sg.xf) {
            mbar_wait(&x_full[as], (xph >> as) & 1);
            xph ^= 1u << as;
            const uint32_t slot = smem_u32(smem + as * SL::A_SLOT);
            const int cbase = sg.xf_c0 + q * BK;
            // 16-byte pieces: piece jp of row r holds channels 8 * (jp ^ (r & 7)) .. + 8 of the panel (128-byte swizzle)
            for (int i = xtid; i < rows * 8; i += XF_WARPS * 32) {
              const int r = i >> 3, jp = i & 7;
              const int l = l0 + sg.row_off + r;
              if (l < 0 || l >= p.xf.L) continue;         // rows outside the tensor stay zero: the conv padding
              const int c = cbase + ((jp ^ (r & 7)) << 3);
              const uint32_t addr = slot + r * (BK * 2) + jp * 16;
              const uint4 v = lds128(addr);
              const float4 a0 = *reinterpret_cast<const float4*>(table + c), a1 = *reinterpret_cast<const float4*>(table + c + 4);
              const float4 b0 = *reinterpret_cast<const float4*>(table + XF_MAXC + c);
              const float4 b1 = *reinterpret_cast<const float4*>(table + XF_MAXC + c + 4);
              float y[8];
              y[0] = fmaf(bf16_lo(v.x), a0.x, b0.x); y[1] = fmaf(bf16_hi(v.x), a0.y, b0.y);
              y[2] = fmaf(bf16_lo(v.y), a0.z, b0.z); y[3] = fmaf(bf16_hi(v.y), a0.w, b0.w);
              y[4] = fmaf(bf16_lo(v.z), a1.x, b1.x); y[5] = fmaf(bf16_hi(v.z), a1.y, b1.y);
              y[6] = fmaf(bf16_lo(v.w), a1.z, b1.z); y[7] = fmaf(bf16_hi(v.w), a1.w, b1.w);
              if (p.xf.silu && !DV_DBG(p, 128)) {           // dbg 128: no SiLU (timing only)
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = silu_f(y[j]);
              }
              sts128(addr, make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7])));
            }
            fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(&a_ready[as]);
          }
          if (++as == NA) as = 0;
        }
      }
    }
  } else {
    pdl_wait();
    // ------------------------------------------------------------------ epilogue: warps 2..17
    // TMEM lane quarter is fixed by hardware to (warp id % 4); the four warps of a quarter split the columns.
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;                    // 0..PARTS-1
    EpiCtx e;
    const uint32_t slab = smem_u32(smem + SL::EPI_OFF + (warp - 2) * EPI_SLAB);
    e.lane = lane; e.srow = lane >> 2; e.spiece = lane & 3;
    e.mine = slab + lane * EPI_PITCH;
    e.coal = slab + e.srow * EPI_PITCH + e.spiece * 16;
    e.sbias = smem_u32(smem + SL::BIAS_OFF);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mi = tile / n_tiles_n;
      const int n0 = (tile % n_tiles_n) * BN;
      const int b = mi / p.tiles_per_batch;
      const int mt = mi % p.tiles_per_batch;
      const int lq = mt * BM + quad * 32;               // first row of this warp's 32-row slab
      const int l = lq + lane;                          // this thread's accumulator row
      const bool valid = l < p.L;
      if (warp == 2 && lane == 0) trace_ev(p, 2, tile, 0);
      const size_t rowbase = static_cast<size_t>(b) * p.L;
      const size_t obase = TRIM ? static_cast<size_t>(b) * p.out_bs : rowbase * p.ld_out;
      uint4 rv[4];
      bool preloaded = false;
      if (p.out_mode == GEMM_OUT_BF16 && p.res != nullptr && part * 32 < BN && n0 + part * 32 < p.N && !DV_DBG(p, 1 | 32)) {
        epi_load_res(p, e, n0 + part * 32, lq, rowbase, rv);
        preloaded = true;
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (warp == 2 && lane == 0) trace_ev(p, 2, tile, 1);
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);

      if (DV_DBG(p, 1)) {
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      } else if (p.out_mode == GEMM_OUT_GEGLU) {
        // tile columns repeat [u 64 | gate 64]; work unit = 16 u columns + their 16 gate columns -> 16 outputs
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        constexpr int NUNIT = BN / 32;                   // 4 per 128 tile columns
        const int orow = lane >> 1, opiece = lane & 1;   // row-contiguous phase: 16 rows x 2 pieces of 16 B
#pragma unroll 1
        for (int un = part; un < NUNIT; un += PARTS) {
          const int cu = (un >> 2) * 128 + (un & 3) * 16;
          const int ocol = (n0 >> 1) + (un >> 2) * 64 + (un & 3) * 16;
          uint32_t u[16], g[16];
          tmem_ld16(taddr + cu, u);
          tmem_ld16(taddr + cu + 64, g);
          tmem_ld_wait();
          if (un + PARTS >= NUNIT) {                          // last TMEM read of this tile by this warp
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          uint32_t packed[8];
          uint32_t packed_lo[SPLIT ? 8 : 1];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 bu = make_float4(0.f, 0.f, 0.f, 0.f), bg = bu;
            if (p.bias) {
              bu = lds128f(e.sbias + (n0 + cu + 4 * j4) * 4);
              bg = lds128f(e.sbias + (n0 + cu + 64 + 4 * j4) * 4);
            }
            const int j = 4 * j4;
            const float2 ua = add_f32x2(make_float2(__uint_as_float(u[j]), __uint_as_float(u[j + 1])), make_float2(bu.x, bu.y));
            const float2 ub = add_f32x2(make_float2(__uint_as_float(u[j + 2]), __uint_as_float(u[j + 3])), make_float2(bu.z, bu.w));
            const float2 ga = add_f32x2(make_float2(__uint_as_float(g[j]), __uint_as_float(g[j + 1])), make_float2(bg.x, bg.y));
            const float2 gb = add_f32x2(make_float2(__uint_as_float(g[j + 2]), __uint_as_float(g[j + 3])), make_float2(bg.z, bg.w));
            if (SPLIT) {                                     // FP32-class mode: exact erf GELU, hi + lo planes
              const float h0 = ua.x * gelu_erf_f(ga.x), h1 = ua.y * gelu_erf_f(ga.y);
              const float h2 = ub.x * gelu_erf_f(gb.x), h3 = ub.y * gelu_erf_f(gb.y);
              packed[2 * j4] = pack_bf16(h0, h1);
              packed[2 * j4 + 1] = pack_bf16(h2, h3);
              packed_lo[2 * j4] = pack_bf16(h0 - bf16_lo(packed[2 * j4]), h1 - bf16_hi(packed[2 * j4]));
              packed_lo[2 * j4 + 1] = pack_bf16(h2 - bf16_lo(packed[2 * j4 + 1]), h3 - bf16_hi(packed[2 * j4 + 1]));
            } else {
              packed[2 * j4] = pack_bf16(ua.x * gelu_fast_f(ga.x), ua.y * gelu_fast_f(ga.y));
              packed[2 * j4 + 1] = pack_bf16(ub.x * gelu_fast_f(gb.x), ub.y * gelu_fast_f(gb.y));
            }
          }
#pragma unroll
          for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {
            const uint32_t* pk = (SPLIT && pl) ? packed_lo : packed;
            __nv_bfloat16* o = (SPLIT && pl) ? reinterpret_cast<__nv_bfloat16*>(p.out_lo) : out;
            sts128(e.mine, make_uint4(pk[0], pk[1], pk[2], pk[3]));
            sts128(e.mine + 16, make_uint4(pk[4], pk[5], pk[6], pk[7]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r = i * 16 + orow;
              if (lq + r < p.L)
                *reinterpret_cast<uint4*>(o + (rowbase + lq + r) * p.ld_out + ocol + opiece * 8) =
                    lds128(slab + r * EPI_PITCH + opiece * 16);
            }
            __syncwarp();
          }
        }
      } else if (p.out_mode == GEMM_OUT_F32_NCL) {
        float* out = reinterpret_cast<float*>(p.out) + static_cast<size_t>(b) * p.N * p.L + (valid ? l : 0);
        constexpr int NCH = BN / 32;
#pragma unroll 1
        for (int ch = part; ch < NCH; ch += PARTS) {
          const int c = ch * 32;
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          if (ch + PARTS >= NCH) {
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bq = lds128f(e.sbias + (n0 + c + 4 * j4) * 4);
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int n = n0 + c + 4 * j4 + h;
              if (valid && n < p.N) out[static_cast<size_t>(n) * p.L] = __uint_as_float(r[4 * j4 + h]) + bb[h];
            }
          }
        }
      } else {
        constexpr int NCH = BN / 32;
        const size_t stat_chunk = (static_cast<size_t>(b) * p.tiles_per_batch + mt) * 4 + quad;
        const bool has_res = p.res != nullptr, has_stats = p.stats != nullptr;
        bool arrived = false;
#pragma unroll 1
        for (int ch = part; ch < NCH; ch += PARTS) {
          const int gcol = n0 + ch * 32;
          const bool last = ch + PARTS >= NCH;
          if (gcol >= p.N) continue;                      // padded output columns are neither read nor written
          arrived |= last;
          const uint32_t tc = taddr + ch * 32;
          if (has_res) {
            if (has_stats) epi_chunk_bf16<true, true, TRIM, SPLIT>(p, e, tc, gcol, lq, rowbase, obase, valid, stat_chunk, last, &tmem_empty_bar[acc], rv, preloaded && ch == part);
            else           epi_chunk_bf16<true, false, TRIM, SPLIT>(p, e, tc, gcol, lq, rowbase, obase, valid, stat_chunk, last, &tmem_empty_bar[acc], rv, preloaded && ch == part);
          } else {
            if (has_stats) epi_chunk_bf16<false, true, TRIM, SPLIT>(p, e, tc, gcol, lq, rowbase, obase, valid, stat_chunk, last, &tmem_empty_bar[acc], rv, preloaded && ch == part);
            else           epi_chunk_bf16<false, false, TRIM, SPLIT>(p, e, tc, gcol, lq, rowbase, obase, valid, stat_chunk, last, &tmem_empty_bar[acc], rv, preloaded && ch == part);
          }
        }
        if (!arrived) {                                   // this warp's columns were all padding
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
      }
      if (warp == 2 && lane == 0) trace_ev(p, 2, tile, 2);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int g_num_sms = 0;
int g_force_ew = 0;   // DV_GEMM_EW: 8 / 16 force the CTA shape of the 128-wide tile kernel (0 = heuristic)

template <int BN, int EW, bool TRIM = false, bool SPLIT = false, bool XF = false>
cudaError_t launch_bn(const GemmParams& p, cudaStream_t st) {
  using SL = SmemLayout<BN, EW>;
  if (SL::dyn_bytes(p.n_pad) > SL::max_bytes) return cudaErrorInvalidValue;
  const int tiles = p.tiles_per_batch * p.B * (p.n_pad / BN);
  const int slots = g_num_sms * (EW == 8 ? 2 : 1);
  const int grid = tiles < slots ? tiles : slots;
  return launch_k(gemm_tc2_kernel<BN, EW, TRIM, SPLIT, XF>, dim3(grid), dim3(96 + EW * 32 + (XF ? XF_WARPS * 32 : 0)),
                  SL::dyn_bytes(p.n_pad), st, p);
}

template <int BN, int EW, bool TRIM = false, bool SPLIT = false, bool XF = false>
cudaError_t set_smem_attr() {
  using SL = SmemLayout<BN, EW>;
  const int want = SL::dyn_bytes(MAX_NPAD);
  return cudaFuncSetAttribute(gemm_tc2_kernel<BN, EW, TRIM, SPLIT, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              want < SL::max_bytes ? want : SL::max_bytes);
}

}  // namespace

cudaError_t gemm_tc2_init() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  {
    const char* ev = getenv("DV_GEMM_EW");
    if (ev) g_force_ew = atoi(ev);
  }
  cudaError_t e;
#define DV_SET(...) if ((e = set_smem_attr<__VA_ARGS__>()) != cudaSuccess) return e;
  DV_SET(128, 16) DV_SET(128, 8) DV_SET(192, 16) DV_SET(64, 16) DV_SET(256, 16)
  DV_SET(128, 16, true) DV_SET(256, 16, true)
  DV_SET(128, 16, false, true) DV_SET(128, 16, true, true)
  DV_SET(64, 16, false, false, true) DV_SET(128, 16, false, false, true) DV_SET(192, 16, false, false, true)
  DV_SET(256, 16, false, false, true)
#undef DV_SET
  return cudaSuccess;
}

cudaError_t launch_gemm_tc2(const GemmParams& p_in, cudaStream_t st) {
  GemmParams p = p_in;
  if (p.out_bs == 0) p.out_bs = static_cast<long long>(p.L) * p.ld_out;
  if (p.trim_col % 32 != 0) return cudaErrorInvalidValue;
  if (g_num_sms == 0) {
    cudaError_t e = gemm_tc2_init();
    if (e != cudaSuccess) return e;
  }
  if (p.n_pad % 128 != 0 || p.n_pad > MAX_NPAD || p.nseg < 1 || p.nseg > GEMM_MAX_SEG) return cudaErrorInvalidValue;
  if (p.out_mode == GEMM_OUT_GEGLU && p.n_pad % 128 != 0) return cudaErrorInvalidValue;
  bool has_xf = false;
  for (int i = 0; i < p.nseg; ++i) {
    const GemmSeg& sg = p.seg[i];
    if (sg.ntap != 1 && sg.ntap != 3) return cudaErrorInvalidValue;
    if (p.map_rows[sg.src] < (sg.ntap == 3 ? 130 : 128) || p.map_rows[sg.src] > A_ROWS_MAX) return cudaErrorInvalidValue;
    has_xf |= sg.xf != 0;
  }
  if (has_xf && (p.xf.C > XF_MAXC || p.xf.C % (16 * p.xf.G) != 0 || p.xf.G > 32 || p.xf.nst < 1 || p.xf.nst > 2))
    return cudaErrorInvalidValue;
  const bool dense = p.out_bs == static_cast<long long>(p.L) * p.ld_out;
  if (p.out_lo != nullptr) {                             // split precision: one tile shape (128 x 128, 16 epilogue warps)
    if (has_xf) return cudaErrorInvalidValue;            // (the FP32-class mode keeps the stand-alone norm kernels)
    if (p.res != nullptr && p.res_lo == nullptr) return cudaErrorInvalidValue;
    if (p.trim_col > 0 || !dense) {
      if (p.out_mode != GEMM_OUT_BF16) return cudaErrorInvalidValue;
      return launch_bn<128, 16, true, true>(p, st);
    }
    return launch_bn<128, 16, false, true>(p, st);
  }
  const int bn = p.bn ? p.bn : gemm_pick_bn(p.tiles_per_batch * p.B, p.n_pad);
  if (p.trim_col > 0 || !dense) {                        // odd-length folded upsample conv: its own instantiations
    if (p.out_mode != GEMM_OUT_BF16 || has_xf) return cudaErrorInvalidValue;
    if (bn == 256 && p.n_pad % 256 == 0) return launch_bn<256, 16, true>(p, st);
    return launch_bn<128, 16, true>(p, st);
  }
  if (bn == 256) {
    if (p.n_pad % 256 != 0) return cudaErrorInvalidValue;
    return has_xf ? launch_bn<256, 16, false, false, true>(p, st) : launch_bn<256, 16>(p, st);
  }
  if (bn == 192) {
    if (p.n_pad % 192 != 0 || p.out_mode == GEMM_OUT_GEGLU) return cudaErrorInvalidValue;
    return has_xf ? launch_bn<192, 16, false, false, true>(p, st) : launch_bn<192, 16>(p, st);
  }
  if (bn == 64 && p.out_mode != GEMM_OUT_GEGLU)          // GEGLU pairs u / gate inside 128 columns
    return has_xf ? launch_bn<64, 16, false, false, true>(p, st) : launch_bn<64, 16>(p, st);
  if (has_xf) return launch_bn<128, 16, false, false, true>(p, st);
  // two small CTAs per SM pay once there is more than one wave of 128-wide tiles (measured on the C2 shapes: 5-10 %
  // faster at levels 0-2, slower for the few long-K tiles of level 3)
  const int tiles = p.tiles_per_batch * p.B * (p.n_pad / 128);
  int ew = g_force_ew ? g_force_ew : (tiles >= g_num_sms + g_num_sms / 8 ? 8 : 16);
  if (ew == 8 && SmemLayout<128, 8>::dyn_bytes(p.n_pad) > SmemLayout<128, 8>::max_bytes) ew = 16;
  return ew == 8 ? launch_bn<128, 8>(p, st) : launch_bn<128, 16>(p, st);
}

}  // namespace DV_NS

// Tap-GEMM on the 5th-gen tensor cores: implicit-GEMM conv1d (k=3, stride 1 or 2, nearest-2x upsample folded
// into the weight packing), 1x1 conv and Linear, all as   D[128 x BN] += A_panel[128 x 64] . W_panel[BN x 64]^T.
//
// Persistent, warp-specialised kernel: one CTA per SM loops over output tiles (static round-robin schedule).
//   warp 0    : TMA producer  - one elected lane issues cp.async.bulk.tensor loads (A: 3-D [C, L, B] box 64x128x1,
//                               W: 2-D [K, N] box 64xBN), 128-byte swizzle, completion on the stage's "full" mbarrier.
//                               Conv taps are the same box shifted by -1/0/+1 rows; rows outside [0, L) are zero
//                               filled by TMA, which is exactly the conv padding. torch.cat inputs are extra A maps.
//   warp 1    : MMA issuer    - lane 0 issues tcgen05.mma.cta_group::1.kind::f16 (BF16 x BF16 -> FP32 in TMEM),
//                               4 x (K=16) per 64-wide panel; tcgen05.commit releases the smem stage / hands the
//                               accumulator to the epilogue. Two TMEM accumulators (2 x BN columns) ping-pong, so
//                               the epilogue of tile i overlaps the main loop of tile i+1.
//   warps 2-9 : epilogue      - tcgen05.ld 32 lanes x 32 columns -> registers -> (+bias) (+residual) (GEGLU) -> bf16,
//                               staged through a per-warp smem slab so that global loads/stores are row-contiguous
//                               (64-byte segments, 8 rows per instruction) instead of one row per thread; optional
//                               per-(32-row chunk, 16-channel block) sum / sum-of-squares of the stored values for the
//                               GroupNorm that consumes this tensor (deterministic shuffle tree, no atomics);
//                               conv_out stores fp32 [B, N, L] directly (lanes = consecutive frames).
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

template <int BN, int EW>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = EW == 8 ? 2 : (BN == 64 ? 6 : (BN == 128 ? 5 : (BN == 192 ? 4 : 3)));
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = EPI_OFF + EW * EPI_SLAB;
  static constexpr int GNS_OFF = BAR_OFF + 256;        // barriers + tmem slot, then mean / rstd of up to 32 groups (fused output GroupNorm)
  static constexpr int BIAS_OFF = GNS_OFF + 256;       // then the bias vector (n_pad floats)
  static constexpr int dyn_bytes(int n_pad) { return BIAS_OFF + n_pad * 4 + 1024; }   // + slack for 1024-byte alignment
  static_assert(dyn_bytes(MAX_NPAD) <= (EW == 8 ? 116224 : 232448), "exceeds the shared memory budget");
};

// debug timeline (debug library only, enabled when p.trace != nullptr): CTA 0; every role (0 producer thread, 1 MMA thread,
// 2 lane 0 of the first epilogue warp) appends (event, tile, SM clock) records to its own region of the buffer with plain
// stores (no atomics: the probe must not add a round trip to what it measures); role 3 holds %globaltimer at kernel entry /
// after griddepcontrol.wait / at exit, which places consecutive launches on one time axis.
constexpr int TRACE_REGION = 160;
__device__ __forceinline__ void trace_ev(const GemmParams& p, int role, int tile, int ev) {
#ifdef DV_DEBUG_HOOKS
  if (p.trace && blockIdx.x == 0) {
    unsigned long long* reg = p.trace + 1 + role * TRACE_REGION;
    const unsigned long long i = reg[0];
    if (i + 1 < TRACE_REGION) {
      unsigned long long t;
      if (role == 3) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); else t = clock64();
      reg[1 + i] = (static_cast<unsigned long long>(role) << 60) | (static_cast<unsigned long long>(tile & 0xFFFF) << 44) |
                   (static_cast<unsigned long long>(ev & 0xF) << 40) | (t & 0xFFFFFFFFFFULL);
      reg[0] = i + 1;
    }
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
  bool trace_warp;  // debug timeline: this is the warp that records epilogue events
  uint32_t stash;   // fused output GroupNorm: shared-memory copy of the chunk being stored (row-contiguous layout), or 0
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
  if (e.trace_warp && e.lane == 0) trace_ev(p, 2, gcol, 6);      // TMEM load complete
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
  // this thread's accumulator row is l = lq + lane of batch item rowbase / L: masked rows are stored as zeros
  const bool row_keep = !TRIM || p.row_len == nullptr || (lq + e.lane) < __ldg(p.row_len + rowbase / p.L);
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
      if (TRIM) {                                         // rare-path instantiation: also ReLU / padding-row mask (prompt encoder)
        if (p.relu) v = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
        if (!row_keep) v = make_float2(0.f, 0.f);
      }
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
  if (e.trace_warp && e.lane == 0) trace_ev(p, 2, gcol, 7);      // math done
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
      const uint4 val = lds128(e.coal + i * 8 * EPI_PITCH);
      if (lq + rrow < p.L && keep && !DV_DBG(p, 8))
        *reinterpret_cast<uint4*>(o + obase + static_cast<size_t>(lq + rrow) * p.ld_out + gcol + e.spiece * 8) = val;
      if (!SPLIT && !TRIM && e.stash) sts128(e.stash + (i * 32 + e.lane) * 16, val);   // kept on chip for the normalisation pass
    }
    __syncwarp();
  }
  if (e.trace_warp && e.lane == 0) trace_ev(p, 2, gcol, 8);      // stores issued
}


// ---- GroupNorm of the output applied by the producing CTAs (GemmGnOut)
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void epi_bar_sync(int nthreads) {   // named barrier 1: the epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
// One thread per CTA: arrive on the batch item's counter, wait for all `S` tiles of the item, depart; the last departure
// resets the counter for the next launch. The wait is bounded (~0.1 s) so that a scheduling assumption that does not hold
// produces wrong numbers (caught by the parity tests) instead of a hung GPU.
__device__ __forceinline__ void gn_item_arrive_wait(unsigned* c, unsigned S, unsigned* gave_up) {
  // release: (after the CTA barrier) the tile's raw store and partials are visible before the arrival is. A release RMW instead
  // of __threadfence() + atomicAdd: the latter compiles to MEMBAR.SC.GPU, the sequentially-consistent fence.
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
  for (int spins = 0; spins < (1 << 17); ++spins) {   // ~0.1 s at the L2 round-trip rate; an item completes within ~50 us
    if (ld_acquire_u32(c) >= S) return;
    if (spins > 4) __nanosleep(40);
  }
  atomicExch(gave_up, 1u);            // sticky: the host side refuses the plan's results (engine.py check_finite / sample_diag)
}
__device__ __forceinline__ void gn_item_depart(unsigned* c, unsigned S) {   // off the critical path: after the CTA barrier
  if (atomicAdd(c, 1u) == 2u * S - 1u) atomicExch(c, 0u);
}
// Sum / sum of squares of one group of one batch item from the producers' partials [chunks][nblk][2]: the same flattened
// lane assignment, fp64 accumulation order and shuffle tree as gn_apply_kernel (csrc/norm.cu), so both give the same bits.
__device__ __forceinline__ void gn_group_fold(const float* base, int chunks, int nblk, int kb0, int nkb, int lane, double& sum,
                                              double& sumsq) {
  sum = 0.0; sumsq = 0.0;
  const int items = nkb * chunks;
  for (int i0 = 0; i0 < items; i0 += 128) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 32 + lane;
      v[u] = make_float2(0.f, 0.f);
      if (i < items) {
        const int kl = i / chunks, ch = i % chunks;
        v[u] = __ldcg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(ch) * nblk + kb0 + kl) * 2));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { sum += static_cast<double>(v[u].x); sumsq += static_cast<double>(v[u].y); }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, m);
    sumsq += __shfl_xor_sync(0xffffffffu, sumsq, m);
  }
}
// Normalise one 32-row x 32-column chunk of the tile this warp has just stored: read back in the row-contiguous layout
// (each lane: 8 columns of 4 rows), y = SiLU?(x * ca + cb) with the coefficients built exactly like gn_apply_kernel's.
// SPLIT: the raw tensor is hi + lo (both planes were just stored by this warp); the normalised tensor gets a lo plane where
// the mode carries one for GEMM operands (p.gn.out_lo, FP32X3) - the same arithmetic as gn_apply_kernel<true>.
template <bool SPLIT>
__device__ __forceinline__ void gn_apply_chunk(const GemmParams& p, const EpiCtx& e, int gcol, int lq, int b, size_t obase,
                                               const float* s_mean, const float* s_rstd, int gfirst, int cpg, uint32_t stash,
                                               uint32_t coef, int n0) {
  const int c8 = gcol + e.spiece * 8;
  if (c8 >= p.N) return;
  const __nv_bfloat16* raw = reinterpret_cast<const __nv_bfloat16*>(p.out);
  const __nv_bfloat16* raw_lo = reinterpret_cast<const __nv_bfloat16*>(p.out_lo);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.gn.out);
  uint4 v[4];
  uint4 vl[SPLIT ? 4 : 1];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = lq + i * 8 + e.srow;
    if (!SPLIT && stash) v[i] = lds128(stash + (i * 32 + e.lane) * 16);      // this warp's own copy from pass 1
    else v[i] = l < p.L ? __ldcg(reinterpret_cast<const uint4*>(raw + obase + static_cast<size_t>(l) * p.ld_out + c8)) : make_uint4(0, 0, 0, 0);
    if (SPLIT)
      vl[i] = l < p.L ? __ldcg(reinterpret_cast<const uint4*>(raw_lo + obase + static_cast<size_t>(l) * p.ld_out + c8)) : make_uint4(0, 0, 0, 0);
  }
  float ca[8], cb[8];
  if (!SPLIT && coef) {
    // per-column y = x * ca + cb of this tile, computed ONCE per column by the CTA (gn_tile_coefficients below) instead of
    // once per lane and chunk behind eight global loads: [BN] ca then [BN] cb floats in shared memory
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 a4 = lds128f(coef + (c8 - n0 + j) * 4);
      const float4 b4 = lds128f(coef + (256 + c8 - n0 + j) * 4);
      ca[j] = a4.x; ca[j + 1] = a4.y; ca[j + 2] = a4.z; ca[j + 3] = a4.w;
      cb[j] = b4.x; cb[j + 1] = b4.y; cb[j + 2] = b4.z; cb[j + 3] = b4.w;
    }
  } else {
  const int grp = c8 / cpg - gfirst;
  const float mean = s_mean[grp], rstd = s_rstd[grp];
#pragma unroll
  for (int j = 0; j < 8; j += 4) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gn.gamma + c8 + j));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.gn.beta + c8 + j));
    const float ga[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = rstd * ga[k];
      ca[j + k] = a;
      cb[j + k] = be[k] - mean * a;
    }
  }
  if (p.gn.scale_shift) {
    const float* ssp = p.gn.scale_shift + static_cast<size_t>(b) * p.gn.ss_ld + c8;
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(ssp + j));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(ssp + p.N + j));
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float s1 = 1.0f + scv[k];
        ca[j + k] *= s1;
        cb[j + k] = cb[j + k] * s1 + shv[k];
      }
    }
  }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = lq + i * 8 + e.srow;
    if (l >= p.L) continue;
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    uint32_t y[4], yl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xa = bf16_lo(w[j]), xc = bf16_hi(w[j]);
      if (SPLIT) {
        const uint32_t wl = j == 0 ? vl[i].x : (j == 1 ? vl[i].y : (j == 2 ? vl[i].z : vl[i].w));
        xa += bf16_lo(wl);
        xc += bf16_hi(wl);
      }
      float a = fmaf(xa, ca[2 * j], cb[2 * j]);
      float c = fmaf(xc, ca[2 * j + 1], cb[2 * j + 1]);
      if (p.gn.silu) { a = silu_f(a); c = silu_f(c); }
      y[j] = pack_bf16(a, c);
      if (SPLIT) yl[j] = pack_bf16(a - bf16_lo(y[j]), c - bf16_hi(y[j]));
    }
    *reinterpret_cast<uint4*>(dst + obase + static_cast<size_t>(l) * p.ld_out + c8) = make_uint4(y[0], y[1], y[2], y[3]);
    if (SPLIT && p.gn.out_lo != nullptr)
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.gn.out_lo) + obase + static_cast<size_t>(l) * p.ld_out + c8) =
          make_uint4(yl[0], yl[1], yl[2], yl[3]);
  }
}

// TRIM: instantiation for the odd-length folded upsample conv (output batch stride != L * ld_out, last half row absent)
template <int BN, int EW, bool TRIM, bool SPLIT = false>
__global__ void __launch_bounds__(64 + EW * 32, EW == 8 ? 2 : 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using SL = SmemLayout<BN, EW>;
  constexpr int PARTS = EW / 4;                       // epilogue warps per TMEM lane quarter
  constexpr int STAGES = SL::STAGES;
  pdl_launch_dependents();
  if (threadIdx.x == 0) { trace_ev(p, 3, 0, 0); trace_ev(p, 0, 0, 8); }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles_n = p.n_pad / BN;
  const int num_tiles = p.tiles_per_batch * p.B * n_tiles_n;
  int total_panels = 0;
  for (int s = 0; s < p.nseg; ++s) total_panels += p.seg[s].npanels;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GEMM_MAX_SRC; ++s) tma_prefetch_desc(&p.mapA[s]);
    tma_prefetch_desc(&p.mapW);
    tma_prefetch_desc(&p.mapW2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EW);
    }
    fence_barrier_init();
  }
  constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;   // power of two >= the two accumulators
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (warp >= 2 && p.bias) {                         // bias vector -> shared memory, once per CTA
    float* sb = reinterpret_cast<float*>(smem + SL::BIAS_OFF);
    for (int i = threadIdx.x - 64; i < p.n_pad; i += EW * 32) sb[i] = __ldg(p.bias + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_ev(p, 0, 0, 9);      // prologue (barriers, TMEM, bias) done
  // everything above touched only static data (weights, descriptors); activations of earlier kernels are read only after
  // griddepcontrol.wait. The producer first puts the W panels of its first ring of stages in flight (weights are static):
  // with PDL that load latency overlaps the previous kernel's tail instead of following it.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int pre = 0;                                        // stages whose expect_tx + W loads are already issued
      if (!DV_DBG(p, 4 | 64)) {
        const int n0 = (blockIdx.x % n_tiles_n) * BN;
        pre = total_panels < STAGES ? total_panels : STAGES;
        int st = 0;
        for (int sgi = 0; sgi < p.nseg && st < pre; ++sgi) {
          const GemmSeg sg = p.seg[sgi];
          const CUtensorMap* mw = sg.wsel ? &p.mapW2 : &p.mapW;
          for (int q = 0; q < sg.npanels && st < pre; ++q, ++st) {
            uint8_t* sb = smem + st * SL::STAGE_BYTES + SL::A_BYTES;
            mbar_arrive_expect_tx(&full_bar[st], SL::STAGE_BYTES);
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sb + h * (64 * BK * 2), mw, &full_bar[st], (sg.wk0 + q) * BK, n0 + h * 64);
          }
        }
      }
      // tile coordinates (two integer divisions by run-time values, a dependent chain of ~300 clk for one thread) are taken
      // BEFORE the wait on the previous kernel for the first tile and one tile ahead afterwards: they sit on the critical path
      // between "previous kernel complete" and the first A load otherwise
      int tile = blockIdx.x;
      int mi = tile / n_tiles_n, n0 = (tile % n_tiles_n) * BN;
      int b = mi / p.tiles_per_batch, l0 = (mi % p.tiles_per_batch) * BM;
      asm volatile("" ::"r"(mi), "r"(n0), "r"(b), "r"(l0));    // (keeps the compiler from sinking them below the wait)
      pdl_wait();
      trace_ev(p, 0, 0, 10);                              // previous kernel complete
      trace_ev(p, 3, 0, 1);
      int stage = 0;
      uint32_t phase = 0;
      int issued = 0;                                     // panels issued by this CTA so far (across tiles)
      for (; tile < num_tiles; tile += gridDim.x, mi = tile / n_tiles_n, n0 = (tile % n_tiles_n) * BN,
                               b = mi / p.tiles_per_batch, l0 = (mi % p.tiles_per_batch) * BM) {
        int kp = 0;
        trace_ev(p, 0, tile, 0);
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          const CUtensorMap* ma = &p.mapA[sg.src];
          const CUtensorMap* mw = sg.wsel ? &p.mapW2 : &p.mapW;
          for (int q = 0; q < sg.npanels; ++q, ++kp, ++issued) {
            uint8_t* sa = smem + stage * SL::STAGE_BYTES;
            uint8_t* sb = sa + SL::A_BYTES;
            if (issued < pre) {                           // first ring: barrier armed and W already in flight
              tma_load_3d(sa, ma, &full_bar[stage], sg.c0 + q * BK, l0 + sg.row_off, b);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (DV_DBG(p, 4)) { mbar_arrive(&full_bar[stage]); if (++stage == STAGES) { stage = 0; phase ^= 1; } continue; }
            mbar_arrive_expect_tx(&full_bar[stage], SL::STAGE_BYTES);
            tma_load_3d(sa, ma, &full_bar[stage], sg.c0 + q * BK, l0 + sg.row_off, b);
            // the W map's box is 64 x 64 rows: a BN-wide tile takes BN / 64 boxes, stacked contiguously
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sb + h * (64 * BK * 2), mw, &full_bar[stage], (sg.wk0 + q) * BK, n0 + h * 64);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else {
      pdl_wait();
    }
    __syncwarp();
  } else if (warp == 1) {
    pdl_wait();
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        trace_ev(p, 1, tile, 0);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        trace_ev(p, 1, tile, 1);
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kp = 0; kp < total_panels; ++kp) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kp == 0) trace_ev(p, 1, tile, 3);           // first panel landed
          const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + SL::A_BYTES);
          if (!DV_DBG(p, 2))
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kp | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        trace_ev(p, 1, tile, 2);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
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
    e.stash = 0;
#ifdef DV_DEBUG_HOOKS
    e.trace_warp = warp == 2 && p.trace != nullptr;
#else
    e.trace_warp = false;
#endif
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
        // Fused output GroupNorm on the LAST tile of this CTA (every tile, when the launch is a single wave): the operand ring
        // is idle by now (all panels consumed: the accumulator is complete), so the stored 16-bit tile is also kept there,
        // 2 KB per (warp, chunk) in the row-contiguous layout, and the normalisation pass reads it back from shared memory
        // instead of from L2 behind its own stores. Earlier tiles of a multi-tile CTA re-read from global.
        constexpr int STASH_K = (NCH + PARTS - 1) / PARTS;
        static_assert(EW * STASH_K * 2048 <= SL::STAGES * SL::STAGE_BYTES, "stash does not fit the operand ring");
        const bool use_stash = !TRIM && !SPLIT && p.gn.out != nullptr && !p.gn.no_stash && tile + static_cast<int>(gridDim.x) >= num_tiles;
        const uint32_t stash_base = smem_u32(smem) + static_cast<uint32_t>((warp - 2) * STASH_K) * 2048u;
#pragma unroll 1
        for (int ch = part; ch < NCH; ch += PARTS) {
          const int gcol = n0 + ch * 32;
          const bool last = ch + PARTS >= NCH;
          if (gcol >= p.N) continue;                      // padded output columns are neither read nor written
          arrived |= last;
          e.stash = use_stash ? stash_base + static_cast<uint32_t>(ch / PARTS) * 2048u : 0u;
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
        if (!TRIM && p.gn.out != nullptr) {
          // ---- GroupNorm of the output, applied here: wait for the other tiles of this batch item, then normalise this tile
          float* s_mean = reinterpret_cast<float*>(smem + SL::GNS_OFF);
          float* s_rstd = s_mean + 32;
          // with the stash (idle operand ring) the CTA also keeps a per-column coefficient table there: thread `etid` owns
          // column n0 + etid and fetches its affine / time-modulation parameters NOW, so that their latency hides behind the
          // item wait; the table sits behind the stash
          const int etid = static_cast<int>(threadIdx.x) - 64;
          const int ecol = n0 + etid;
          const bool coef_thread = use_stash && etid < BN && ecol < p.N;
          const uint32_t coef_base = use_stash ? smem_u32(smem) + static_cast<uint32_t>(EW * STASH_K) * 2048u : 0u;
          static_assert(EW * STASH_K * 2048 + 2 * 256 * 4 <= SL::STAGES * SL::STAGE_BYTES, "coefficient table does not fit the ring");
          float cg = 0.f, cbt = 0.f, csc = 0.f, csh = 0.f;
          if (coef_thread) {
            cg = __ldg(p.gn.gamma + ecol);
            cbt = __ldg(p.gn.beta + ecol);
            if (p.gn.scale_shift) {
              const float* ssp = p.gn.scale_shift + static_cast<size_t>(b) * p.gn.ss_ld;
              csc = __ldg(ssp + ecol);
              csh = __ldg(ssp + p.N + ecol);
            }
          }
          if (warp == 2 && lane == 0) trace_ev(p, 2, tile, 3);    // pass 1 done
          epi_bar_sync(EW * 32);                          // every warp's raw store and partials are issued
          if (warp == 2 && lane == 0)
            gn_item_arrive_wait(p.gn.counter + b, static_cast<unsigned>(p.tiles_per_batch * n_tiles_n), p.gn.counter + p.B);
          epi_bar_sync(EW * 32);
          if (warp == 2 && lane == 0) trace_ev(p, 2, tile, 4);    // item complete
          if (warp == 2 + EW - 1 && lane == 0)              // (a warp that folds no group unless the tile spans EW groups)
            gn_item_depart(p.gn.counter + b, static_cast<unsigned>(p.tiles_per_batch * n_tiles_n));
          const int cpg = p.N / p.gn.G;
          const int gfirst = n0 / cpg;
          const int glast = (min(n0 + BN, p.N) - 1) / cpg;
          const int chunks = p.tiles_per_batch * 4;
          const float* sbase = p.stats + static_cast<size_t>(b) * chunks * p.stats_nblk * 2;
          for (int g = gfirst + (warp - 2); g <= glast; g += EW) {
            double sum, sumsq;
            gn_group_fold(sbase, chunks, p.stats_nblk, (g * cpg) >> 4, cpg >> 4, lane, sum, sumsq);
            if (lane == 0) {
              const double n = static_cast<double>(p.L) * cpg;
              const double mean = sum / n;
              double var = sumsq / n - mean * mean;
              if (var < 0.0) var = 0.0;
              s_mean[g - gfirst] = static_cast<float>(mean);
              s_rstd[g - gfirst] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.gn.eps)));
            }
          }
          epi_bar_sync(EW * 32);
          if (warp == 2 && lane == 0) trace_ev(p, 2, tile, 5);    // statistics folded
          if (use_stash) {
            if (coef_thread) {                                    // the same arithmetic, in the same order, as the per-lane path
              const int grp = ecol / cpg - gfirst;
              const float mean = s_mean[grp], rstd = s_rstd[grp];
              float a = rstd * cg;
              float bb = cbt - mean * a;
              if (p.gn.scale_shift) {
                const float s1 = 1.0f + csc;
                a *= s1;
                bb = bb * s1 + csh;
              }
              sts_f32(coef_base + etid * 4, a);
              sts_f32(coef_base + (256 + etid) * 4, bb);
            }
            epi_bar_sync(EW * 32);
          }
#pragma unroll 1
          for (int ch = part; ch < NCH; ch += PARTS) {
            const int gcol = n0 + ch * 32;
            if (gcol >= p.N) continue;
            gn_apply_chunk<SPLIT>(p, e, gcol, lq, b, obase, s_mean, s_rstd, gfirst, cpg,
                                  use_stash ? stash_base + static_cast<uint32_t>(ch / PARTS) * 2048u : 0u, coef_base, n0);
          }
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
  if (threadIdx.x == 0) { trace_ev(p, 0, 0, 11); trace_ev(p, 3, 0, 2); }
}

// ---------------------------------------------------------------- CUDA-core reference of the same contract
__global__ void gemm_ref_kernel(GemmParams p, GemmRefSrc s0, GemmRefSrc s1, GemmRefSrc s2, const __nv_bfloat16* W,
                                int k_tot) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;   // packed output column
  const int l = blockIdx.y;
  const int b = blockIdx.z;
  if (n >= p.n_pad) return;
  GemmRefSrc srcs[3] = {s0, s1, s2};
  float acc = 0.f;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg sg = p.seg[s];
    const GemmRefSrc& src = srcs[sg.src];
    for (int t = 0; t < sg.ntap; ++t) {
      const int ls = l + sg.row_off + t;
      for (int c = 0; c < sg.npanels * BK; ++c) {
        const int cs = sg.c0 + c;
        const int k = (sg.wk0 + t * sg.tap_stride) * BK + c;
        float a = 0.f;
        if (ls >= 0 && ls < src.L && cs < src.C)
          a = h16_to_f32(reinterpret_cast<const __nv_bfloat16*>(src.ptr)[(static_cast<size_t>(b) * src.L + ls) * src.ld + cs]);
        acc += a * h16_to_f32(W[static_cast<size_t>(n) * k_tot + k]);
      }
    }
  }
  if (p.bias) acc += p.bias[n];
  // out is a dense fp32 [B, L, n_pad] scratch in the reference (raw accumulator + bias)
  reinterpret_cast<float*>(p.out)[(static_cast<size_t>(b) * p.L + l) * p.n_pad + n] = acc;
}

inline int min_i(int a, int b) { return a < b ? a : b; }
int g_num_sms = 0;
int g_force_ew = 0;   // DV_GEMM_EW: 8 / 16 force the CTA shape of the 128-wide tile kernel (0 = heuristic)

template <int BN, int EW, bool TRIM = false, bool SPLIT = false>
cudaError_t launch_bn(const GemmParams& p, cudaStream_t st) {
  using SL = SmemLayout<BN, EW>;
  const int tiles = p.tiles_per_batch * p.B * (p.n_pad / BN);
  const int slots = g_num_sms * (EW == 8 ? 2 : 1);
  const int grid = tiles < slots ? tiles : slots;
  return launch_k(gemm_tc_kernel<BN, EW, TRIM, SPLIT>, dim3(grid), dim3(64 + EW * 32), SL::dyn_bytes(p.n_pad), st, p);
}

}  // namespace

int gemm_pick_bn(int m_tiles, int n_pad) {
  if (g_num_sms == 0) return 128;
  // small problems (a handful of row tiles: batch 1-4): 64-wide tiles put twice as many SMs to work and halve the W bytes
  // each CTA streams; the main loop of a lone CTA is bound by what one SM can pull from L2 (~0.3 us per 32 KB panel)
  if (m_tiles * (n_pad / 128) <= g_num_sms / 4 && !getenv("DV_GEMM_NO64")) return 64;
  // N = 384 (level 2): 2 x 64 tiles of 128 x 128 rows would be 192 tiles = 1.3 waves; 192-wide tiles make it one wave
  if (n_pad == 384 && m_tiles * 2 <= g_num_sms && !getenv("DV_GEMM_NO192")) return 192;
  if (n_pad % 256 != 0) return 128;
  // wide tiles halve the A traffic and the shared-memory bandwidth per FLOP, but only pay when they still fill the GPU
  return (m_tiles * (n_pad / 256) >= (g_num_sms * 4) / 5) ? 256 : 128;
}

cudaError_t gemm_tc_init() {
  {
    cudaError_t e2 = gemm_tc2_init();
    if (e2 != cudaSuccess) return e2;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<128, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SmemLayout<128, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<128, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<128, 8>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  {
    const char* ev = getenv("DV_GEMM_EW");
    if (ev) g_force_ew = atoi(ev);
  }
  e = cudaFuncSetAttribute(gemm_tc_kernel<192, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<192, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<64, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<64, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<128, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<128, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<256, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<256, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<128, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<128, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<128, 16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<128, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<256, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<256, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<192, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SmemLayout<192, 16>::dyn_bytes(MAX_NPAD));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_tc_kernel<256, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SmemLayout<256, 16>::dyn_bytes(MAX_NPAD));
}

cudaError_t launch_gemm_tc(const GemmParams& p_in, cudaStream_t st) {
  GemmParams p = p_in;
  if (p.out_bs == 0) p.out_bs = static_cast<long long>(p.L) * p.ld_out;
  if (p.trim_col % 32 != 0) return cudaErrorInvalidValue;
  if (g_num_sms == 0) {
    cudaError_t e = gemm_tc_init();
    if (e != cudaSuccess) return e;
  }
  if (p.n_pad % 128 != 0 || p.n_pad > MAX_NPAD || p.nseg < 1 || p.nseg > GEMM_MAX_SEG) return cudaErrorInvalidValue;
  if (p.out_mode == GEMM_OUT_GEGLU && p.n_pad % 128 != 0) return cudaErrorInvalidValue;
  // shared conv taps / a GroupNorm fused into the operand path are features of the two-ring kernel (csrc/gemm_tc2.cu)
  for (int i = 0; i < p.nseg; ++i)
    if (p.seg[i].ntap != 1 || p.seg[i].xf != 0) return p.gn.out ? cudaErrorInvalidValue : DV_NS::launch_gemm_tc2(p_in, st);
  for (int i = 0; i < GEMM_MAX_SRC; ++i)
    if (p.map_rows[i] != 128) return cudaErrorInvalidValue;
  const bool dense = p.out_bs == static_cast<long long>(p.L) * p.ld_out;
  const int bn = p.bn ? p.bn : gemm_pick_bn(p.tiles_per_batch * p.B, p.n_pad);
  if (p.gn.out != nullptr) {
    // fused output GroupNorm: plain 16-bit epilogue with statistics, and every tile of a batch item must be co-scheduled in one
    // round of the persistent tile loop (the item barrier waits on them): checked for the narrowest tile any branch below picks
    if (p.out_mode != GEMM_OUT_BF16 || !p.stats || !p.gn.counter || p.relu || p.row_len || p.trim_col > 0 || !dense ||
        p.gn.G < 1 || p.N % (16 * p.gn.G) != 0)
      return cudaErrorInvalidValue;
    if (p.tiles_per_batch * (p.n_pad / 64) > g_num_sms) return cudaErrorInvalidValue;
    if ((min_i(256, p.N) + p.N / p.gn.G - 1) / (p.N / p.gn.G) + 1 > 32) return cudaErrorInvalidValue;
  }
  if (p.out_lo != nullptr) {                             // split precision (hi + lo output planes): the same tile shapes
    if (p.res != nullptr && p.res_lo == nullptr) return cudaErrorInvalidValue;
    if (p.trim_col > 0 || !dense) {
      if (p.out_mode != GEMM_OUT_BF16) return cudaErrorInvalidValue;
      return launch_bn<128, 16, true, true>(p, st);
    }
    static const bool one_shape = getenv("DV_GEMM_SPLIT128") != nullptr;     // A/B: the single 128 x 128 shape of the first version
    if (!one_shape) {
      if (bn == 256 && p.n_pad % 256 == 0) return launch_bn<256, 16, false, true>(p, st);
      if (bn == 192 && p.n_pad % 192 == 0 && p.out_mode != GEMM_OUT_GEGLU) return launch_bn<192, 16, false, true>(p, st);
    }
    return launch_bn<128, 16, false, true>(p, st);
  }
  if (p.relu || p.row_len) {                             // prompt-encoder epilogue options live in the rare-path instantiation
    if (p.out_mode != GEMM_OUT_BF16) return cudaErrorInvalidValue;
    if (p.trim_col == 0) p.trim_col = 1 << 30;           // (no trimmed half row)
    if (bn == 256 && p.n_pad % 256 == 0) return launch_bn<256, 16, true>(p, st);
    return launch_bn<128, 16, true>(p, st);
  }
  if (p.trim_col > 0 || !dense) {                        // odd-length folded upsample conv: its own instantiations
    if (p.out_mode != GEMM_OUT_BF16) return cudaErrorInvalidValue;
    if (bn == 256 && p.n_pad % 256 == 0) return launch_bn<256, 16, true>(p, st);
    return launch_bn<128, 16, true>(p, st);
  }
  if (bn == 256) {
    if (p.n_pad % 256 != 0) return cudaErrorInvalidValue;
    return launch_bn<256, 16>(p, st);
  }
  if (bn == 192) {
    if (p.n_pad % 192 != 0 || p.out_mode == GEMM_OUT_GEGLU) return cudaErrorInvalidValue;
    return launch_bn<192, 16>(p, st);
  }
  if (bn == 64 && p.out_mode != GEMM_OUT_GEGLU) return launch_bn<64, 16>(p, st);   // GEGLU pairs u / gate inside 128 columns
  // two small CTAs per SM pay once there is more than one wave of 128-wide tiles (measured on the C2 shapes: 5-10 %
  // faster at levels 0-2, slower for the few long-K tiles of level 3)
  const int tiles = p.tiles_per_batch * p.B * (p.n_pad / 128);
  const int ew = g_force_ew ? g_force_ew : (tiles >= g_num_sms + g_num_sms / 8 ? 8 : 16);
  return ew == 8 ? launch_bn<128, 8>(p, st) : launch_bn<128, 16>(p, st);
}

cudaError_t launch_gemm_ref(const GemmParams& p, const GemmRefSrc* srcs, const void* W, int k_tot, cudaStream_t st) {
  for (int i = 0; i < p.nseg; ++i)
    if (p.seg[i].xf || p.seg[i].wsel || p.seg[i].src >= 3) return cudaErrorInvalidValue;
  dim3 grid((p.n_pad + 127) / 128, p.L, p.B);
  gemm_ref_kernel<<<grid, 128, 0, st>>>(p, srcs[0], srcs[1], srcs[2], reinterpret_cast<const __nv_bfloat16*>(W), k_tot);
  return cudaGetLastError();
}

}  // namespace DV_NS

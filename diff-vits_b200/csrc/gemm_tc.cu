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

#include "kernels.h"
#include "ptx.cuh"

namespace dv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int EPI_PITCH = 80;                       // bytes per staged row: 32 bf16 (64 B) + 16 B pad
constexpr int EPI_SLAB = 32 * EPI_PITCH;            // per epilogue warp

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 128 ? 6 : 4;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = EPI_OFF + EPI_WARPS * EPI_SLAB;
  static constexpr int TOTAL = BAR_OFF + 256;          // barriers + tmem slot
  static constexpr int DYN_BYTES = TOTAL + 1024;       // slack for 1024-byte alignment
  static_assert(DYN_BYTES <= 232448, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using SL = SmemLayout<BN>;
  constexpr int STAGES = SL::STAGES;
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
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mi = tile / n_tiles_n;
        const int n0 = (tile % n_tiles_n) * BN;
        const int b = mi / p.tiles_per_batch;
        const int l0 = (mi % p.tiles_per_batch) * BM;
        int kp = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          const CUtensorMap* ma = &p.mapA[sg.src];
          for (int q = 0; q < sg.npanels; ++q, ++kp) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * SL::STAGE_BYTES;
            uint8_t* sb = sa + SL::A_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], SL::STAGE_BYTES);
            tma_load_3d(sa, ma, &full_bar[stage], sg.c0 + q * BK, l0 + sg.row_off, b);
            // the W map's box is 64 x 128 rows: a 256-wide tile takes two boxes
#pragma unroll
            for (int h = 0; h < BN / 128; ++h)
              tma_load_2d(sb + h * (128 * BK * 2), &p.mapW, &full_bar[stage], kp * BK, n0 + h * 128);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kp = 0; kp < total_panels; ++kp) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + SL::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kp | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: warps 2..9
    // TMEM lane quarter is fixed by hardware to (warp id % 4); the two warps of a quarter split the columns.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* slab = smem + SL::EPI_OFF + (warp - 2) * EPI_SLAB;
    const int srow = lane >> 2, spiece = lane & 3;       // coalesced phase: 8 rows x 4 pieces of 16 B per instruction
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
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      const size_t rowbase = static_cast<size_t>(b) * p.L;

      if (p.out_mode == GEMM_OUT_GEGLU) {
        // tile columns repeat [u 64 | gate 64]; this warp takes 32-wide (u, gate) chunk pairs with index % 2 == half
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        constexpr int NPAIR = BN / 64;
        for (int pr = half; pr < NPAIR; pr += 2) {
          const int cu = (pr >> 1) * 128 + (pr & 1) * 32;
          const int ocol = (n0 >> 1) + (pr >> 1) * 64 + (pr & 1) * 32;
          uint32_t u[32], g[32];
          tmem_ld32(taddr + cu, u);
          tmem_ld32(taddr + cu + 64, g);
          tmem_ld_wait();
          if (pr + 2 >= NPAIR) {                          // last TMEM read of this tile by this warp
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float u0 = __uint_as_float(u[j]), u1 = __uint_as_float(u[j + 1]);
            float g0 = __uint_as_float(g[j]), g1 = __uint_as_float(g[j + 1]);
            if (p.bias) {
              u0 += __ldg(p.bias + n0 + cu + j);
              u1 += __ldg(p.bias + n0 + cu + j + 1);
              g0 += __ldg(p.bias + n0 + cu + 64 + j);
              g1 += __ldg(p.bias + n0 + cu + 64 + j + 1);
            }
            packed[j >> 1] = pack_bf16(u0 * gelu_erf_f(g0), u1 * gelu_erf_f(g1));
          }
          uint4* mine = reinterpret_cast<uint4*>(slab + lane * EPI_PITCH);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mine[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + srow;
            if (lq + r < p.L)
              *reinterpret_cast<uint4*>(out + (rowbase + lq + r) * p.ld_out + ocol + spiece * 8) =
                  *reinterpret_cast<const uint4*>(slab + r * EPI_PITCH + spiece * 16);
          }
          __syncwarp();
        }
      } else if (p.out_mode == GEMM_OUT_F32_NCL) {
        float* out = reinterpret_cast<float*>(p.out) + static_cast<size_t>(b) * p.N * p.L + (valid ? l : 0);
        constexpr int NCH = BN / 32;
        for (int ch = half * (NCH / 2); ch < (half + 1) * (NCH / 2); ++ch) {
          const int c = ch * 32;
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          if (ch + 1 == (half + 1) * (NCH / 2)) {
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c + j;
              if (n < p.N) {
                float v = __uint_as_float(r[j]);
                if (p.bias) v += __ldg(p.bias + n);
                out[static_cast<size_t>(n) * p.L] = v;  // lanes = consecutive l: coalesced
              }
            }
          }
        }
      } else {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(p.res);
        constexpr int NCH = BN / 32;
        for (int ch = half * (NCH / 2); ch < (half + 1) * (NCH / 2); ++ch) {
          const int c = ch * 32;
          const int gcol = n0 + c;
          const bool live = gcol < p.N;                   // padded output columns are neither read nor written
          uint4 rv[4];
          if (res && live) {                              // coalesced residual fetch, in flight during the TMEM load
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + srow;
              rv[i] = (lq + r < p.L)
                          ? *reinterpret_cast<const uint4*>(res + (rowbase + lq + r) * p.ld_res + gcol + spiece * 8)
                          : make_uint4(0, 0, 0, 0);
            }
          }
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          if (res && live) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(slab + (i * 8 + srow) * EPI_PITCH + spiece * 16) = rv[i];
          }
          tmem_ld_wait();
          if (ch + 1 == (half + 1) * (NCH / 2)) {
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          if (!live) continue;
          __syncwarp();
          uint4* mine = reinterpret_cast<uint4*>(slab + lane * EPI_PITCH);
          uint32_t rr[16];
          if (res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = mine[j];
              rr[4 * j] = t.x; rr[4 * j + 1] = t.y; rr[4 * j + 2] = t.z; rr[4 * j + 3] = t.w;
            }
          }
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
            if (p.bias) {
              v0 += __ldg(p.bias + gcol + j);
              v1 += __ldg(p.bias + gcol + j + 1);
            }
            if (res) {
              v0 += bf16_lo(rr[j >> 1]);
              v1 += bf16_hi(rr[j >> 1]);
            }
            packed[j >> 1] = pack_bf16(v0, v1);
          }
          if (p.stats) {
            // statistics of the STORED (bf16-rounded) values, per 16-channel block, over this warp's 32 rows
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            if (valid) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a = bf16_lo(packed[j]), c2 = bf16_hi(packed[j]);
                s0 += a + c2; q0 += a * a + c2 * c2;
                const float d = bf16_lo(packed[8 + j]), e = bf16_hi(packed[8 + j]);
                s1 += d + e; q1 += d * d + e * e;
              }
            }
            s0 = warp_sum(s0); q0 = warp_sum(q0); s1 = warp_sum(s1); q1 = warp_sum(q1);
            if (lane == 0) {
              const size_t chunk = (static_cast<size_t>(b) * p.tiles_per_batch + mt) * 4 + quad;
              *reinterpret_cast<float4*>(p.stats + (chunk * p.stats_nblk + (gcol >> 4)) * 2) = make_float4(s0, q0, s1, q1);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mine[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rrow = i * 8 + srow;
            if (lq + rrow < p.L)
              *reinterpret_cast<uint4*>(out + (rowbase + lq + rrow) * p.ld_out + gcol + spiece * 8) =
                  *reinterpret_cast<const uint4*>(slab + rrow * EPI_PITCH + spiece * 16);
          }
          __syncwarp();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
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
  int k = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg sg = p.seg[s];
    const GemmRefSrc& src = srcs[sg.src];
    const int ls = l + sg.row_off;
    for (int c = 0; c < sg.npanels * BK; ++c, ++k) {
      const int cs = sg.c0 + c;
      float a = 0.f;
      if (ls >= 0 && ls < src.L && cs < src.C)
        a = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src.ptr)[(static_cast<size_t>(b) * src.L + ls) * src.ld + cs]);
      acc += a * __bfloat162float(W[static_cast<size_t>(n) * k_tot + k]);
    }
  }
  if (p.bias) acc += p.bias[n];
  // out is a dense fp32 [B, L, n_pad] scratch in the reference (raw accumulator + bias)
  reinterpret_cast<float*>(p.out)[(static_cast<size_t>(b) * p.L + l) * p.n_pad + n] = acc;
}

int g_num_sms = 0;

template <int BN>
cudaError_t launch_bn(const GemmParams& p, cudaStream_t st) {
  using SL = SmemLayout<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::DYN_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.tiles_per_batch * p.B * (p.n_pad / BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, SL::DYN_BYTES, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

int gemm_pick_bn(int m_tiles, int n_pad) {
  if (n_pad % 256 != 0) return 128;
  if (g_num_sms == 0) return 128;
  // wide tiles halve the A traffic and the shared-memory bandwidth per FLOP, but only pay when they still fill the GPU
  return (m_tiles * (n_pad / 256) >= (g_num_sms * 4) / 5) ? 256 : 128;
}

cudaError_t gemm_tc_init() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SmemLayout<128>::DYN_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SmemLayout<256>::DYN_BYTES);
}

cudaError_t launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  if (g_num_sms == 0) {
    cudaError_t e = gemm_tc_init();
    if (e != cudaSuccess) return e;
  }
  if (p.n_pad % 128 != 0 || p.nseg < 1 || p.nseg > GEMM_MAX_SEG) return cudaErrorInvalidValue;
  if (p.out_mode == GEMM_OUT_GEGLU && p.n_pad % 128 != 0) return cudaErrorInvalidValue;
  const int bn = p.bn ? p.bn : gemm_pick_bn(p.tiles_per_batch * p.B, p.n_pad);
  if (bn == 256) {
    if (p.n_pad % 256 != 0) return cudaErrorInvalidValue;
    return launch_bn<256>(p, st);
  }
  return launch_bn<128>(p, st);
}

cudaError_t launch_gemm_ref(const GemmParams& p, const GemmRefSrc* srcs, const void* W, int k_tot, cudaStream_t st) {
  dim3 grid((p.n_pad + 127) / 128, p.L, p.B);
  gemm_ref_kernel<<<grid, 128, 0, st>>>(p, srcs[0], srcs[1], srcs[2], reinterpret_cast<const __nv_bfloat16*>(W), k_tot);
  return cudaGetLastError();
}

}  // namespace dv

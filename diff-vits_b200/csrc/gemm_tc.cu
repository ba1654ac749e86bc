// Tap-GEMM on the 5th-gen tensor cores: implicit-GEMM conv1d (k=3, stride 1 or 2, nearest-2x upsample folded
// into the weight packing), 1x1 conv and Linear, all as   D[128 x BN] += A_panel[128 x 64] . W_panel[BN x 64]^T.
//
//   warp 0 : TMA producer  - one elected lane issues cp.async.bulk.tensor loads (A: 3-D [C, L, B] box 64x128x1,
//                            W: 2-D [K, N] box 64xBN), 128-byte swizzle, completion on the stage's "full" mbarrier.
//                            Conv taps are the same box shifted by -1/0/+1 rows; rows outside [0, L) are zero
//                            filled by TMA, which is exactly the conv padding. torch.cat inputs are extra A maps.
//   warp 1 : MMA issuer    - lane 0 issues tcgen05.mma.cta_group::1.kind::f16 (BF16 x BF16 -> FP32 in TMEM),
//                            4 x (K=16) per 64-wide panel; tcgen05.commit releases the smem stage / signals the
//                            epilogue. Also owns the TMEM allocation (BN columns).
//   warps 2-5 : epilogue   - tcgen05.ld 32 lanes x 32 columns -> registers -> (+bias) (+residual) (GEGLU)
//                            -> bf16 channel-last store, or fp32 [B, N, L] transposed store (conv_out).
//
// One 128 x BN output tile per CTA, 2 CTAs per SM (3 stages x 32 KB each, 128 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.
#include <cstdio>

#include "kernels.h"
#include "ptx.cuh"

namespace dv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128;          // barriers + tmem slot
  static constexpr int DYN_BYTES = TOTAL + 1024;       // slack for 1024-byte alignment
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using SL = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int b = blockIdx.x / p.tiles_per_batch;
  const int l0 = (blockIdx.x % p.tiles_per_batch) * BM;
  const int n0 = blockIdx.y * BN;

  int total_panels = 0;
  for (int s = 0; s < p.nseg; ++s) total_panels += p.seg[s].npanels;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GEMM_MAX_SRC; ++s) tma_prefetch_desc(&p.mapA[s]);
    tma_prefetch_desc(&p.mapW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
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
          tma_load_2d(sb, &p.mapW, &full_bar[stage], kp * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kp = 0; kp < total_panels; ++kp) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + SL::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kp | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int l = l0 + row;
    const bool valid = l < p.L;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const size_t row_idx = static_cast<size_t>(b) * p.L + (valid ? l : 0);

    if (p.out_mode == GEMM_OUT_GEGLU) {
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + row_idx * p.ld_out + (n0 / 2);
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        uint32_t u[32], g[32];
        tmem_ld32(taddr + c, u);
        tmem_ld32(taddr + BN / 2 + c, g);
        tmem_ld_wait();
        if (valid) {
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float u0 = __uint_as_float(u[j]), u1 = __uint_as_float(u[j + 1]);
            float g0 = __uint_as_float(g[j]), g1 = __uint_as_float(g[j + 1]);
            if (p.bias) {
              u0 += __ldg(p.bias + n0 + c + j);
              u1 += __ldg(p.bias + n0 + c + j + 1);
              g0 += __ldg(p.bias + n0 + BN / 2 + c + j);
              g1 += __ldg(p.bias + n0 + BN / 2 + c + j + 1);
            }
            packed[j >> 1] = pack_bf16(u0 * gelu_erf_f(g0), u1 * gelu_erf_f(g1));
          }
          uint4* dst = reinterpret_cast<uint4*>(out + c);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
    } else if (p.out_mode == GEMM_OUT_F32_NCL) {
      float* out = reinterpret_cast<float*>(p.out) + static_cast<size_t>(b) * p.N * p.L + (valid ? l : 0);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c, r);
        tmem_ld_wait();
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
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + row_idx * p.ld_out + n0;
      const __nv_bfloat16* res =
          p.res ? reinterpret_cast<const __nv_bfloat16*>(p.res) + row_idx * p.ld_res + n0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c, r);
        tmem_ld_wait();
        if (valid && n0 + c < p.N) {
          uint32_t rr[16];
          if (res) {
            const uint4* rs = reinterpret_cast<const uint4*>(res + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 t = rs[j];
              rr[4 * j] = t.x; rr[4 * j + 1] = t.y; rr[4 * j + 2] = t.z; rr[4 * j + 3] = t.w;
            }
          }
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
            if (p.bias) {
              v0 += __ldg(p.bias + n0 + c + j);
              v1 += __ldg(p.bias + n0 + c + j + 1);
            }
            if (res) {
              v0 += bf16_lo(rr[j >> 1]);
              v1 += bf16_hi(rr[j >> 1]);
            }
            packed[j >> 1] = pack_bf16(v0, v1);
          }
          uint4* dst = reinterpret_cast<uint4*>(out + c);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
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

}  // namespace

cudaError_t launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  constexpr int BN = 128;
  using SL = SmemLayout<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::DYN_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (p.n_pad % BN != 0 || p.nseg < 1 || p.nseg > GEMM_MAX_SEG) return cudaErrorInvalidValue;
  dim3 grid(p.tiles_per_batch * p.B, p.n_pad / BN);
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, SL::DYN_BYTES, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_gemm_ref(const GemmParams& p, const GemmRefSrc* srcs, const void* W, int k_tot, cudaStream_t st) {
  dim3 grid((p.n_pad + 127) / 128, p.L, p.B);
  GemmRefSrc z{nullptr, 0, 0, 0};
  gemm_ref_kernel<<<grid, 128, 0, st>>>(p, srcs[0], srcs[1], srcs[2], reinterpret_cast<const __nv_bfloat16*>(W), k_tot);
  (void)z;
  return cudaGetLastError();
}

}  // namespace dv

// Probe: does tcgen05.mma accept a K-major SWIZZLE_128B A operand whose start address is shifted by whole rows
// (+128 B per row, i.e. NOT a multiple of the 1024-byte swizzle atom)? This is what lets the three taps of a k=3 conv read
// ONE staged [130 x 64] activation box with three row-shifted descriptors instead of three TMA loads.
// Variants: descriptor base_offset field = 0, or (start_address >> 7) & 7 as the PTX ISA prescribes for unaligned starts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I diff-vits_b200/csrc -o tools/ubench/umma_shift tools/ubench/umma_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace dv;

constexpr int ROWS = 144;   // staged A rows (>= 128 + max shift)
constexpr int N = 64;

__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t kbyte) {   // byte offset of (row, byte kbyte of the 128-byte row)
  return row * 128 + ((((kbyte >> 4) ^ (row & 7)) << 4) | (kbyte & 15));
}

__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int use_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // ROWS x 128 B
  uint8_t* sb = smem + ROWS * 128;          // N x 128 B (ROWS * 128 = 18432 = 18 atoms: still 1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sa + sw128(r, k * 2)) = A[i];
  }
  for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sb + sw128(r, k * 2)) = B[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    uint64_t adesc = umma_desc_sw128(a_addr);
    if (use_base_offset) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sb));
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (threadIdx.x < 128) {
    const int warp = threadIdx.x >> 5;
    uint32_t r[32];
    for (int c = 0; c < N; c += 32) {
      tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[threadIdx.x * N + c + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  std::vector<__nv_bfloat16> hA(ROWS * 64), hB(N * 64);
  std::vector<float> fA(ROWS * 64), fB(N * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = v; }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = v; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = ROWS * 128 + N * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 1, 2, 3, 7, 8, 9, 15}) {
      cudaMemset(dD, 0, 128 * N * 4);
      probe<<<1, 128, smem>>>(dA, dB, dD, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> hD(128 * N);
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += fA[(m + shift) * 64 + k] * fB[n * 64 + k];
          maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
        }
      printf("base_offset_field=%d shift=%2d rows: %s  max|err| = %g  (%s)\n", bo, shift, maxerr < 1e-3 ? "OK  " : "FAIL", maxerr,
             cudaGetErrorString(e));
    }
  return 0;
}

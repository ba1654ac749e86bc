// Store-path probe for the GEMM epilogue: how fast can one SM push a 128-row x 256-column bf16 tile (64 KB) into a row-major
// [M, ld] tensor, as a function of the row-segment length one warp-wide STG.128 covers? 148 CTAs x 512 threads (16 warps), every
// CTA writes `tiles` different tiles (a strided sweep over a 64 MB buffer, so lines are not rewritten while still in flight).
//   seg = 64  : 8 rows x 64 B per instruction  (what csrc/gemm_tc.cu does: 32-column chunks per warp)
//   seg = 128 : 4 rows x 128 B, seg = 256 : 2 rows x 256 B, seg = 512 : 1 row x 512 B
// Build: nvcc -arch=sm_100a -O3 -o stg stg.cu ; prints clk per tile and B/clk/SM for each pattern, with all SMs active and alone.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int SEG>
__global__ void __launch_bounds__(512) k(uint4* out, int ld_bytes, int tiles, long long* clk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int LPR = SEG / 16;            // lanes per row segment
  constexpr int RPI = 32 / LPR;            // rows per instruction
  // the tile: 128 rows x 512 bytes. warp w covers a [rows x SEG bytes] block: blocks are laid out so that the 16 warps x 4
  // instructions cover 64 KB exactly
  constexpr int BLK_ROWS = 4 * RPI;        // rows per warp block (4 instructions)
  constexpr int BLKS_PER_ROWBAND = 512 / SEG;
  const int blk = warp;                    // 16 blocks per pass; passes = (128 / BLK_ROWS) * BLKS_PER_ROWBAND / 16
  constexpr int PASSES = (128 / BLK_ROWS) * BLKS_PER_ROWBAND / 16;
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    char* base = reinterpret_cast<char*>(out) + (static_cast<size_t>(blockIdx.x) * tiles + t) * 128 * static_cast<size_t>(ld_bytes) % (size_t(1) << 30);
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int b = ps * 16 + blk;
      const int band = b / BLKS_PER_ROWBAND, col = (b % BLKS_PER_ROWBAND) * SEG;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = band * BLK_ROWS + i * RPI + lane / LPR;
        *reinterpret_cast<uint4*>(base + static_cast<size_t>(row) * ld_bytes + col + (lane % LPR) * 16) = v;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int SEG>
void run(uint4* buf, long long* clk, int grid, int tiles) {
  k<SEG><<<grid, 512>>>(buf, 1536, tiles, clk);
  cudaDeviceSynchronize();
  k<SEG><<<grid, 512>>>(buf, 1536, tiles, clk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mx = 0, av = 0;
  for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; av += h[i]; }
  av /= grid;
  printf("seg %3d B  grid %3d  tiles %2d : %8.0f clk per tile (issue side, mean; max %.0f)  %6.1f B/clk/SM\n", SEG, grid, tiles, av / tiles,
         mx / tiles, 65536.0 * tiles / av);
}

int main() {
  uint4* buf;
  long long* clk;
  cudaMalloc(&buf, size_t(1) << 30);
  cudaMalloc(&clk, 148 * sizeof(long long));
  for (int grid : {148, 1}) {
    for (int tiles : {1, 4}) {
      run<64>(buf, clk, grid, tiles);
      run<128>(buf, clk, grid, tiles);
      run<256>(buf, clk, grid, tiles);
      run<512>(buf, clk, grid, tiles);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

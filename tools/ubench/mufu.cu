// Throughput probe: MUFU.EX2 f32 vs f16x2 vs bf16x2, 3-input max, FFMA2. Build: nvcc -arch=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = __float_as_uint(a0), h1 = __float_as_uint(a1), h2 = __float_as_uint(a2), h3 = __float_as_uint(a3);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 3) {   // 3-input max
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(a2));
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(a2), "f"(a3));
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(a0));
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(a0), "f"(a1));
    } else if (MODE == 4) {   // 2-input max
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a0) : "f"(a1));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a1) : "f"(a2));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a2) : "f"(a3));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a3) : "f"(a0));
    } else if (MODE == 5) {   // cvt.rn.f16x2.f32
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(a0), "f"(a1));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(a1), "f"(a2));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(a2), "f"(a3));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h3) : "f"(a3), "f"(a0));
      a0 += __uint_as_float(h0); a1 += __uint_as_float(h1);
    } else if (MODE == 6) {   // fma.rn.f32x2
      asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}"
                   : "+f"(a0), "+f"(a1) : "f"(a2), "f"(a3));
      asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}"
                   : "+f"(a2), "+f"(a3) : "f"(a0), "f"(a1));
      asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}"
                   : "+f"(a0), "+f"(a1) : "f"(a2), "f"(a3));
      asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}"
                   : "+f"(a2), "+f"(a3) : "f"(a0), "f"(a1));
    } else if (MODE == 7) {   // plain ffma
      a0 = fmaf(a0, a1, a2); a1 = fmaf(a1, a2, a3); a2 = fmaf(a2, a3, a0); a3 = fmaf(a3, a0, a1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}

template <int MODE>
void run(const char* name, float* out) {
  const int iters = 4096, blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = double(blocks) * threads * iters * 4;   // thread-level instructions
  printf("%-28s %8.3f ms  %8.2f Ginstr/s (thread-level)  = %.2f per clk per SM @1.9GHz\n", name, ms, instr / ms / 1e6,
         instr / ms / 1e6 / 148 / 1.9);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("ex2.approx.ftz.f32", out);
  run<1>("ex2.approx.f16x2", out);
  run<2>("ex2.approx.ftz.bf16x2", out);
  run<3>("max.f32 (3 inputs)", out);
  run<4>("max.f32 (2 inputs)", out);
  run<5>("cvt.rn.f16x2.f32 (+2 fadd/4)", out);
  run<6>("fma.rn.f32x2", out);
  run<7>("fma.rn.f32", out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

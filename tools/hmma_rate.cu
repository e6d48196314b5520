// Micro-benchmark: issue rate of the legacy tensor path (mma.sync.m16n8k16 f16 -> f32, SASS HMMA.16816.F32) on
// sm_100a, per SM, as a function of warps per SM and independent accumulator chains per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hmma_rate tools/hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CH>
__global__ void k(float* out, int iters, uint32_t seed) {
  float acc[CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  uint32_t a = seed + threadIdx.x, b = seed * 3 + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) mma(acc[c], a, a + c, b, b + c, a ^ b, b + 1);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  if (s == 12345.f) out[0] = s;
}

template <int CH>
void run(int warps, int iters) {
  float* d;
  cudaMalloc(&d, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<CH><<<148, warps * 32>>>(d, 10, 1);
  cudaEventRecord(e0);
  k<CH><<<148, warps * 32>>>(d, iters, 1);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double n = (double)iters * CH * warps;               // HMMAs per SM
  double cyc = ms * 1e-3 * clk_khz * 1e3;
  printf("warps/SM %2d chains %d: %.2f HMMA/clk/SM (%.1f clk per HMMA per SMSP at 4 SMSP)  -> %.0f dense TFLOP/s, %.2f TB/s of int4 weights at M=1\n",
         warps, CH, n / cyc, 4.0 * cyc / n, n / cyc * 148 * clk_khz * 1e3 * 4096 * 2 / 2 / 1e12 * 1.0,
         n / cyc * 148 * clk_khz * 1e3 * 256 * 0.51953 / 1e12);
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<1>(w, 20000);
    run<4>(w, 20000);
    run<8>(w, 20000);
  }
  return 0;
}

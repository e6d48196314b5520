// Glue kernels the reference's fused modules call through awq_ext:
//   rmsnorm       <- awq_ext.layernorm_forward_cuda   (awq/modules/fused/norm.py:33-36)
//   silu_and_mul  <- awq_ext.silu_and_mul             (awq/modules/fused/moe.py:76)
// fp16 in/out, fp32 math.  Bandwidth-trivial (KBs per decode step); kept simple.
#include "common.cuh"
#include "kernels.h"

namespace b200awq {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one CTA per row
__global__ void __launch_bounds__(256)
    rmsnorm_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ out, int hidden,
                   float eps) {
  __shared__ float wsum[8];
  pdl_trigger();
  pdl_wait();
  const __half* xr = x + (int64_t)blockIdx.x * hidden;
  __half* orow = out + (int64_t)blockIdx.x * hidden;
  float ss = 0.f;
  const bool vec = (hidden % 8) == 0 && (reinterpret_cast<uintptr_t>(xr) % 16) == 0;
  if (vec) {
    for (int i = threadIdx.x * 8; i < hidden; i += blockDim.x * 8) {
      uint4 v = *reinterpret_cast<const uint4*>(xr + i);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        ss += f.x * f.x + f.y * f.y;
      }
    }
  } else {
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
      float f = __half2float(xr[i]);
      ss += f * f;
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += (i < (blockDim.x >> 5)) ? wsum[i] : 0.f;
  const float rs = rsqrtf(tot / static_cast<float>(hidden) + eps);
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
    float f = __half2float(xr[i]) * rs * __half2float(w[i]);
    orow[i] = __float2half_rn(f);
  }
}

__global__ void __launch_bounds__(256)
    silu_mul_kernel(const __half* __restrict__ gu, __half* __restrict__ out, int rows, int d) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * d) return;
  const int r = static_cast<int>(i / d), j = static_cast<int>(i % d);
  const float g = __half2float(gu[(int64_t)r * 2 * d + j]);
  const float u = __half2float(gu[(int64_t)r * 2 * d + d + j]);
  out[i] = __float2half_rn(g / (1.f + __expf(-g)) * u);
}

cudaError_t rmsnorm(const void* x, const void* w, void* out, int rows, int hidden, float eps, cudaStream_t st) {
  return launch_kernel(rmsnorm_kernel, dim3(rows), dim3(256), 0, st, reinterpret_cast<const __half*>(x),
                       reinterpret_cast<const __half*>(w), reinterpret_cast<__half*>(out), hidden, eps);
}
cudaError_t silu_and_mul(const void* gate_up, void* out, int rows, int d, cudaStream_t st) {
  const int64_t n = (int64_t)rows * d;
  return launch_kernel(silu_mul_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st,
                       reinterpret_cast<const __half*>(gate_up), reinterpret_cast<__half*>(out), rows, d);
}

}  // namespace b200awq

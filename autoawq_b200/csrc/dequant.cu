// awq_ext.dequantize_weights_cuda replacement: W[K, N] fp16 = (q - z) * s, GEMM layout, bit-exact with
// awq/utils/packing_utils.py:87-102 (exact integer difference, one RN rounding of the product).
// Pure streaming kernel: reads K*N/2 bytes, writes 2*K*N bytes.  Each thread owns WPT words (8*WPT
// columns) and ROWS consecutive k-rows of one quantisation group, so scales / zeros are fetched once.
#include "common.cuh"
#include "kernels.h"

namespace b200awq {

template <int WPT, int ROWS>
__global__ void __launch_bounds__(256)
    dequant_gemm_kernel(const int32_t* __restrict__ qweight, const __half* __restrict__ scales,
                        const int32_t* __restrict__ qzeros, __half* __restrict__ out, int K, int N, int G) {
  pdl_trigger();
  pdl_wait();
  const int NW = N >> 3;
  const int nvec = NW / WPT;  // word-vectors per row
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = static_cast<int>(gid % nvec);
  const int64_t rb = gid / nvec;
  const int r0 = static_cast<int>(rb * ROWS);
  if (r0 >= K) return;
  const int g = r0 / G;
  const int wc = v * WPT;

  ZeroPairs zp[WPT];
  uint4 sc[WPT];
#pragma unroll
  for (int w = 0; w < WPT; ++w) {
    zp[w] = awq_zero_pairs(static_cast<uint32_t>(qzeros[(int64_t)g * NW + wc + w]));
    sc[w] = *reinterpret_cast<const uint4*>(scales + (int64_t)g * N + (wc + w) * 8);
  }
  uint32_t q[ROWS][WPT];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    if (r0 + i < K) {
      if constexpr (WPT == 4) {
        uint4 t = ldg_stream_u4(qweight + (int64_t)(r0 + i) * NW + wc);
        q[i][0] = t.x; q[i][1] = t.y; q[i][2] = t.z; q[i][3] = t.w;
      } else {
#pragma unroll
        for (int w = 0; w < WPT; ++w) q[i][w] = ldg_stream_u1(qweight + (int64_t)(r0 + i) * NW + wc + w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    if (r0 + i < K) {
#pragma unroll
      for (int w = 0; w < WPT; ++w) {
        uint4 o = awq_dequant_word(q[i][w], zp[w], sc[w]);
        *reinterpret_cast<uint4*>(out + (int64_t)(r0 + i) * N + (wc + w) * 8) = o;
      }
    }
  }
}

cudaError_t dequantize_gemm(const int32_t* qweight, const void* scales, const int32_t* qzeros, void* out, int K,
                            int N, int G, cudaStream_t st) {
  const int NW = N / 8;
  const bool vec = (N % 32) == 0 && (reinterpret_cast<uintptr_t>(qweight) % 16) == 0;
  const bool rows8 = (G % 8) == 0;
  const int wpt = vec ? 4 : 1;
  const int rows = rows8 ? 8 : 1;
  const int64_t threads = (int64_t)(NW / wpt) * ((K + rows - 1) / rows);
  const int blocks = static_cast<int>((threads + 255) / 256);
  const __half* s = reinterpret_cast<const __half*>(scales);
  __half* o = reinterpret_cast<__half*>(out);
  const dim3 grid(blocks), block(256);
  if (vec && rows8) return launch_kernel(dequant_gemm_kernel<4, 8>, grid, block, 0, st, qweight, s, qzeros, o, K, N, G);
  if (vec) return launch_kernel(dequant_gemm_kernel<4, 1>, grid, block, 0, st, qweight, s, qzeros, o, K, N, G);
  if (rows8) return launch_kernel(dequant_gemm_kernel<1, 8>, grid, block, 0, st, qweight, s, qzeros, o, K, N, G);
  return launch_kernel(dequant_gemm_kernel<1, 1>, grid, block, 0, st, qweight, s, qzeros, o, K, N, G);
}

}  // namespace b200awq

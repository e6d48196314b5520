// HBM streaming micro-benchmark: how fast can 148 persistent CTAs pull a row-major [K, rowbytes] matrix
// through a bulk-copy shared-memory ring, as a function of the contiguous segment width and tile order?
// (Design input for the GEMV: the AWQ GEMM layout is contiguous along N only.)
//   membw <K> <rowbytes> <segbytes> <rows_per_tile> <order 0|1> <stages> [iters]
//   order 0: tiles column-block major (a CTA owns a contiguous run: same columns, consecutive rows)
//   order 1: tiles row major (consecutive tiles = adjacent column blocks of the same rows)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

__global__ void __launch_bounds__(64, 1) k_stream(const uint8_t* __restrict__ src, int K, int rowbytes, int seg, int rpt, int order, int stages, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tile_bytes = seg * rpt;
  uint64_t* full = (uint64_t*)(sm + (size_t)stages * tile_bytes);
  uint64_t* empty = full + stages;
  const int ncb = rowbytes / seg, nkt = K / rpt;
  const long long T = (long long)ncb * nkt;
  const long long t0 = T * blockIdx.x / gridDim.x, t1 = T * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // producer warp: lane r issues row r of the tile (rpt <= 32) ; lane 0 arms the barrier
    long long i = 0;
    for (long long t = t0; t < t1; ++t, ++i) {
      const int st = (int)(i % stages);
      const uint32_t ph = (uint32_t)((i / stages) & 1);
      if (threadIdx.x == 0) { mb_wait(&empty[st], ph ^ 1); mb_expect(&full[st], tile_bytes); }
      __syncwarp();
      long long cb, kt;
      if (order == 0) { cb = t / nkt; kt = t % nkt; } else { kt = t / ncb; cb = t % ncb; }
      for (int r = threadIdx.x; r < rpt; r += 32)
        bulk(sm + (size_t)st * tile_bytes + (size_t)r * seg, src + ((size_t)(kt * rpt + r)) * rowbytes + (size_t)cb * seg, seg, &full[st]);
    }
  } else if (threadIdx.x == 32) {
    long long i = 0;
    unsigned long long acc = 0;
    for (long long t = t0; t < t1; ++t, ++i) {
      const int st = (int)(i % stages);
      const uint32_t ph = (uint32_t)((i / stages) & 1);
      mb_wait(&full[st], ph);
      acc += *(volatile unsigned long long*)(sm + (size_t)st * tile_bytes);
      mb_arrive(&empty[st]);
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

int main(int argc, char** argv) {
  if (argc < 7) { printf("usage\n"); return 1; }
  const int K = atoi(argv[1]), rowbytes = atoi(argv[2]), seg = atoi(argv[3]), rpt = atoi(argv[4]), order = atoi(argv[5]), stages = atoi(argv[6]);
  const int iters = argc > 7 ? atoi(argv[7]) : 20;
  const size_t bytes = (size_t)K * rowbytes;
  const int nbuf = (int)(600000000ull / bytes) + 2;
  uint8_t* d; unsigned long long* sink;
  cudaMalloc(&d, bytes * nbuf); cudaMemset(d, 1, bytes * nbuf); cudaMalloc(&sink, 8);
  const size_t smem = (size_t)stages * seg * rpt + 2 * stages * 8 + 64;
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k_stream<<<sms, 64, smem>>>(d + (size_t)(w % nbuf) * bytes, K, rowbytes, seg, rpt, order, stages, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) k_stream<<<sms, 64, smem>>>(d + (size_t)(it % nbuf) * bytes, K, rowbytes, seg, rpt, order, stages, sink);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("K=%d rowbytes=%d seg=%d rpt=%d order=%d stages=%d smem=%zu : %.2f us/iter  %.1f GB/s  (%s)\n", K, rowbytes, seg, rpt, order, stages, smem, ms * 1e3 / iters, bytes / (ms * 1e-3 / iters) / 1e9, cudaGetErrorString(err));
  return 0;
}

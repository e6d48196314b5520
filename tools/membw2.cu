// Read-bandwidth micro-benchmarks, part 2 (LDG vs TMA-2D issue strategies).  See membw.cu.
//   membw2 <mode> <K> <rowbytes> [p1] [p2]
//   mode 0: LDG.128 fully contiguous, grid = SMs*p1 CTAs of 256 threads, unroll 8
//   mode 1: LDG.128, 128-byte row segments (8 lanes x 16 B, 4 rows per warp instruction), tile = 256 columns x 64 rows
//   mode 2: TMA 2-D box {p1 bytes wide x p2 rows}, 8 producer lanes each with 2 private stages (like gemv v3)
//   mode 3: TMA 2-D box {p1 x p2}, one shared ring of 16 stages, single producer lane
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
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
__device__ __forceinline__ void tma2d(void* dst, const void* tm, uint64_t* b, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)), "l"(tm), "r"(s32(b)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint4 ldg4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__global__ void __launch_bounds__(256) k_ldg_contig(const uint4* __restrict__ src, size_t n16, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n16; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ldg4(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n16; i += stride) { uint4 v = ldg4(src + i); acc ^= v.x ^ v.w; }
  if (acc == 0x12345) *sink = acc;
}

// tile = 256 columns (128 B) x 64 rows per warp-iteration; warps grid-stride over tiles
__global__ void __launch_bounds__(256) k_ldg_seg(const uint8_t* __restrict__ src, int K, int rowbytes, unsigned* sink) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int ncb = rowbytes / 128, nkt = K / 64;
  const long long T = (long long)ncb * nkt;
  const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
  unsigned acc = 0;
  for (long long t = gw; t < T; t += nw) {
    const long long cb = t / nkt, kt = t % nkt;
    const uint8_t* base = src + (size_t)(kt * 64) * rowbytes + (size_t)cb * 128 + g * 16;
    uint4 v[16];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int r = 0; r < 4; ++r) v[b * 4 + r] = ldg4(base + (size_t)(16 * b + 4 * tig + r) * rowbytes);
#pragma unroll
    for (int u = 0; u < 16; ++u) acc ^= v[u].x ^ v[u].w;
  }
  if (acc == 0x12345) *sink = acc;
}

__global__ void __launch_bounds__(288, 1) k_tma(const __grid_constant__ CUtensorMap tm, int K, int rowbytes, int bw, int br, int shared_ring, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tile_bytes = bw * br;
  const int NS = 16;
  const int stage_bytes = (tile_bytes + 1023) & ~1023;
  uint64_t* full = (uint64_t*)(sm + (size_t)NS * stage_bytes);
  uint64_t* empty = full + NS;
  const int ncb = rowbytes / bw, nkt = K / br;
  const long long T = (long long)ncb * nkt;
  const long long t0 = T * blockIdx.x / gridDim.x, t1 = T * (blockIdx.x + 1) / gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long ntile = t1 - t0;
  if (shared_ring) {
    if (threadIdx.x == 0) {
      long long i = 0;
      for (long long t = t0; t < t1; ++t, ++i) {
        const int st = (int)(i % NS); const uint32_t ph = (uint32_t)((i / NS) & 1);
        mb_wait(&empty[st], ph ^ 1); mb_expect(&full[st], tile_bytes);
        const long long cb = t / nkt, kt = t % nkt;
        tma2d(sm + (size_t)st * stage_bytes, &tm, &full[st], (int)(cb * (bw / 4)), (int)(kt * br));
      }
    } else if (threadIdx.x == 32) {
      long long i = 0; unsigned long long acc = 0;
      for (long long t = t0; t < t1; ++t, ++i) {
        const int st = (int)(i % NS); const uint32_t ph = (uint32_t)((i / NS) & 1);
        mb_wait(&full[st], ph); acc += *(volatile unsigned long long*)(sm + (size_t)st * stage_bytes); mb_arrive(&empty[st]);
      }
      if (acc == 0x1234567) *sink = acc;
    }
    return;
  }
  if (warp == 0) {
    if (lane < 8) {
      const int w = lane;
      const long long a = t0 + ntile * w / 8, bnd = t0 + ntile * (w + 1) / 8;
      long long j = 0;
      for (long long t = a; t < bnd; ++t, ++j) {
        const int st = w * 2 + (int)(j & 1); const uint32_t ph = (uint32_t)((j >> 1) & 1);
        mb_wait(&empty[st], ph ^ 1); mb_expect(&full[st], tile_bytes);
        const long long cb = t / nkt, kt = t % nkt;
        tma2d(sm + (size_t)st * stage_bytes, &tm, &full[st], (int)(cb * (bw / 4)), (int)(kt * br));
      }
    }
  } else if (lane == 0) {
    const int w = warp - 1;
    const long long a = t0 + ntile * w / 8, bnd = t0 + ntile * (w + 1) / 8;
    long long j = 0; unsigned long long acc = 0;
    for (long long t = a; t < bnd; ++t, ++j) {
      const int st = w * 2 + (int)(j & 1); const uint32_t ph = (uint32_t)((j >> 1) & 1);
      mb_wait(&full[st], ph); acc += *(volatile unsigned long long*)(sm + (size_t)st * stage_bytes); mb_arrive(&empty[st]);
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = atoi(argv[1]), K = atoi(argv[2]), rowbytes = atoi(argv[3]);
  const int p1 = argc > 4 ? atoi(argv[4]) : 0, p2 = argc > 5 ? atoi(argv[5]) : 0;
  const size_t bytes = (size_t)K * rowbytes;
  const int nbuf = (int)(600000000ull / bytes) + 2, iters = 20;
  uint8_t* d; unsigned long long* sink;
  cudaMalloc(&d, bytes * nbuf); cudaMemset(d, 1, bytes * nbuf); cudaMalloc(&sink, 8);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CUtensorMap tms[64];
  size_t smem = 0;
  if (mode >= 2) {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncFn enc = (EncFn)f;
    for (int b = 0; b < nbuf && b < 64; ++b) {
      cuuint64_t dims[2] = {(cuuint64_t)rowbytes / 4, (cuuint64_t)K}; cuuint64_t str[1] = {(cuuint64_t)rowbytes};
      cuuint32_t box[2] = {(cuuint32_t)p1 / 4, (cuuint32_t)p2}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tms[b], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d + (size_t)b * bytes, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       p1 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    smem = 16 * (size_t)(((p1 * p2) + 1023) & ~1023) + 2 * 16 * 8 + 64;
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  auto run = [&](int it) {
    const uint8_t* src = d + (size_t)(it % nbuf) * bytes;
    if (mode == 0) k_ldg_contig<<<sms * (p1 ? p1 : 8), 256>>>((const uint4*)src, bytes / 16, (unsigned*)sink);
    else if (mode == 1) k_ldg_seg<<<sms * (p1 ? p1 : 8), 256>>>(src, K, rowbytes, (unsigned*)sink);
    else k_tma<<<sms, 288, smem>>>(tms[it % (nbuf < 64 ? nbuf : 64)], K, rowbytes, p1, p2, mode == 3, sink);
  };
  for (int w = 0; w < 3; ++w) run(w);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) run(it);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("mode=%d K=%d rowbytes=%d p1=%d p2=%d : %.2f us/iter %.1f GB/s (%s)\n", mode, K, rowbytes, p1, p2, ms * 1e3 / iters, bytes / (ms * 1e-3 / iters) / 1e9, cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Tensor-core path (M > 8 tokens): Y[M, N] = X[M, K] . deq(W), fp16 x fp16 -> fp32 in TMEM (tcgen05).
//
// Orientation ("swap-AB"): the MMA's M dimension is the OUTPUT-FEATURE axis n (128 per tile, the only
// M the 1-CTA UMMA runs at full rate), the MMA's N dimension is the token axis (BT = 32..256 per tile).
//   A operand = W^T tile [128 n x 64 k]  - produced IN-KERNEL: 4 warps read packed int4 words, dequantise
//               in registers (bit-exact with the dequant kernel) and store fp16 into shared memory in the
//               UMMA canonical 128B-swizzled layout (MN-major for the GEMM layout, whose words hold 8
//               consecutive n of one k; K-major for the GEMV / GEMVFast layouts, whose words hold
//               consecutive k of one n).
//   B operand = X tile [BT tokens x 64 k], K-major, 128B swizzle, loaded by TMA (cp.async.bulk.tensor.2d).
//   D         = [128 lanes (n) x BT columns (tokens)] fp32 in tensor memory.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = dequant producers, then epilogue (tcgen05.ld -> +bias -> fp16 -> global).
// Pipeline: NS smem stages, one "full" mbarrier per stage (128 dequant arrivals + 1 TMA expect_tx),
// one "empty" mbarrier per stage (tcgen05.commit), tmem_full / tmem_empty between MMA and epilogue.
// Persistent CTAs walk (n_tile, m_tile, k_split) work items.  Split-K (only for M <= 64, where the
// problem is HBM-bound and 148 SMs must all stream weights) reduces through fp32 atomics into the
// caller's zeroed workspace; the last CTA of a tile rounds to fp16 and restores the zeros.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace b200awq {

constexpr int kTileN = 128;  // output features per tile (UMMA M)
constexpr int kBK = 64;      // k per pipeline stage (one 128B swizzle row of fp16)
constexpr int kAStageBytes = kTileN * kBK * 2;  // 16 KB

struct TcParams {
  const int32_t* qweight;
  const __half* scales;
  const int32_t* qzeros;  // FAST layout: scaled zeros (fp16) reinterpret
  const __half* bias;
  __half* y;
  float* acc_ws;
  int* tickets;
  int M, K, N, G;
  int zw;          // GEMV layout: zeros width
  int n_tiles, m_tiles, ksplit;
  int a_desc_variant;
};

template <int BT>
struct TcCfg {
  static constexpr int kXStageBytes = BT * kBK * 2;
  static constexpr int kStageBytes = kAStageBytes + kXStageBytes;
  static constexpr int kStages = (BT >= 256) ? 4 : (BT >= 128 ? 6 : 8);
  static constexpr int kTmemCols = BT < 32 ? 32 : BT;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------- A-tile producers
// GEMM layout: thread dt owns word column c = dt % 16 (8 n) and rows kk = dt/16 + 8 j.
struct GemmLayoutLoader {
  uint32_t q[8];
  __device__ __forceinline__ void load(const TcParams& p, int nt, int k0, int dt) {
    const int NW = p.N >> 3;
    const int c = dt & 15, rb = dt >> 4;
    const int wc = nt * 16 + c;
    const bool ok = wc < NW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      q[j] = 0u;
      if (ok) q[j] = ldg_stream_u1(p.qweight + (int64_t)(k0 + rb + 8 * j) * NW + wc);
    }
  }
  __device__ __forceinline__ void store(const TcParams& p, int nt, int k0, int dt, uint8_t* a_stage) const {
    const int NW = p.N >> 3;
    const int c = dt & 15, rb = dt >> 4;
    const int wc = nt * 16 + c;
    const bool ok = wc < NW;
    int cur_g = -1;
    ZeroPairs zp;
    uint4 sc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (k0 + rb + 8 * j) / p.G;
      if (g != cur_g) {
        cur_g = g;
        if (ok) {
          zp = awq_zero_pairs(static_cast<uint32_t>(p.qzeros[(int64_t)g * NW + wc]));
          sc = *reinterpret_cast<const uint4*>(p.scales + (int64_t)g * p.N + wc * 8);
        } else {
          zp = awq_zero_pairs(0u);
        }
      }
      uint4 o = awq_dequant_word(q[j], zp, sc);
      if (!ok) o = make_uint4(0, 0, 0, 0);
      // MN-major SW128: (n/64)*8192 + (k/8)*1024 + (k%8)*128 + (((n%64)/8) ^ (k%8))*16 ; k = rb + 8j
      const uint32_t off = (uint32_t)(c >> 3) * 8192u + (uint32_t)j * 1024u + (uint32_t)rb * 128u +
                           (uint32_t)(((c & 7) ^ rb) << 4);
      *reinterpret_cast<uint4*>(a_stage + off) = o;
    }
  }
};

// GEMV layout: thread dt owns k-word cw = dt % 8 (8 consecutive k) and rows n = dt/8 + 16 j.
struct GemvLayoutLoader {
  uint32_t q[8];
  __device__ __forceinline__ void load(const TcParams& p, int nt, int k0, int dt) {
    const int KW = p.K >> 3;
    const int cw = dt & 7, rb = dt >> 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = nt * kTileN + rb + 16 * j;
      q[j] = 0u;
      if (n < p.N) q[j] = ldg_stream_u1(p.qweight + (int64_t)n * KW + (k0 >> 3) + cw);
    }
  }
  __device__ __forceinline__ void store(const TcParams& p, int nt, int k0, int dt, uint8_t* a_stage) const {
    const int cw = dt & 7, rb = dt >> 3;
    const int g = (k0 + cw * 8) / p.G;
    const __half2 r16 = __float2half2_rn(0.0625f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int nl = rb + 16 * j;
      const int n = nt * kTileN + nl;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (n < p.N) {
        const __half s = p.scales[(int64_t)n * (p.zw * 8) + g];
        const uint32_t zword = static_cast<uint32_t>(p.qzeros[(int64_t)n * p.zw + (g >> 3)]);
        const float zf = static_cast<float>((zword >> (4 * (g & 7))) & 0xFu);
        const __half2 zA = __float2half2_rn(1024.f + zf);   // exact
        const __half2 zB = __float2half2_rn(-(64.f + zf));  // exact
        const __half2 s2 = __half2half2(s);
        RawPairs r = awq_raw_pairs(q[j]);
        // pairs (k0,k4) (k1,k5) (k2,k6) (k3,k7)
        const uint32_t d0 = h2_as_u32(__hmul2(__hsub2(u32_as_h2(r.p[0]), zA), s2));
        const uint32_t d1 = h2_as_u32(__hmul2(__hfma2(u32_as_h2(r.p[1]), r16, zB), s2));
        const uint32_t d2 = h2_as_u32(__hmul2(__hsub2(u32_as_h2(r.p[2]), zA), s2));
        const uint32_t d3 = h2_as_u32(__hmul2(__hfma2(u32_as_h2(r.p[3]), r16, zB), s2));
        o.x = __byte_perm(d0, d1, 0x5410);  // (k0, k1)
        o.y = __byte_perm(d2, d3, 0x5410);  // (k2, k3)
        o.z = __byte_perm(d0, d1, 0x7632);  // (k4, k5)
        o.w = __byte_perm(d2, d3, 0x7632);  // (k6, k7)
      }
      // K-major SW128: row n * 128 B, 16-byte chunk (k/8) ^ (n % 8)
      const uint32_t off = (uint32_t)nl * 128u + (uint32_t)((cw ^ (nl & 7)) << 4);
      *reinterpret_cast<uint4*>(a_stage + off) = o;
    }
  }
};

// GEMVFast layout: thread dt owns row n = dt of the tile (row group dt/4, run dt%4): 2 x 16 B = 64 k.
struct FastLayoutLoader {
  uint4 q[2];
  __device__ __forceinline__ void load(const TcParams& p, int nt, int k0, int dt) {
    const int n = nt * kTileN + dt;
    q[0] = q[1] = make_uint4(0, 0, 0, 0);
    if (n < p.N) {
      const int16_t* base = reinterpret_cast<const int16_t*>(p.qweight) + (int64_t)(n >> 2) * p.K + (int64_t)k0 +
                            (n & 3) * 16;  // block k0/64 starts at int16 offset k0
      q[0] = ldg_stream_u4(base);
      q[1] = ldg_stream_u4(base + 8);
    }
  }
  __device__ __forceinline__ void store(const TcParams& p, int nt, int k0, int dt, uint8_t* a_stage) const {
    const int n = nt * kTileN + dt;
    const __half* sz_ptr = reinterpret_cast<const __half*>(p.qzeros);
    const __half2 r16 = __float2half2_rn(0.0625f);
    const __half2 m1024 = __float2half2_rn(1024.f), m64 = __float2half2_rn(-64.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      __half2 s2 = __float2half2_rn(0.f), z2 = s2;
      if (n < p.N) {
        const int g = (k0 + 32 * h) / p.G;
        s2 = __half2half2(p.scales[(int64_t)g * p.N + n]);
        z2 = __half2half2(sz_ptr[(int64_t)g * p.N + n]);
      }
      const uint32_t w[4] = {q[h].x, q[h].y, q[h].z, q[h].w};
      uint32_t d[4][4];  // [word u][r']
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        RawPairs r = awq_raw_pairs(w[u]);
        d[u][0] = h2_as_u32(__hfma2(__hsub2(u32_as_h2(r.p[0]), m1024), s2, z2));
        d[u][1] = h2_as_u32(__hfma2(__hfma2(u32_as_h2(r.p[1]), r16, m64), s2, z2));
        d[u][2] = h2_as_u32(__hfma2(__hsub2(u32_as_h2(r.p[2]), m1024), s2, z2));
        d[u][3] = h2_as_u32(__hfma2(__hfma2(u32_as_h2(r.p[3]), r16, m64), s2, z2));
      }
#pragma unroll
      for (int rp = 0; rp < 4; ++rp) {
        uint4 o = make_uint4(d[0][rp], d[1][rp], d[2][rp], d[3][rp]);  // k = 32h + 8rp + 0..7
        if (n >= p.N) o = make_uint4(0, 0, 0, 0);
        const int cc = 4 * h + rp;
        const uint32_t off = (uint32_t)dt * 128u + (uint32_t)((cc ^ (dt & 7)) << 4);
        *reinterpret_cast<uint4*>(a_stage + off) = o;
      }
    }
  }
};

template <int LAYOUT>
struct LoaderOf;
template <> struct LoaderOf<0> { using T = GemmLayoutLoader; };
template <> struct LoaderOf<1> { using T = GemvLayoutLoader; };
template <> struct LoaderOf<2> { using T = FastLayoutLoader; };

// --------------------------------------------------------------------------------------- kernel
template <int BT, int LAYOUT>
__global__ void __launch_bounds__(192, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmx, const TcParams p) {
  using Cfg = TcCfg<BT>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;                                  // NS x 16 KB
  uint8_t* x_base = smem + (size_t)NS * kAStageBytes;      // NS x BT*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * Cfg::kStageBytes);
  uint64_t* full = bars;            // [NS]
  uint64_t* empty = bars + NS;      // [NS]
  uint64_t* tmem_full = bars + 2 * NS;
  uint64_t* tmem_empty = bars + 2 * NS + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
  int* s_flag = reinterpret_cast<int*>(bars + 2 * NS + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 128 + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int KS = p.K / kBK;  // k-steps in total
  const int n_work = p.n_tiles * p.m_tiles * p.ksplit;

  if (warp == 0) {
    // ================================================================= TMA producer (X tiles)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int ks = w % p.ksplit;
        const int mt = (w / p.ksplit) % p.m_tiles;
        const int s_begin = (int)((int64_t)KS * ks / p.ksplit), s_end = (int)((int64_t)KS * (ks + 1) / p.ksplit);
        for (int s = s_begin; s < s_end; ++s) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::kXStageBytes);
          tma_load_2d(x_base + (size_t)stage * Cfg::kXStageBytes, &tmx, &full[stage], s * kBK, mt * BT);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileN, BT, LAYOUT == 0 ? 1 : 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int ks = w % p.ksplit;
        const int s_begin = (int)((int64_t)KS * ks / p.ksplit), s_end = (int)((int64_t)KS * (ks + 1) / p.ksplit);
        mbar_wait(tmem_empty, acc_phase ^ 1);  // epilogue has drained the accumulator
        tc_fence_after();
        for (int s = s_begin; s < s_end; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_base + (size_t)stage * kAStageBytes);
          const uint32_t x_addr = smem_u32(x_base + (size_t)stage * Cfg::kXStageBytes);
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16) {
            uint64_t da, db;
            if (LAYOUT == 0) {
              // MN-major, SW128: LBO = stride between 64-n atoms (8192), SBO = stride between 8-k atoms (1024)
              if (p.a_desc_variant == 0)
                da = umma_smem_desc(a_addr + k16 * 2048, 8192, 1024);
              else
                da = umma_smem_desc(a_addr + k16 * 2048, 1024, 8192);
            } else {
              da = umma_smem_desc(a_addr + k16 * 32, 16, 1024);  // K-major SW128
            }
            db = umma_smem_desc(x_addr + k16 * 32, 16, 1024);    // K-major SW128
            umma_f16_ss(tmem_base, da, db, idesc, (s > s_begin || k16 > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees this smem stage when the MMAs above retire
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full);  // accumulator complete
        acc_phase ^= 1;
      }
    }
  } else {
    // ================================================================= dequant producers + epilogue
    const int dt = threadIdx.x - 64;  // 0..127
    const int q4 = warp & 3;          // TMEM lane quadrant this warp may read
    typename LoaderOf<LAYOUT>::T cur, nxt;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int ks = w % p.ksplit;
      const int mt = (w / p.ksplit) % p.m_tiles;
      const int nt = w / (p.ksplit * p.m_tiles);
      const int s_begin = (int)((int64_t)KS * ks / p.ksplit), s_end = (int)((int64_t)KS * (ks + 1) / p.ksplit);
      if (s_begin < s_end) nxt.load(p, nt, s_begin * kBK, dt);
      for (int s = s_begin; s < s_end; ++s) {
        cur = nxt;
        if (s + 1 < s_end) nxt.load(p, nt, (s + 1) * kBK, dt);  // next step's words in flight
        mbar_wait(&empty[stage], phase ^ 1);
        cur.store(p, nt, s * kBK, dt, a_base + (size_t)stage * kAStageBytes);
        fence_proxy_async_smem();
        mbar_arrive(&full[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      // ---- epilogue for this work item
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after();
      const int n = nt * kTileN + q4 * 32 + lane;
      const bool n_ok = n < p.N;
      const float bias_v = (p.bias != nullptr && n_ok) ? __half2float(p.bias[n]) : 0.f;
      const int m0 = mt * BT;
#pragma unroll 1
      for (int c0 = 0; c0 < BT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (n_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = m0 + c0 + j;
            if (m < p.M) {
              const float f = __uint_as_float(v[j]);
              if (p.ksplit == 1)
                p.y[(int64_t)m * p.N + n] = __float2half_rn(f + bias_v);
              else
                atomicAdd(&p.acc_ws[(int64_t)m * p.N + n], f);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      acc_phase ^= 1;
      if (p.ksplit > 1) {
        __threadfence();
        named_bar_sync(1, 128);
        if (dt == 0) {
          const int prev = atomicAdd(&p.tickets[nt * p.m_tiles + mt], 1);
          *s_flag = (prev == p.ksplit - 1);
        }
        named_bar_sync(1, 128);
        const bool last = *s_flag != 0;
        named_bar_sync(1, 128);  // everyone has read the flag before a later item rewrites it
        if (last) {
          __threadfence();
          if (n_ok) {
            for (int j = 0; j < BT; ++j) {
              const int m = m0 + j;
              if (m >= p.M) break;
              float* a = &p.acc_ws[(int64_t)m * p.N + n];
              const float f = ldcg_f1(a);
              *a = 0.f;
              p.y[(int64_t)m * p.N + n] = __float2half_rn(f + bias_v);
            }
          }
          if (dt == 0) p.tickets[nt * p.m_tiles + mt] = 0;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// -------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t ld;
  int M, K, BT;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && ld == o.ld && M == o.M && K == o.K && BT == o.BT; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.ld * 0x9E3779B97F4A7C15ull + (size_t)k.M * 1315423911u + (size_t)k.K * 2654435761u + (size_t)k.BT;
    return h;
  }
};

// X[M, K] fp16 (row pitch ld elements) -> box {64 k, BT rows}, 128B swizzle, zero fill out of bounds
static cudaError_t make_x_tmap(const void* x, int64_t ld, int M, int K, int BT, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{x, ld, M, K, BT};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return cudaSuccess;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return cudaErrorNotSupported;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BT};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return cudaSuccess;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = B200AWQ_SM_COUNT_FALLBACK;
  }
  return n;
}

template <int BT, int LAYOUT>
static cudaError_t launch_tc(const CUtensorMap& tm, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BT>;
  auto kern = gemm_tc_kernel<BT, LAYOUT>;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int n_work = p.n_tiles * p.m_tiles * p.ksplit;
  const int grid = n_work < sm_count() ? n_work : sm_count();
  kern<<<grid, 192, Cfg::kSmemBytes, st>>>(tm, p);
  return cudaGetLastError();
}

template <int LAYOUT>
static cudaError_t dispatch_bt(int BT, const CUtensorMap& tm, const TcParams& p, cudaStream_t st) {
  switch (BT) {
    case 32: return launch_tc<32, LAYOUT>(tm, p, st);
    case 64: return launch_tc<64, LAYOUT>(tm, p, st);
    case 128: return launch_tc<128, LAYOUT>(tm, p, st);
    default: return launch_tc<256, LAYOUT>(tm, p, st);
  }
}

static int zeros_width_tc(int K, int G) {
  const int mult = G >= 128 ? 1 : (G == 64 ? 2 : 4);
  int base = ((K / G) + 7) / 8;
  return ((base + mult - 1) / mult) * mult;
}

cudaError_t gemm_tc(const GemmArgs& a, int layout, float* acc_ws, int* tickets, cudaStream_t st) {
  if (a.K % kBK != 0 || a.G % 32 != 0) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0 || (a.ldx % 8) != 0) return cudaErrorMisalignedAddress;
  const int BT = a.M <= 32 ? 32 : (a.M <= 64 ? 64 : (a.M <= 128 ? 128 : 256));
  TcParams p;
  p.qweight = a.qweight;
  p.scales = reinterpret_cast<const __half*>(a.scales);
  p.qzeros = a.qzeros;
  p.bias = reinterpret_cast<const __half*>(a.bias);
  p.y = reinterpret_cast<__half*>(a.y);
  p.acc_ws = acc_ws;
  p.tickets = tickets;
  p.M = a.M; p.K = a.K; p.N = a.N; p.G = a.G;
  p.zw = zeros_width_tc(a.K, a.G);
  p.n_tiles = (a.N + kTileN - 1) / kTileN;
  p.m_tiles = (a.M + BT - 1) / BT;
  p.a_desc_variant = knob(3);
  const int KS = a.K / kBK;
  int ksplit = 1;
  const int tiles = p.n_tiles * p.m_tiles;
  if (a.M <= kMaxSplitM && acc_ws != nullptr && tickets != nullptr && tiles < sm_count() && tiles <= 4096) {
    ksplit = sm_count() / tiles;
    if (ksplit < 1) ksplit = 1;
    while (ksplit > 1 && KS / ksplit < 4) --ksplit;  // at least 4 k-steps per slice
  }
  const int forced = knob(1);
  if (forced > 0 && a.M <= kMaxSplitM && acc_ws != nullptr && tickets != nullptr) ksplit = forced > KS ? KS : forced;
  p.ksplit = ksplit;
  CUtensorMap tm;
  cudaError_t e = make_x_tmap(a.x, a.ldx, a.M, a.K, BT, &tm);
  if (e != cudaSuccess) return e;
  switch (layout) {
    case 0: return dispatch_bt<0>(BT, tm, p, st);
    case 1: return dispatch_bt<1>(BT, tm, p, st);
    default: return dispatch_bt<2>(BT, tm, p, st);
  }
}

}  // namespace b200awq

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
// Warp roles (448 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> +bias -> fp16 -> global), warps 6.. = dequant producers (16 for the GEMM layout).
// Pipeline: NS smem stages, one "full" mbarrier per stage (8 producer-warp arrivals + 1 TMA expect_tx),
// one "empty" mbarrier per stage (tcgen05.commit); two TMEM accumulator buffers with tmem_full /
// tmem_empty mbarriers, so the epilogue of one tile overlaps the main loop of the next.
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
  int has_tmq;     // GEMM layout: a tensor map over qweight is available for L2 prefetch
  int g_shift;     // log2(G) when G is a power of two, else 31 (G == K: one group) - no integer division on device
};

template <int BT>
struct TcCfg {
  static constexpr int kXStageBytes = BT * kBK * 2;
  static constexpr int kStageBytes = kAStageBytes + kXStageBytes;
  static constexpr int kStages = (BT >= 256) ? 4 : (BT >= 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * (BT < 32 ? 32 : BT);  // two accumulator buffers
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------- A-tile producers
// 256 or 512 producer threads.  load() issues the global loads of one k-step (packed words plus,
// when the quantisation group changes, the group's zeros / scales) ONE STEP AHEAD of store(), which
// dequantises from registers and writes the swizzled fp16 tile: no global latency on the critical path.

// GEMM layout: thread dt owns word column c = dt % 16 (8 n) and rows kk = dt/16 + 16 j, j = 0..3.
// (A 512-thread variant with 2 rows per thread measured slower: 725 vs 838 TFLOP/s at M = 4096.)
// NG = quantisation groups per 64-row k-step: 1 for G >= 64, 2 for G == 32 (rows < 32 / >= 32).  Keeping the
// slot small (9 registers for NG = 1) is what allows a 6-deep register prefetch ring.
template <int NG>
struct GemmLayoutLoaderT {
  static constexpr int kThreads = 256;
  static constexpr int kDepth = NG == 1 ? 6 : 4;
  uint32_t q[4];
  uint32_t zq[NG];
  uint4 sc[NG];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int h = 0; h < NG; ++h) { zq[h] = 0; sc[h] = make_uint4(0, 0, 0, 0); }
  }
  __device__ __forceinline__ void load(const TcParams& p, int nt, int k0, int dt) {
    // 32-bit word offsets (K * N/8 < 2^32 for every real shape) keep the address arithmetic to a few
    // instructions per load; the first version spent ~90 instructions per k-step on 64-bit multiplies.
    const uint32_t NW = (uint32_t)p.N >> 3;
    const uint32_t c = dt & 15, rb = dt >> 4;
    const uint32_t wc = (uint32_t)nt * 16 + c;
    const bool ok = wc < NW;
    const uint32_t row0 = ((uint32_t)k0 + rb) * NW + wc;
    const uint32_t row16 = 16u * NW;
    const int32_t* src = p.qweight + row0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = 0u;
      if (ok) q[j] = ldg_stream_u1(src + j * row16);
    }
#pragma unroll
    for (int h = 0; h < NG; ++h) {
      const int g = (k0 + (int)rb + 32 * h) >> p.g_shift;   // G is a power of two, or the single group G == K
      if (ok) {
        const uint32_t goff = (uint32_t)g * NW + wc;
        zq[h] = static_cast<uint32_t>(__ldg(p.qzeros + goff));
        sc[h] = __ldg(reinterpret_cast<const uint4*>(p.scales) + goff);  // 8 halves per word column
      }
    }
  }
  __device__ __forceinline__ void store(const TcParams& p, int nt, int k0, int dt, uint32_t a_stage) const {
    const int c = dt & 15, rb = dt >> 4;
    // columns past N keep q = 0, zeros = 0, scales = 0 from init(): they dequantise to exact zeros
    ZeroPairs zp[NG];
#pragma unroll
    for (int h = 0; h < NG; ++h) zp[h] = awq_zero_pairs(zq[h]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = NG == 1 ? 0 : (j >> 1);
      const uint4 o = awq_dequant_word(q[j], zp[h], sc[h]);
      // MN-major SW128: (n/64)*8192 + (k/8)*1024 + (k%8)*128 + (((n%64)/8) ^ (k%8))*16 ; k = rb + 16 j
      const uint32_t off = (uint32_t)(c >> 3) * 8192u + (uint32_t)(2 * j + (rb >> 3)) * 1024u +
                           (uint32_t)(rb & 7) * 128u + (uint32_t)(((c & 7) ^ (rb & 7)) << 4);
      sts_u4(a_stage + off, o);
    }
  }
};

// GEMV layout: thread dt owns k-word cw = dt % 8 (8 consecutive k) and rows n = dt/8 + 32 j, j = 0..3.
struct GemvLayoutLoader {
  static constexpr int kThreads = 256;
  static constexpr int kDepth = 4;
  uint32_t q[4];
  uint32_t zs[4];  // per row: fp16 scale in the low half, zero-point (0..15) in the high half
  __device__ __forceinline__ void init() { zs[0] = zs[1] = zs[2] = zs[3] = 0; }
  __device__ __forceinline__ void load(const TcParams& p, int nt, int k0, int dt) {
    const int KW = p.K >> 3;
    const int cw = dt & 7, rb = dt >> 3;
    const int g = (k0 + cw * 8) >> p.g_shift;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = nt * kTileN + rb + 32 * j;
      q[j] = 0u;
      if (n < p.N) {
        q[j] = ldg_stream_u1(p.qweight + (int64_t)n * KW + (k0 >> 3) + cw);
        const uint32_t s = __half_as_ushort(__ldg(p.scales + (int64_t)n * (p.zw * 8) + g));
        const uint32_t zword = static_cast<uint32_t>(__ldg(p.qzeros + (int64_t)n * p.zw + (g >> 3)));
        zs[j] = s | (((zword >> (4 * (g & 7))) & 0xFu) << 16);
      }
    }
  }
  __device__ __forceinline__ void store(const TcParams& p, int nt, int k0, int dt, uint32_t a_stage) const {
    const int cw = dt & 7, rb = dt >> 3;
    const __half2 r16 = __float2half2_rn(0.0625f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nl = rb + 32 * j;
      const int n = nt * kTileN + nl;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (n < p.N) {
        const float zf = static_cast<float>(zs[j] >> 16);
        const __half2 zA = __float2half2_rn(1024.f + zf);   // exact
        const __half2 zB = __float2half2_rn(-(64.f + zf));  // exact
        const __half2 s2 = __half2half2(__ushort_as_half(static_cast<unsigned short>(zs[j] & 0xffffu)));
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
      sts_u4(a_stage + off, o);
    }
  }
};

// GEMVFast layout: thread dt owns row n = dt/2 of the tile and the 32-k half h = dt%2 of the step.
struct FastLayoutLoader {
  static constexpr int kThreads = 256;
  static constexpr int kDepth = 4;
  uint4 q;
  uint32_t ss;  // scale (low half) | scaled zero (high half)
  __device__ __forceinline__ void init() { ss = 0; }
  __device__ __forceinline__ void load(const TcParams& p, int nt, int k0, int dt) {
    const int n = nt * kTileN + (dt >> 1), h = dt & 1;
    q = make_uint4(0, 0, 0, 0);
    const int g = (k0 + 32 * h) >> p.g_shift;
    if (n < p.N) {
      // 64-k block k0/64 of row group n/4 starts at int16 offset k0; run (n%4) * 16; half h * 8
      q = ldg_stream_u4(reinterpret_cast<const int16_t*>(p.qweight) + (int64_t)(n >> 2) * p.K + (int64_t)k0 +
                        (n & 3) * 16 + h * 8);
      const __half* sz_ptr = reinterpret_cast<const __half*>(p.qzeros);
      ss = static_cast<uint32_t>(__half_as_ushort(__ldg(p.scales + (int64_t)g * p.N + n))) |
           (static_cast<uint32_t>(__half_as_ushort(__ldg(sz_ptr + (int64_t)g * p.N + n))) << 16);
    }
  }
  __device__ __forceinline__ void store(const TcParams& p, int nt, int k0, int dt, uint32_t a_stage) const {
    const int nl = dt >> 1, h = dt & 1;
    const bool ok = nt * kTileN + nl < p.N;
    const __half2 r16 = __float2half2_rn(0.0625f);
    const __half2 m1024 = __float2half2_rn(1024.f), m64 = __float2half2_rn(-64.f);
    const __half2 s2 = __half2half2(__ushort_as_half(static_cast<unsigned short>(ss & 0xffffu)));
    const __half2 z2 = __half2half2(__ushort_as_half(static_cast<unsigned short>(ss >> 16)));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t d[4][4];  // [word u][r'] : pair (k, k+1), k = 32h + 2u + 8r'
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
      if (!ok) o = make_uint4(0, 0, 0, 0);
      const int cc = 4 * h + rp;
      const uint32_t off = (uint32_t)nl * 128u + (uint32_t)((cc ^ (nl & 7)) << 4);
      sts_u4(a_stage + off, o);
    }
  }
};

template <int LAYOUT>
struct LoaderOf;
template <> struct LoaderOf<0> { using T = GemmLayoutLoaderT<1>; };  // GEMM layout, G >= 64
template <> struct LoaderOf<3> { using T = GemmLayoutLoaderT<2>; };  // GEMM layout, G == 32
template <> struct LoaderOf<1> { using T = GemvLayoutLoader; };
template <> struct LoaderOf<2> { using T = FastLayoutLoader; };

// --------------------------------------------------------------------------------------- kernel
// warps: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..5 = epilogue (TMEM lane quadrant = warp % 4),
// 6..13 = dequant producers.  Two TMEM accumulator buffers: the epilogue of tile i overlaps the main loop
// of tile i+1.
template <int LAYOUT>
constexpr int tc_threads() { return 64 + 128 + LoaderOf<LAYOUT>::T::kThreads; }

template <int BT, int LAYOUT>
__global__ void __launch_bounds__(tc_threads<LAYOUT>(), 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmq,
                   const TcParams p) {
  using Cfg = TcCfg<BT>;
  constexpr int NS = Cfg::kStages;
  constexpr int NPROD = LoaderOf<LAYOUT>::T::kThreads;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;                                  // NS x 16 KB
  uint8_t* x_base = smem + (size_t)NS * kAStageBytes;      // NS x BT*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * Cfg::kStageBytes);
  uint64_t* full = bars;            // [NS]
  uint64_t* empty = bars + NS;      // [NS]
  uint64_t* tmem_full = bars + 2 * NS;       // [2]
  uint64_t* tmem_empty = bars + 2 * NS + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);
  int* s_flag = reinterpret_cast<int*>(bars + 2 * NS + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], NPROD / 32 + 1);  // one elected arrival per producer warp + the TMA expect_tx
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
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
      // GEMM layout: pull the packed weights of the next kL2Ahead k-steps from HBM into L2 (TMA prefetch, no
      // smem destination) so that the producers' register prefetch only has to cover L2 latency.  Weights
      // do not depend on the predecessor kernel: this starts before the PDL wait.
      constexpr int kL2Ahead = 16;
      int wp = blockIdx.x, sp = 0, sp_end = 0;
      bool pf_valid = (LAYOUT == 0 || LAYOUT == 3) && p.has_tmq && wp < n_work;
      auto pf_range = [&]() {
        const int ks = wp % p.ksplit;
        sp = (int)((int64_t)KS * ks / p.ksplit);
        sp_end = (int)((int64_t)KS * (ks + 1) / p.ksplit);
      };
      auto pf_step = [&]() {
        const int nt = wp / (p.ksplit * p.m_tiles);
        tma_prefetch_l2_2d(&tmq, nt * 16, sp * kBK);
        if (++sp == sp_end) {
          wp += gridDim.x;
          if (wp < n_work) pf_range(); else pf_valid = false;
        }
      };
      if (pf_valid) pf_range();
      for (int i = 0; i < kL2Ahead && pf_valid; ++i) pf_step();
      pdl_wait();  // the activations are the predecessor's output
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
          if (pf_valid) pf_step();
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileN, BT, (LAYOUT == 0 || LAYOUT == 3) ? 1 : 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int ks = w % p.ksplit;
        const int s_begin = (int)((int64_t)KS * ks / p.ksplit), s_end = (int)((int64_t)KS * (ks + 1) / p.ksplit);
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tmem_empty[buf], use ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BT);
        for (int s = s_begin; s < s_end; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_base + (size_t)stage * kAStageBytes);
          const uint32_t x_addr = smem_u32(x_base + (size_t)stage * Cfg::kXStageBytes);
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16) {
            uint64_t da, db;
            if (LAYOUT == 0 || LAYOUT == 3) {
              // MN-major, SW128: LBO = stride between 64-n atoms (8192), SBO = stride between 8-k atoms (1024)
              da = umma_smem_desc(a_addr + k16 * 2048, 8192, 1024);
            } else {
              da = umma_smem_desc(a_addr + k16 * 32, 16, 1024);  // K-major SW128
            }
            db = umma_smem_desc(x_addr + k16 * 32, 16, 1024);    // K-major SW128
            umma_f16_ss(d_tmem, da, db, idesc, (s > s_begin || k16 > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees this smem stage when the MMAs above retire
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[buf]);  // accumulator complete
      }
    }
  } else if (warp < 6) {
    // ================================================================= epilogue (4 warps)
    const int et = threadIdx.x - 64;  // 0..127
    const int q4 = warp & 3;          // TMEM lane quadrant this warp may read
    pdl_wait();                       // outputs / workspace may alias memory the predecessor still uses
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int mt = (w / p.ksplit) % p.m_tiles;
      const int nt = w / (p.ksplit * p.m_tiles);
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tmem_full[buf], use);
      tc_fence_after();
      const int n = nt * kTileN + q4 * 32 + lane;
      const bool n_ok = n < p.N;
      const float bias_v = (p.bias != nullptr && n_ok) ? __half2float(p.bias[n]) : 0.f;
      const int m0 = mt * BT;
#pragma unroll 1
      for (int c0 = 0; c0 < BT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BT + c0), v);
        tmem_ld_wait();
        if (n_ok && m0 + c0 < p.M) {
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
      mbar_arrive(&tmem_empty[buf]);
      if (p.ksplit > 1) {
        __threadfence();
        named_bar_sync(1, 128);
        if (et == 0) {
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
          if (et == 0) p.tickets[nt * p.m_tiles + mt] = 0;
        }
      }
    }
  } else {
    // ================================================================= dequant producers (8 warps)
    // Global loads run kPrefetch k-steps ahead of the dequantisation (register ring): a k-step is only
    // ~500 MMA cycles, well below the DRAM / L2 latency a 1-deep prefetch would expose every step.
    constexpr int kPrefetch = LoaderOf<LAYOUT>::T::kDepth;
    const int dt = threadIdx.x - 192;  // 0..255
    const uint32_t a_base_s = smem_u32(a_base);
    typename LoaderOf<LAYOUT>::T ring[kPrefetch];
    // The (work item, k-step) sequence of this CTA is ONE stream: the load cursor runs kPrefetch steps ahead
    // of the store cursor straight through tile boundaries, so the ring never drains between tiles.
    struct Cursor {
      int w, s, s_end, nt;
      bool valid;
    };
    auto set_range = [&](Cursor& c) {
      const int ks = c.w % p.ksplit;
      c.nt = c.w / (p.ksplit * p.m_tiles);
      c.s = (int)((int64_t)KS * ks / p.ksplit);
      c.s_end = (int)((int64_t)KS * (ks + 1) / p.ksplit);
    };
    auto advance = [&](Cursor& c) {
      if (++c.s == c.s_end) {
        c.w += gridDim.x;
        c.valid = c.w < n_work;
        if (c.valid) set_range(c);
      }
    };
    Cursor L, S;
    L.w = S.w = blockIdx.x;
    L.valid = S.valid = blockIdx.x < n_work;
    if (L.valid) { set_range(L); set_range(S); }
#pragma unroll
    for (int d = 0; d < kPrefetch; ++d) {
      ring[d].init();
      if (L.valid) {
        ring[d].load(p, L.nt, L.s * kBK, dt);
        advance(L);
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    while (S.valid) {
#pragma unroll
      for (int d = 0; d < kPrefetch; ++d) {
        if (S.valid) {
          mbar_wait(&empty[stage], phase ^ 1);
          ring[d].store(p, S.nt, S.s * kBK, dt, a_base_s + (uint32_t)stage * kAStageBytes);
          fence_proxy_async_smem();   // every writer: generic-proxy stores -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);  // 8 arrivals per stage instead of 256
          if (++stage == NS) { stage = 0; phase ^= 1; }
          advance(S);
          if (L.valid) {
            ring[d].load(p, L.nt, L.s * kBK, dt);
            advance(L);
          }
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
  uint64_t inner, outer, pitch;
  uint32_t bi, bo;
  int kind;  // bit 0: element type, bit 1: 128B swizzle
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && pitch == o.pitch && bi == o.bi && bo == o.bo &&
           kind == o.kind;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.inner * 0x9E3779B97F4A7C15ull + (size_t)k.outer * 1315423911u + (size_t)k.pitch * 2654435761u +
         (size_t)k.bi * 31u + (size_t)k.bo * 131u + (size_t)k.kind;
    return h;
  }
};

cudaError_t make_tmap_2d(const void* ptr, int elem_kind, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                         uint32_t box_inner, uint32_t box_outer, CUtensorMap* out, bool swizzle128) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, inner, outer, pitch_bytes, box_inner, box_outer, elem_kind | (swizzle128 ? 2 : 0)};
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
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_INT32;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return cudaSuccess;
}

// X[M, K] fp16 (row pitch ld elements) -> box {64 k, BT rows}
static cudaError_t make_x_tmap(const void* x, int64_t ld, int M, int K, int BT, CUtensorMap* out) {
  return make_tmap_2d(x, 0, (uint64_t)K, (uint64_t)M, (uint64_t)ld * 2, kBK, (uint32_t)BT, out);
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
static cudaError_t launch_tc(const CUtensorMap& tm, const CUtensorMap& tmq, const TcParams& p, cudaStream_t st) {
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
  return launch_kernel(kern, dim3(grid), dim3(tc_threads<LAYOUT>()), Cfg::kSmemBytes, st, tm, tmq, p);
}

template <int LAYOUT>
static cudaError_t dispatch_bt(int BT, const CUtensorMap& tm, const CUtensorMap& tmq, const TcParams& p,
                               cudaStream_t st) {
  switch (BT) {
    case 32: return launch_tc<32, LAYOUT>(tm, tmq, p, st);
    case 64: return launch_tc<64, LAYOUT>(tm, tmq, p, st);
    case 128: return launch_tc<128, LAYOUT>(tm, tmq, p, st);
    default: return launch_tc<256, LAYOUT>(tm, tmq, p, st);
  }
}

static int zeros_width_tc(int K, int G) {
  const int mult = G >= 128 ? 1 : (G == 64 ? 2 : 4);
  int base = ((K / G) + 7) / 8;
  return ((base + mult - 1) / mult) * mult;
}

cudaError_t gemm_tc(const GemmArgs& a, int layout, float* acc_ws, int* tickets, cudaStream_t st) {
  if (a.K % kBK != 0 || a.G % 32 != 0) return cudaErrorNotSupported;
  const bool g_pow2 = (a.G & (a.G - 1)) == 0;
  if (!g_pow2 && a.G != a.K) return cudaErrorNotSupported;  // AWQ group sizes: 32 / 64 / 128 / whole row
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
  p.g_shift = 31;
  if (g_pow2) {
    p.g_shift = 0;
    while ((1 << p.g_shift) < a.G) ++p.g_shift;
  }
  p.n_tiles = (a.N + kTileN - 1) / kTileN;
  p.m_tiles = (a.M + BT - 1) / BT;
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
  CUtensorMap tm, tmq;
  cudaError_t e = make_x_tmap(a.x, a.ldx, a.M, a.K, BT, &tm);
  if (e != cudaSuccess) return e;
  tmq = tm;
  p.has_tmq = 0;
  if (layout == 0 && (a.N % 4) == 0 && (reinterpret_cast<uintptr_t>(a.qweight) & 15) == 0 && (a.N / 8) % 4 == 0) {
    // qweight [K, N/8] int32 -> box {16 words = one 128-column tile, 64 rows}; only used for L2 prefetch
    if (make_tmap_2d(a.qweight, 1, (uint64_t)(a.N / 8), (uint64_t)a.K, (uint64_t)(a.N / 8) * 4, 16, kBK, &tmq, false) ==
        cudaSuccess)
      p.has_tmq = 1;
  }
  switch (layout) {
    case 0: return a.G >= 64 ? dispatch_bt<0>(BT, tm, tmq, p, st) : dispatch_bt<3>(BT, tm, tmq, p, st);
    case 1: return dispatch_bt<1>(BT, tm, tmq, p, st);
    default: return dispatch_bt<2>(BT, tm, tmq, p, st);
  }
}

}  // namespace b200awq

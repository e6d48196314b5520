// CUDA-core GEMV path (M <= 8 tokens) for the three AWQ layouts.  HBM-bound: the job of these kernels
// is to stream the packed int4 weights once, with 128-bit coalesced loads and enough bytes in flight
// to cover DRAM latency, and to keep the ALU cost per weight below the issue budget that 6.5+ TB/s
// leaves (about 46 weights / clock / SM on a 148-SM B200).
//
// Arithmetic (all layouts): the 4-bit code q is used directly as an fp16 *subnormal* bit pattern
// (q * 2^-24, or q * 2^-20 for the nibbles that sit 4 bits higher), multiplied with the fp16 activation
// and accumulated in fp32 by FHFMA (fma.rn.f32.f16: exact product, one fp32 rounding) - one ALU op per
// weight plus 5/8 op of LOP3/SHF unpack.  Zero-point and scale are applied once per (group, column):
//     y[n] += s[g,n] * ( 2^24 * sum_k x[k] q[k,n]  -  z[g,n] * sum_k x[k] )
// The cross-thread / cross-CTA (split-K) reduction is fp32; the result is rounded to fp16 once.
#include "common.cuh"
#include "kernels.h"

namespace b200awq {

constexpr float kScaleA = 16777216.0f;  // 2^24: code read through mask 0x000f000f
constexpr float kScaleB = 1048576.0f;   // 2^20: code read through mask 0x00f000f0

// ======================================================================= GEMM layout [K, N/8]
// CTA = TX x TY threads.  Thread (tx, ty) owns WPT consecutive words (8*WPT columns) and RPT consecutive
// k-rows; a CTA covers TN = 8*WPT*TX columns x KC = RPT*TY rows.  grid = (ceil(N/TN), ceil(K/KC)).
template <int WPT>
struct WordVec;
template <>
struct WordVec<4> {
  using T = uint4;
  static __device__ __forceinline__ void load(const int32_t* p, uint32_t (&w)[4]) {
    uint4 v = ldg_stream_u4(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
};
template <>
struct WordVec<2> {
  static __device__ __forceinline__ void load(const int32_t* p, uint32_t (&w)[2]) {
    uint2 v = ldg_stream_u2(p);
    w[0] = v.x; w[1] = v.y;
  }
};
template <>
struct WordVec<1> {
  static __device__ __forceinline__ void load(const int32_t* p, uint32_t (&w)[1]) { w[0] = ldg_stream_u1(p); }
};

template <int WPT, int MT, int TX, int TY, int RPT>
__global__ void __launch_bounds__(TX* TY)
    gemv_gemm_layout_kernel(const __half* __restrict__ x, int64_t ldx, const int32_t* __restrict__ qweight,
                            const __half* __restrict__ scales, const int32_t* __restrict__ qzeros,
                            const __half* __restrict__ bias, __half* __restrict__ y, float* __restrict__ acc_ws,
                            int* __restrict__ tickets, int M, int K, int N, int G) {
  constexpr int NT = TX * TY;
  constexpr int CPT = 8 * WPT;        // columns per thread
  constexpr int TN = CPT * TX;        // columns per CTA
  constexpr int KC = RPT * TY;        // rows per CTA
  constexpr int RB = (RPT < 8) ? RPT : 8;  // rows per load batch
  static_assert(RPT % RB == 0, "RPT must be a multiple of the row batch");
  static_assert(RPT % 2 == 0, "x is read as half2");

  __shared__ __align__(16) __half xs[MT][KC];
  __shared__ float xsum[MT][TY];           // per (token, ty): sum of x over that thread-row's RPT rows
  __shared__ float red[TY][MT][TN + 4];    // +4 floats: de-phase the banks of consecutive ty
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int NW = N >> 3;                          // words per row
  const int k0 = blockIdx.y * KC;                 // first row of this CTA
  const int wc0 = blockIdx.x * (TN / 8) + tx * WPT;  // first word column of this thread
  const bool col_ok = wc0 < NW;                   // N % (8*WPT) == 0 is guaranteed by the launcher

  // ---- prologue: activations slice -> smem (zero padded past K / past M) ----------------------------
  for (int i = tid; i < MT * KC; i += NT) {
    const int m = i / KC, kk = i % KC;
    __half v = __float2half(0.f);
    if (m < M && k0 + kk < K) v = x[(int64_t)m * ldx + k0 + kk];
    xs[m][kk] = v;
  }
  __syncthreads();
  if (tid < MT * TY) {
    const int m = tid / TY, t = tid % TY;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RPT; ++i) s += __half2float(xs[m][t * RPT + i]);
    xsum[m][t] = s;
  }

  // ---- main loop: raw accumulation ------------------------------------------------------------------
  float acc[MT][CPT];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[m][c] = 0.f;

  const int r0 = k0 + ty * RPT;
  const int32_t* wp = qweight + (int64_t)r0 * NW + wc0;
#pragma unroll
  for (int b = 0; b < RPT / RB; ++b) {
    uint32_t q[RB][WPT];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = r0 + b * RB + i;
      if (col_ok && r < K) {
        WordVec<WPT>::load(wp + (int64_t)(b * RB + i) * NW, q[i]);
      } else {
#pragma unroll
        for (int w = 0; w < WPT; ++w) q[i][w] = 0u;
      }
    }
    // activations of these RB rows for every token: RB halves = RB/2 words each
    uint32_t xr[MT][RB / 2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < RB / 2; ++i)
        xr[m][i] = *reinterpret_cast<const uint32_t*>(&xs[m][ty * RPT + b * RB + 2 * i]);
#pragma unroll
    for (int i = 0; i < RB; ++i) {
#pragma unroll
      for (int w = 0; w < WPT; ++w) {
        const uint32_t word = q[i][w];
        const uint32_t w8 = word >> 8;
        const uint32_t p0 = word & 0x000f000fu, p1 = word & 0x00f000f0u;
        const uint32_t p2 = w8 & 0x000f000fu, p3 = w8 & 0x00f000f0u;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint16_t xh = (i & 1) ? hi16(xr[m][i / 2]) : lo16(xr[m][i / 2]);
          float* a = &acc[m][w * 8];
          a[0] = fhfma(lo16(p0), xh, a[0]);
          a[1] = fhfma(hi16(p0), xh, a[1]);
          a[2] = fhfma(lo16(p1), xh, a[2]);
          a[3] = fhfma(hi16(p1), xh, a[3]);
          a[4] = fhfma(lo16(p2), xh, a[4]);
          a[5] = fhfma(hi16(p2), xh, a[5]);
          a[6] = fhfma(lo16(p3), xh, a[6]);
          a[7] = fhfma(hi16(p3), xh, a[7]);
        }
      }
    }
  }

  // ---- cross-thread reduction over ty through smem --------------------------------------------------
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int c = 0; c < CPT; c += 4)
      *reinterpret_cast<float4*>(&red[ty][m][tx * CPT + c]) =
          make_float4(acc[m][c], acc[m][c + 1], acc[m][c + 2], acc[m][c + 3]);
  __syncthreads();

  // ---- fold scale / zero-point per (group, column), write or accumulate ----------------------------
  const int n_base = blockIdx.x * TN;
  const bool split = gridDim.y > 1;
  for (int o = tid; o < MT * TN; o += NT) {
    const int m = o / TN, c = o % TN;
    const int n = n_base + c;
    if (m >= M || n >= N) continue;
    const int j = c & 7;                       // column inside its word
    const float cs = ((j >> 1) & 1) ? kScaleB : kScaleA;
    const int zshift = 4 * ((j >> 1) + 4 * (j & 1));  // AWQ_REVERSE_ORDER[j] * 4
    float val = 0.f;
    // thread-rows are grouped by quantisation group: ty -> (k0 + ty*RPT) / G
    int t = 0;
    while (t < TY) {
      const int krow = k0 + t * RPT;
      if (krow >= K) break;
      const int g = krow / G;
      float raw = 0.f, xsg = 0.f;
      // consecutive thread-rows in the same group
      while (t < TY && (k0 + t * RPT) < K && (k0 + t * RPT) / G == g) {
        raw += red[t][m][c];
        xsg += xsum[m][t];
        ++t;
      }
      const float s = __half2float(scales[(int64_t)g * N + n]);
      const float z = static_cast<float>((static_cast<uint32_t>(qzeros[(int64_t)g * NW + (n >> 3)]) >> zshift) & 0xFu);
      val += s * (cs * raw - z * xsg);
    }
    if (!split) {
      if (bias != nullptr) val += __half2float(bias[n]);
      y[(int64_t)m * N + n] = __float2half_rn(val);
    } else {
      atomicAdd(&acc_ws[(int64_t)m * N + n], val);
    }
  }
  if (!split) return;

  // ---- split-K: the last CTA of this column block rounds, adds bias, and re-zeroes the scratch -------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&tickets[blockIdx.x], 1);
    s_last = (prev == static_cast<int>(gridDim.y) - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int o = tid; o < MT * TN; o += NT) {
    const int m = o / TN, c = o % TN;
    const int n = n_base + c;
    if (m >= M || n >= N) continue;
    float* p = &acc_ws[(int64_t)m * N + n];
    float val = ldcg_f1(p);
    *p = 0.f;
    if (bias != nullptr) val += __half2float(bias[n]);
    y[(int64_t)m * N + n] = __float2half_rn(val);
  }
  if (tid == 0) tickets[blockIdx.x] = 0;
}

template <int WPT, int MT, int TX, int TY, int RPT>
static cudaError_t launch_gemv_gemm_layout(const GemmArgs& a, float* acc_ws, int* tickets, cudaStream_t st) {
  constexpr int TN = 8 * WPT * TX, KC = RPT * TY;
  dim3 grid((a.N + TN - 1) / TN, (a.K + KC - 1) / KC);
  gemv_gemm_layout_kernel<WPT, MT, TX, TY, RPT><<<grid, TX * TY, 0, st>>>(
      reinterpret_cast<const __half*>(a.x), a.ldx, a.qweight, reinterpret_cast<const __half*>(a.scales), a.qzeros,
      reinterpret_cast<const __half*>(a.bias), reinterpret_cast<__half*>(a.y), acc_ws, tickets, a.M, a.K, a.N, a.G);
  return cudaGetLastError();
}

// Thread-rows must not straddle a quantisation group: RPT | G.  G is a multiple of 32 in every AWQ
// checkpoint (32 / 64 / 128 / K); the launcher falls back to RPT = 2 for exotic group sizes.
cudaError_t gemv_gemm_layout(const GemmArgs& a, float* acc_ws, int* tickets, cudaStream_t st) {
  const bool vec4 = (a.N % 32) == 0 && (reinterpret_cast<uintptr_t>(a.qweight) % 16) == 0;
  const bool rpt16 = (a.G % 16) == 0;
  if (!rpt16) {
    // generic slow-but-correct shape: one word per thread, 2 rows per thread
    if (a.M <= 1) return launch_gemv_gemm_layout<1, 1, 32, 4, 2>(a, acc_ws, tickets, st);
    if (a.M <= 2) return launch_gemv_gemm_layout<1, 2, 32, 4, 2>(a, acc_ws, tickets, st);
    if (a.M <= 4) return launch_gemv_gemm_layout<1, 4, 32, 4, 2>(a, acc_ws, tickets, st);
    return launch_gemv_gemm_layout<1, 8, 32, 4, 2>(a, acc_ws, tickets, st);
  }
  if (a.M <= 1) {
    if (vec4) return launch_gemv_gemm_layout<4, 1, 16, 8, 16>(a, acc_ws, tickets, st);
    return launch_gemv_gemm_layout<1, 1, 32, 4, 16>(a, acc_ws, tickets, st);
  }
  const bool vec2 = (a.N % 16) == 0 && (reinterpret_cast<uintptr_t>(a.qweight) % 8) == 0;
  if (a.M <= 2) {
    if (vec2) return launch_gemv_gemm_layout<2, 2, 16, 8, 16>(a, acc_ws, tickets, st);
    return launch_gemv_gemm_layout<1, 2, 32, 4, 16>(a, acc_ws, tickets, st);
  }
  if (a.M <= 4) return launch_gemv_gemm_layout<1, 4, 32, 4, 16>(a, acc_ws, tickets, st);
  return launch_gemv_gemm_layout<1, 8, 32, 4, 16>(a, acc_ws, tickets, st);
}

}  // namespace b200awq

// ======================================================================= GEMV layout [N, K/8]
// Output row n is a contiguous K/2-byte run: a warp streams it 512 B per instruction (lane = 32
// consecutive k), keeps two fp32 accumulators per (row, token) (the two nibble scale classes), folds
// scale / zero once per 32-k chunk (always inside one quantisation group since G % 32 == 0) and finishes
// with a 5-step warp-shuffle reduction.  No split-K, deterministic.  Activations live in shared memory,
// permuted so that the four 16-byte pieces a lane needs for one chunk are each a conflict-free
// 512-byte warp access:  piece c of chunk (it, lane)  ->  slot (it * 4 + c) * 32 + lane.
namespace b200awq {

template <int MT, int RW>
__global__ void __launch_bounds__(256)
    gemv_gemv_layout_kernel(const __half* __restrict__ x, int64_t ldx, const int32_t* __restrict__ qweight,
                            const __half* __restrict__ scales, const int32_t* __restrict__ qzeros,
                            const __half* __restrict__ bias, __half* __restrict__ y, int M, int K, int N, int G,
                            int zw) {
  extern __shared__ __align__(16) uint4 xs_perm[];  // [MT][nit*128] uint4
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nit = (K + 1023) / 1024;                 // 1024-k passes per row
  const int slots = nit * 128;                       // uint4 slots per token
  // prologue: stage x (8 halves per slot)
  for (int i = tid; i < MT * slots; i += blockDim.x) {
    const int m = i / slots, s = i % slots;
    const int it = s / 128, c = (s % 128) / 32, l = s % 32;
    const int k = it * 1024 + l * 32 + c * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m < M && k < K) {
      const __half* src = x + (int64_t)m * ldx + k;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        v = *reinterpret_cast<const uint4*>(src);
      } else {
        __half t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = src[j];
        v = *reinterpret_cast<uint4*>(t);
      }
    }
    xs_perm[i] = v;
  }
  __syncthreads();

  const int KW = K >> 3;
  const int n0 = (blockIdx.x * 8 + warp) * RW;
  float val[RW][MT];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) val[r][m] = 0.f;

  for (int it = 0; it < nit; ++it) {
    const int kc = it * 1024 + lane * 32;  // first k of this lane's chunk
    const bool k_ok = kc < K;
    uint4 q[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      q[r] = make_uint4(0, 0, 0, 0);
      if (k_ok && n0 + r < N) q[r] = ldg_stream_u4(qweight + (int64_t)(n0 + r) * KW + (kc >> 3));
    }
    const int g = k_ok ? kc / G : 0;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      // 32 activations of this chunk
      uint32_t xr[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 v = xs_perm[m * slots + (it * 4 + c) * 32 + lane];
        xr[4 * c + 0] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
      }
      float xsum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 f = __half22float2(u32_as_h2(xr[i]));
        xsum += f.x + f.y;
      }
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        float aA = 0.f, aB = 0.f;
        const uint32_t w[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // word j: k = kc + 8j + i at nibble i; x pairs xr[4j + i/2]
          const uint32_t w8 = w[j] >> 8;
          const uint32_t p0 = w[j] & 0x000f000fu;  // (k0, k4)  class A
          const uint32_t p1 = w[j] & 0x00f000f0u;  // (k1, k5)  class B
          const uint32_t p2 = w8 & 0x000f000fu;    // (k2, k6)  class A
          const uint32_t p3 = w8 & 0x00f000f0u;    // (k3, k7)  class B
          const uint32_t x01 = xr[4 * j + 0], x23 = xr[4 * j + 1], x45 = xr[4 * j + 2], x67 = xr[4 * j + 3];
          aA = fhfma(lo16(p0), lo16(x01), aA);
          aA = fhfma(hi16(p0), lo16(x45), aA);
          aB = fhfma(lo16(p1), hi16(x01), aB);
          aB = fhfma(hi16(p1), hi16(x45), aB);
          aA = fhfma(lo16(p2), lo16(x23), aA);
          aA = fhfma(hi16(p2), lo16(x67), aA);
          aB = fhfma(lo16(p3), hi16(x23), aB);
          aB = fhfma(hi16(p3), hi16(x67), aB);
        }
        if (k_ok && n0 + r < N) {
          const float s = __half2float(scales[(int64_t)(n0 + r) * (zw * 8) + g]);
          const uint32_t zword = static_cast<uint32_t>(qzeros[(int64_t)(n0 + r) * zw + (g >> 3)]);
          const float z = static_cast<float>((zword >> (4 * (g & 7))) & 0xFu);
          val[r][m] += s * (kScaleA * aA + kScaleB * aB - z * xsum);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float v = val[r][m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && n0 + r < N && m < M) {
        if (bias != nullptr) v += __half2float(bias[n0 + r]);
        y[(int64_t)m * N + n0 + r] = __float2half_rn(v);
      }
    }
}

template <int MT, int RW>
static cudaError_t launch_gemv_gemv_layout(const GemmArgs& a, int zw, cudaStream_t st) {
  const int nit = (a.K + 1023) / 1024;
  const size_t smem = (size_t)MT * nit * 128 * sizeof(uint4);
  auto kern = gemv_gemv_layout_kernel<MT, RW>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  const int rows_per_cta = 8 * RW;
  kern<<<(a.N + rows_per_cta - 1) / rows_per_cta, 256, smem, st>>>(
      reinterpret_cast<const __half*>(a.x), a.ldx, a.qweight, reinterpret_cast<const __half*>(a.scales), a.qzeros,
      reinterpret_cast<const __half*>(a.bias), reinterpret_cast<__half*>(a.y), a.M, a.K, a.N, a.G, zw);
  return cudaGetLastError();
}

static int zeros_width(int K, int G) {  // awq/modules/linear/gemv.py:12-24
  const int mult = G >= 128 ? 1 : (G == 64 ? 2 : 4);
  int base = ((K / G) + 7) / 8;
  return ((base + mult - 1) / mult) * mult;
}

cudaError_t gemv_gemv_layout(const GemmArgs& a0, cudaStream_t st) {
  const int zw = zeros_width(a0.K, a0.G);
  // keep the staged activations under ~200 KB of shared memory: at most `cap` tokens per pass
  const int nit = (a0.K + 1023) / 1024;
  const size_t per_tok = (size_t)nit * 128 * 16;
  int done = 0;
  while (done < a0.M) {
    GemmArgs a = a0;
    int m = a0.M - done;
    int mt = m <= 1 ? 1 : (m <= 2 ? 2 : (m <= 4 ? 4 : 8));
    while (mt > 1 && per_tok * mt > 200 * 1024) mt >>= 1;
    if (m > mt) m = mt;
    a.M = m;
    a.x = reinterpret_cast<const __half*>(a0.x) + (int64_t)done * a0.ldx;
    a.y = reinterpret_cast<__half*>(a0.y) + (int64_t)done * a0.N;
    cudaError_t e;
    if (mt == 1) e = launch_gemv_gemv_layout<1, 2>(a, zw, st);
    else if (mt == 2) e = launch_gemv_gemv_layout<2, 2>(a, zw, st);
    else if (mt == 4) e = launch_gemv_gemv_layout<4, 1>(a, zw, st);
    else e = launch_gemv_gemv_layout<8, 1>(a, zw, st);
    if (e != cudaSuccess) return e;
    done += m;
  }
  return cudaSuccess;
}

}  // namespace b200awq

// ======================================================================= GEMVFast layout
// qweight int16 [N/4, K] (awq/modules/linear/gemv_fast.py:26-65).  Decoded structure: for the 4-row
// group R and the 64-k block b, the 64 int16 at [R, 64b .. 64b+63] are 4 runs of 16 (one per row
// r = 0..3); in a run, int16 i (i = 0..15) nibble r' holds k = 64b + 32(i/8) + (i%8) + 8r'.  Hence a 16-byte
// piece is 32 consecutive-in-k weights of ONE row, and (word >> 4r') & 0x000f000f is the natural pair
// (k, k+1) with k = base + 2u + 8r' - the layout was built for exactly this unpack.
// scales / scaled zeros are [groups(padded), N] fp16 with W = q*s + sz, sz = -(z*s) rounded to fp16.
// A warp streams one row group: 512 B per instruction = 4 blocks x 4 rows; lane l -> block l/8, row
// (l%8)/2, half l%2.  Reduction over the 8 lanes of a row by shuffles (xor 1, 8, 16).
namespace b200awq {

template <int MT>
__global__ void __launch_bounds__(128)
    gemv_fast_layout_kernel(const __half* __restrict__ x, int64_t ldx, const int16_t* __restrict__ qweight,
                            const __half* __restrict__ scales, const __half* __restrict__ szeros,
                            const __half* __restrict__ bias, __half* __restrict__ y, int M, int K, int N, int G) {
  extern __shared__ __align__(16) uint4 xs4[];  // [MT][K/8] uint4, natural order
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KV = K >> 3;
  for (int i = tid; i < MT * KV; i += blockDim.x) {
    const int m = i / KV, k = (i % KV) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m < M) {
      const __half* src = x + (int64_t)m * ldx + k;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        v = *reinterpret_cast<const uint4*>(src);
      } else {
        __half t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = src[j];
        v = *reinterpret_cast<uint4*>(t);
      }
    }
    xs4[i] = v;
  }
  __syncthreads();

  const int R = blockIdx.x * 4 + warp;           // row group
  if (R * 4 >= N) return;
  const int bo = lane >> 3, r = (lane & 7) >> 1, h = lane & 1;
  const int n = R * 4 + r;
  const int16_t* wrow = qweight + (int64_t)R * K;  // K int16 per row group... (N/4 rows of K int16)
  float val[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) val[m] = 0.f;
  const int nit = (K + 255) / 256;
  constexpr int UN = 4;
  for (int it0 = 0; it0 < nit; it0 += UN) {
    uint4 q[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int it = it0 + u;
      const int kc = it * 256 + bo * 64 + h * 32;
      q[u] = make_uint4(0, 0, 0, 0);
      // int16 offset inside the row group: block (it*4+bo)*64 + run r*16 + half h*8
      if (it < nit && kc < K) q[u] = ldg_stream_u4(wrow + (int64_t)(it * 4 + bo) * 64 + r * 16 + h * 8);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int it = it0 + u;
      const int kc = it * 256 + bo * 64 + h * 32;
      if (it >= nit || kc >= K) continue;
      const int g = kc / G;
      const float s = __half2float(scales[(int64_t)g * N + n]);
      const float sz = __half2float(szeros[(int64_t)g * N + n]);
      const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        uint32_t xr[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v = xs4[m * KV + (kc >> 3) + c];
          xr[4 * c + 0] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
        }
        float xsum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 f = __half22float2(u32_as_h2(xr[i]));
          xsum += f.x + f.y;
        }
        float aA = 0.f, aB = 0.f;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const uint32_t w8 = w[uu] >> 8;
          const uint32_t p0 = w[uu] & 0x000f000fu, p1 = w[uu] & 0x00f000f0u;
          const uint32_t p2 = w8 & 0x000f000fu, p3 = w8 & 0x00f000f0u;
          aA = fhfma(lo16(p0), lo16(xr[uu]), aA);
          aA = fhfma(hi16(p0), hi16(xr[uu]), aA);
          aB = fhfma(lo16(p1), lo16(xr[uu + 4]), aB);
          aB = fhfma(hi16(p1), hi16(xr[uu + 4]), aB);
          aA = fhfma(lo16(p2), lo16(xr[uu + 8]), aA);
          aA = fhfma(hi16(p2), hi16(xr[uu + 8]), aA);
          aB = fhfma(lo16(p3), lo16(xr[uu + 12]), aB);
          aB = fhfma(hi16(p3), hi16(xr[uu + 12]), aB);
        }
        val[m] += s * (kScaleA * aA + kScaleB * aB) + sz * xsum;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    float v = val[m];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if ((lane & 0x19) == 0 && m < M && n < N) {  // lanes 0, 2, 4, 6: one per row
      if (bias != nullptr) v += __half2float(bias[n]);
      y[(int64_t)m * N + n] = __float2half_rn(v);
    }
  }
}

template <int MT>
static cudaError_t launch_gemv_fast(const FastArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)MT * (a.K / 8) * sizeof(uint4);
  auto kern = gemv_fast_layout_kernel<MT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  const int groups = a.N / 4;
  kern<<<(groups + 3) / 4, 128, smem, st>>>(reinterpret_cast<const __half*>(a.x), a.ldx, a.qweight,
                                            reinterpret_cast<const __half*>(a.scales),
                                            reinterpret_cast<const __half*>(a.szeros),
                                            reinterpret_cast<const __half*>(a.bias), reinterpret_cast<__half*>(a.y),
                                            a.M, a.K, a.N, a.G);
  return cudaGetLastError();
}

cudaError_t gemv_fast_layout(const FastArgs& a0, cudaStream_t st) {
  const size_t per_tok = (size_t)(a0.K / 8) * 16;
  int done = 0;
  while (done < a0.M) {
    FastArgs a = a0;
    int m = a0.M - done;
    int mt = m <= 1 ? 1 : (m <= 2 ? 2 : (m <= 4 ? 4 : 8));
    while (mt > 1 && per_tok * mt > 200 * 1024) mt >>= 1;
    if (m > mt) m = mt;
    a.M = m;
    a.x = reinterpret_cast<const __half*>(a0.x) + (int64_t)done * a0.ldx;
    a.y = reinterpret_cast<__half*>(a0.y) + (int64_t)done * a0.N;
    cudaError_t e;
    if (mt == 1) e = launch_gemv_fast<1>(a, st);
    else if (mt == 2) e = launch_gemv_fast<2>(a, st);
    else if (mt == 4) e = launch_gemv_fast<4>(a, st);
    else e = launch_gemv_fast<8>(a, st);
    if (e != cudaSuccess) return e;
    done += m;
  }
  return cudaSuccess;
}

}  // namespace b200awq

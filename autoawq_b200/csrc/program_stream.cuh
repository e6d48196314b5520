// Decode program, stream variant: the persistent kernel of program.cu re-built around a ONE-TIME RE-LAYOUT of the
// packed weights (the "stream format", restated in numpy in oracle/stream_format.py; the reference's precedent for
// a post-load re-layout is awq/modules/linear/exllama.py:66-79) so that the grid-wide hand-off between two
// dependent linears shrinks to "store, poll".
//
// Why (round-1 measurements, profiles/r01_program_phase_timeline.log): with the checkpoint's GEMM layout
// [K, N/8] a DRAM-efficient tile is >= 128 bytes = 256 columns wide, so N = 4096 has only 16 column blocks for 148
// CTAs: split-K with a fan-in of ~9 CTAs per column block is forced, and its cost - ~110 k 64-bit REDs per op into
// L2, every CTA polling 64-bit sums, a reclamation protocol for the accumulator rows - was ~6 us per op boundary
// against ~4 us of weight streaming.  In the stream format ANY partition is contiguous in memory, so the work is
// cut OUTPUT-STATIONARY: a CTA owns whole 16-column sets (all of K), its 8 consumer warps split the CTA's units
// (set, 128 rows of K) evenly, partial sums meet in shared memory, and the CTA publishes FINISHED fp16 outputs:
//   * no cross-CTA reduction, no atomics, no fixed-point packing, nothing to zero or reclaim, no duty warp;
//   * the hand-off word is (fp16 value | 16-bit tag): one plain 32-bit store by the owner, polled by the consumers
//     with ld.relaxed.gpu - a word is valid when its tag equals the tag of (run, op), so buffers never need
//     clearing and a run is bit-reproducible (fixed summation order);
//   * every CTA reads the whole activation row (K x 4 bytes from L2), RMSNorm needs no second grid-wide pass;
//   * SiLU*mul is fused into the PRODUCER: the stream format of a gate|up linear pairs gate column j and up
//     column j in one lane, so the consumer of `down` polls d words instead of 2 d.
// The weight stream itself is as before: a producer warp keeps a shared-memory ring of bulk copies
// (cp.async.bulk, plain 1-D: every warp's byte range is contiguous) full ACROSS op boundaries.
//
// Included by program.cu (one translation unit: shares the watchdog / debug symbols).
#pragma once

namespace b200awq {

constexpr int kSpWarps = 8;                       // consumer warps
constexpr int kSpThreads = 32 + kSpWarps * 32;    // producer warp + consumers
constexpr int kSpCons = kSpWarps * 32;
constexpr int kSpStageBytes = 8576;               // 8 units of G >= 128 (8 x 1072), 15 of G = 64, 28 of G = 32
constexpr int kSpSPW = 2;                         // ring stages per consumer warp
constexpr int kSpStages = kSpWarps * kSpSPW;
constexpr int kSpAux = 48;                        // group constants per unit: 32 B scales + 8 B zeros + 8 B pad
constexpr int kSpLMax = 32;                       // 16-column sets one CTA may touch in one op
constexpr int kSpRows = 4;                        // hand-off rows in rotation (op i publishes into row i % 4)
constexpr int kSpXsumMax = 1024;                  // units along K (K / UK) an op may have

struct __align__(128) SpOp {
  const uint8_t* wstream;      // stream-format weights
  const uint32_t* cta_begin;   // [grid + 1] first unit of every CTA (unit = set * NU + j)
  const __half* bias;
  __half* y;                   // fp16 output of the per-op path (every buffer holds the same values after a run)
  const __half* src;           // source in plain global memory (src_op < 0)
  const __half* norm_w;        // RMSNorm weight [K]
  __half* xout;                // RMSNorm prologue: where the recorded norm wanted its output, or null
  __half* act_out;             // mode 1: where the recorded SiLU*mul wanted its output, or null
  int K, N;
  int uk_shift, F, NU, unit_bytes, ups;
  int mode;                    // 0: plain sets, 1: gate|up pairs (publishes silu(gate) * up, N / 2 columns)
  int prologue;                // kProCopy / kProRmsnorm
  int src_op, src_off;         // >= 0: the source is op src_op's published row, from column src_off
  float eps;
  int pad_[4];
};
static_assert(sizeof(SpOp) == 128, "SpOp layout");

__host__ __device__ constexpr size_t sp_fixed_smem() {
  return (size_t)kSpStages * kSpStageBytes + (size_t)kSpLMax * kSpWarps * 16 * 4 + (size_t)kSpXsumMax * 4 +
         2 * kSpStages * 8 + 256;
}
static_assert(sp_fixed_smem() % 16 == 0, "xs must stay 16-byte aligned");

__device__ __forceinline__ uint4 ld_relaxed_u4(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_relaxed_u32(void* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t sp_tag(int base, int op) { return (uint32_t)((base + op) % 65535 + 1); }

// set -> original columns (oracle/stream_format.py:set_columns)
__device__ __forceinline__ void sp_cols(int mode, int N, int s, int g, int& lo, int& hi) {
  if (mode == 0) {
    lo = 16 * s + g;
    hi = lo + 8;
  } else {
    lo = 8 * s + g;
    hi = (N >> 1) + lo;
  }
}

// ------------------------------------------------------------------------------------------ re-layout kernel
// One thread per output word / per group-constant slot; run once per linear at program creation.
__global__ void __launch_bounds__(256)
    stream_pack_kernel(const int32_t* __restrict__ qweight, const __half* __restrict__ scales,
                       const int32_t* __restrict__ qzeros, uint8_t* __restrict__ out, int K, int N, int G, int mode) {
  const int UK = G < 128 ? G : 128, F = UK >> 4, NU = K / UK, UB = F * 128 + kSpAux;
  const int NW = N >> 3;
  const int wpu = F * 32 + 12;   // 32-bit slots per unit: fragment words + 8 scale pairs + 2 zero words + 2 pad
  const int64_t total = (int64_t)(N >> 4) * NU * wpu;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int slot = (int)(i % wpu);
    const int64_t unit = i / wpu;
    const int j = (int)(unit % NU), s = (int)(unit / NU);
    uint8_t* ub = out + unit * UB;
    auto nib = [&](int k, int col) -> uint32_t {
      const uint32_t w = (uint32_t)qweight[(int64_t)k * NW + (col >> 3)];
      const int jj = col & 7;
      return (w >> (4 * ((jj >> 1) + 4 * (jj & 1)))) & 0xFu;   // 4 * AWQ_REVERSE_ORDER[jj]
    };
    if (slot < F * 32) {
      int f, lane;
      if (F >= 4) {          // [quad][lane][4]
        const int quad = slot >> 7, r = slot & 127;
        lane = r >> 2;
        f = quad * 4 + (r & 3);
      } else {               // [lane][2]
        lane = slot >> 1;
        f = slot & 1;
      }
      const int g = lane >> 2, tig = lane & 3;
      int lo, hi;
      sp_cols(mode, N, s, g, lo, hi);
      const int k0 = j * UK + 16 * f + 2 * tig;
      const uint32_t w = nib(k0, lo) | nib(k0, hi) << 4 | nib(k0 + 8, lo) << 8 | nib(k0 + 8, hi) << 12 |
                         nib(k0 + 1, lo) << 16 | nib(k0 + 1, hi) << 20 | nib(k0 + 9, lo) << 24 | nib(k0 + 9, hi) << 28;
      reinterpret_cast<uint32_t*>(ub)[slot] = w;
    } else {
      const int a = slot - F * 32;      // 0..7 scale pairs, 8..9 zero words, 10..11 pad
      const int grp = (j * UK) / G;
      uint32_t v = 0;
      if (a < 8) {
        int lo, hi;
        sp_cols(mode, N, s, a, lo, hi);
        const __half sl = scales[(int64_t)grp * N + lo], sh = scales[(int64_t)grp * N + hi];
        v = (uint32_t)__half_as_ushort(sl) | (uint32_t)__half_as_ushort(sh) << 16;
      } else if (a < 10) {
        for (int b = 0; b < 4; ++b) {
          const int g = (a - 8) * 4 + b;
          int lo, hi;
          sp_cols(mode, N, s, g, lo, hi);
          auto znib = [&](int col) -> uint32_t {
            const uint32_t w = (uint32_t)qzeros[(int64_t)grp * NW + (col >> 3)];
            const int jj = col & 7;
            return (w >> (4 * ((jj >> 1) + 4 * (jj & 1)))) & 0xFu;
          };
          v |= (znib(lo) | znib(hi) << 4) << (8 * b);
        }
      }
      reinterpret_cast<uint32_t*>(ub + F * 128)[a] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------ the kernel
// F fragments of one unit (one 16-column set, UK rows of K) times the activations: acc += raw codes . x
template <int F>
__device__ __forceinline__ void sp_unit_mma(const uint8_t* __restrict__ up, const uint32_t* __restrict__ xsu, int lane,
                                            bool xl, float (&acc)[4]) {
  constexpr uint32_t MA = 0x000f000fu, MB = 0x00f000f0u, MG = 0x64006400u;
  const int tig = lane & 3;
  if constexpr (F >= 4) {
#pragma unroll
    for (int qd = 0; qd < F / 4; ++qd) {
      const uint4 q = *reinterpret_cast<const uint4*>(up + qd * 512 + lane * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w = (&q.x)[j], w8 = w >> 8;
        uint2 xb = make_uint2(0u, 0u);
        if (xl) xb = *reinterpret_cast<const uint2*>(xsu + ((qd * 4 + j) * 4 + tig) * 2);
        mma_16816(acc, lop3_and_or(w, MA, MG), lop3_and_or(w, MB, MG), lop3_and_or(w8, MA, MG), lop3_and_or(w8, MB, MG),
                  xb.x, xb.y);
      }
    }
  } else {
    const uint2 q = *reinterpret_cast<const uint2*>(up + lane * 8);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t w = (&q.x)[j], w8 = w >> 8;
      uint2 xb = make_uint2(0u, 0u);
      if (xl) xb = *reinterpret_cast<const uint2*>(xsu + (j * 4 + tig) * 2);
      mma_16816(acc, lop3_and_or(w, MA, MG), lop3_and_or(w, MB, MG), lop3_and_or(w8, MA, MG), lop3_and_or(w8, MB, MG),
                xb.x, xb.y);
    }
  }
}

// debug stamps (knob 3 = 2), per op and CTA (first 8 CTAs, first 32 ops):
// [0] op begin, [1] source row complete (poll over), [2] activations staged, [3] warp 0's first chunk landed,
// [4] warp 0 finished its units, [5] all warps finished, [6] outputs published, [7] unused
#define SP_STAMP(slot)                                                                                       \
  do {                                                                                                       \
    if (dbg == 2 && ct == 0 && blockIdx.x < 8 && op < 32) g_prog_dbg[(op * 8 + blockIdx.x) * 8 + (slot)] = prog_timer(); \
  } while (0)

__global__ void __launch_bounds__(kSpThreads, 1)
    stream_program_kernel(const SpOp* __restrict__ ops, int n_ops, uint32_t* __restrict__ rows, int row_stride,
                          int* __restrict__ state, int dbg) {
  extern __shared__ __align__(1024) uint8_t sp_smem[];
  uint8_t* ring = sp_smem;
  float* part = reinterpret_cast<float*>(sp_smem + (size_t)kSpStages * kSpStageBytes);   // [LMax][8 warps][16]
  float* xsum = part + kSpLMax * kSpWarps * 16;                                           // [K / UK]
  uint64_t* full = reinterpret_cast<uint64_t*>(xsum + kSpXsumMax);
  uint64_t* empty = full + kSpStages;
  int* misc = reinterpret_cast<int*>(empty + kSpStages);
  float* wsum = reinterpret_cast<float*>(misc);        // [8]
  int* wfirst = misc + 8;                              // [8] first local set each warp touched (-1: none)
  int* wlast = misc + 16;                              // [8]
  uint32_t* xs = reinterpret_cast<uint32_t*>(sp_smem + sp_fixed_smem());   // activations in B-fragment order

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = gridDim.x, bid = blockIdx.x;
  const int base = state[0];     // tag base of this run (advanced by the last CTA to leave, see the end)

  if (tid == 0) {
    for (int s = 0; s < kSpStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ============================================================ producer: the weight stream of ALL ops
    if (lane < kSpWarps) {
      const int w = lane;
      int stage_i = 0;
      uint32_t ph = 0;
      bool dead = false;      // a wait was abandoned (watchdog): stop feeding, fall through to the common exit
      for (int op = 0; op < n_ops && !dead; ++op) {
        const SpOp* o = ops + op;
        const uint32_t u0 = o->cta_begin[bid], u1 = o->cta_begin[bid + 1];
        const uint32_t nu = u1 - u0;
        const uint32_t ua = u0 + (uint32_t)((uint64_t)nu * w / kSpWarps), ub = u0 + (uint32_t)((uint64_t)nu * (w + 1) / kSpWarps);
        const int UB = o->unit_bytes, ups = o->ups;
        const uint8_t* src = o->wstream;
        for (uint32_t u = ua; u < ub; u += ups) {
          const int n = (int)(ub - u) < ups ? (int)(ub - u) : ups;
          const int stage = w * kSpSPW + stage_i;
          if (!prog_mbar_wait(&empty[stage], ph ^ 1, kWEmpty, op)) {
            dead = true;
            break;
          }
          mbar_arrive_expect_tx(&full[stage], (uint32_t)(n * UB));
          bulk_load_1d(ring + (size_t)stage * kSpStageBytes, src + (size_t)u * UB, (uint32_t)(n * UB), &full[stage]);
          if (++stage_i == kSpSPW) { stage_i = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ================================================================ consumers
    const int cw = warp - 1;
    const int ct = tid - 32;
    const int g = lane >> 2, tig = lane & 3;
    const bool xl = g == 0;       // M = 1: token 0 is column n = 0 of the MMA's B operand, supplied by the g = 0 lanes
    int stage_i = 0;
    uint32_t ph = 0;

    for (int op = 0; op < n_ops; ++op) {
      const SpOp* o = ops + op;
      const int K = o->K, N = o->N, NU = o->NU, F = o->F, UB = o->unit_bytes, ups = o->ups, mode = o->mode;
      const uint32_t u0 = o->cta_begin[bid], u1 = o->cta_begin[bid + 1];
      const uint32_t nu = u1 - u0;
      const uint32_t ua = u0 + (uint32_t)((uint64_t)nu * cw / kSpWarps), ub = u0 + (uint32_t)((uint64_t)nu * (cw + 1) / kSpWarps);
      const int set0 = (int)(u0 / NU);                            // first set of the CTA
      SP_STAMP(0);

      // ---- stage the activations (whole row, every CTA): poll the producer's published row / read global memory,
      //      apply the recorded RMSNorm, write them in B-fragment order and keep the per-unit sums sum_k x_k.
      //      Thread t takes 8 consecutive k per pass (k = 2048 pass + 8 t) and sums squares in the order of
      //      aux.cu's rmsnorm_kernel, so the norm reproduces the stand-alone kernel bit for bit.
      {
        const int uk_shift = o->uk_shift;
        const int seg = (1 << uk_shift) >> 3;                      // lanes per unit (each lane holds 8 consecutive k)
        const bool from_row = o->src_op >= 0;
        const uint32_t* row = from_row ? rows + (size_t)(o->src_op % kSpRows) * row_stride + o->src_off : nullptr;
        const uint32_t want = from_row ? sp_tag(base, o->src_op) : 0u;
        const __half* src = o->src;
        const bool norm = o->prologue == kProRmsnorm;
        float ss = 0.f;
        // fragment order: k = 16 ks + j -> word ks * 8 + ((j & 7) >> 1) * 2 + (j >> 3)   (a word = the pair (j, j + 1))
        auto frag_ptr = [&](int c) { return xs + (c >> 4) * 8 + ((c >> 3) & 1); };   // + 2 * pair index
        auto unit_sums = [&](int c, bool ok, const uint32_t (&h)[4]) {
          float sx = 0.f;
          if (ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(u32_as_h2(h[q]));
              sx += f.x + f.y;
            }
          }
          for (int d = 1; d < seg; d <<= 1) sx += __shfl_xor_sync(0xffffffffu, sx, d);
          if (ok && (lane & (seg - 1)) == 0) xsum[c >> uk_shift] = sx;
        };
        for (int cb = cw * 256; cb < K; cb += kSpCons * 8) {      // warp-uniform trip count
          const int c = cb + lane * 8;
          const bool ok = c < K;
          uint32_t h[4] = {0u, 0u, 0u, 0u};                        // the eight fp16 values as four pairs
          if (ok) {
            if (from_row) {
              uint4 v0, v1;
              ProgWatch wd;
              for (;;) {
                v0 = ld_relaxed_u4(row + c);
                v1 = ld_relaxed_u4(row + c + 4);
                if ((v0.x >> 16) == want && (v0.y >> 16) == want && (v0.z >> 16) == want && (v0.w >> 16) == want &&
                    (v1.x >> 16) == want && (v1.y >> 16) == want && (v1.z >> 16) == want && (v1.w >> 16) == want)
                  break;
                if (wd.tick(kWCopy, op)) break;
              }
              h[0] = (v0.x & 0xffffu) | (v0.y << 16);
              h[1] = (v0.z & 0xffffu) | (v0.w << 16);
              h[2] = (v1.x & 0xffffu) | (v1.y << 16);
              h[3] = (v1.z & 0xffffu) | (v1.w << 16);
            } else {
              const uint4 v = ldg_stream_u4(src + c);
              h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
            }
            uint32_t* dst = frag_ptr(c);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dst[2 * q] = h[q];
              const float2 f = __half22float2(u32_as_h2(h[q]));
              ss += f.x * f.x + f.y * f.y;
            }
          }
          if (!norm) unit_sums(c, ok, h);
        }
        SP_STAMP(1);
        if (norm) {
          ss = prog_warp_sum(ss);
          if (lane == 0) wsum[cw] = ss;
          named_bar_sync_gv(1, kSpCons);
          float tot = 0.f;
#pragma unroll
          for (int i = 0; i < kSpWarps; ++i) tot += wsum[i];
          const float rs = rsqrtf(tot / static_cast<float>(K) + o->eps);
          const __half* nw = o->norm_w;
          __half* xout = o->xout;
          int xlo = 0, xhi = 0;
          if (xout != nullptr) {       // this CTA's share of the norm's recorded output buffer
            const int u8 = K >> 3;
            xlo = (int)((int64_t)u8 * bid / nblk) << 3;
            xhi = (int)((int64_t)u8 * (bid + 1) / nblk) << 3;
          }
          for (int cb = cw * 256; cb < K; cb += kSpCons * 8) {    // the thread's own chunks again
            const int c = cb + lane * 8;
            const bool ok = c < K;
            uint32_t h[4] = {0u, 0u, 0u, 0u};
            if (ok) {
              uint32_t* dst = frag_ptr(c);
              const uint4 wv = __ldg(reinterpret_cast<const uint4*>(nw + c));
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 a = __half22float2(u32_as_h2(dst[2 * q]));
                const float2 wq = __half22float2(u32_as_h2((&wv.x)[q]));
                // arithmetic of aux.cu's rmsnorm_kernel: fp16(x * rs * w)
                h[q] = h2_as_u32(__halves2half2(__float2half_rn(a.x * rs * wq.x), __float2half_rn(a.y * rs * wq.y)));
                dst[2 * q] = h[q];
              }
              if (c >= xlo && c < xhi) *reinterpret_cast<uint4*>(xout + c) = make_uint4(h[0], h[1], h[2], h[3]);
            }
            unit_sums(c, ok, h);
          }
        }
        if (lane == 0) {
          wfirst[cw] = -1;
          wlast[cw] = -1;
        }
        named_bar_sync_gv(1, kSpCons);
      }
      SP_STAMP(2);

      // ---- this warp's run of units
      {
        int s_cur = (int)(ua / NU), j = (int)(ua - (uint32_t)s_cur * NU);
        float ylo = 0.f, yhi = 0.f;
        int first_ls = -1, last_ls = -1;
        auto flush = [&]() {
          const int ls = s_cur - set0;
          if (tig == 0) {
            float* p = part + ((size_t)ls * kSpWarps + cw) * 16;
            p[g] = ylo;
            p[g + 8] = yhi;
          }
          if (first_ls < 0) first_ls = ls;
          last_ls = ls;
          ylo = yhi = 0.f;
        };
        for (uint32_t u = ua; u < ub; u += ups) {
          const int n = (int)(ub - u) < ups ? (int)(ub - u) : ups;
          const int stage = cw * kSpSPW + stage_i;
          prog_mbar_wait(&full[stage], ph, kWFull, op);
          if (u == ua) SP_STAMP(3);
          const uint8_t* st = ring + (size_t)stage * kSpStageBytes;
          for (int i = 0; i < n; ++i) {
            const uint8_t* up = st + (size_t)i * UB;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t* xsu = xs + (size_t)j * F * 8;
            if (F == 8) sp_unit_mma<8>(up, xsu, lane, xl, acc);
            else if (F == 4) sp_unit_mma<4>(up, xsu, lane, xl, acc);
            else sp_unit_mma<2>(up, xsu, lane, xl, acc);
            // fold with the unit's group constants: y += s * (S - (1024 + c z) * X) / c   (csrc/gemv_tile.cuh:v3_fold)
            const uint8_t* ax = up + F * 128;
            const float2 sc = __half22float2(u32_as_h2(*reinterpret_cast<const uint32_t*>(ax + 4 * g)));
            const uint32_t zb = ax[32 + g];
            const float X = xsum[j];
            ylo += sc.x * (acc[0] - (1024.f + static_cast<float>(zb & 0xFu)) * X);
            yhi += (sc.y * 0.0625f) * (acc[2] - (1024.f + 16.f * static_cast<float>(zb >> 4)) * X);
            if (++j == NU) {
              flush();
              j = 0;
              ++s_cur;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage_i == kSpSPW) { stage_i = 0; ph ^= 1; }
        }
        if (j != 0 && ua < ub) flush();     // the run ended inside a set
        if (lane == 0) {
          wfirst[cw] = first_ls;
          wlast[cw] = last_ls;
        }
      }
      SP_STAMP(4);
      named_bar_sync_gv(1, kSpCons);
      SP_STAMP(5);

      // ---- finish: sum the warps' partial sums in a fixed order, publish (fp16 | tag) and the per-op-path tensors
      {
        const int nsets = nu == 0 ? 0 : (int)((u1 - 1) / NU) - set0 + 1;
        const uint32_t tagw = sp_tag(base, op) << 16;
        uint32_t* out_row = rows + (size_t)(op % kSpRows) * row_stride;
        const __half* bias = o->bias;
        __half* y = o->y;
        for (int t = ct; t < nsets * 8; t += kSpCons) {
          const int ls = t >> 3, gg = t & 7;
          float lo = 0.f, hi = 0.f;
#pragma unroll
          for (int w = 0; w < kSpWarps; ++w) {
            if (wfirst[w] >= 0 && wfirst[w] <= ls && ls <= wlast[w]) {
              const float* p = part + ((size_t)ls * kSpWarps + w) * 16;
              lo += p[gg];
              hi += p[gg + 8];
            }
          }
          int clo, chi;
          sp_cols(mode, N, set0 + ls, gg, clo, chi);
          if (bias != nullptr) {
            lo += __half2float(bias[clo]);
            hi += __half2float(bias[chi]);
          }
          const __half hlo = __float2half_rn(lo), hhi = __float2half_rn(hi);
          y[clo] = hlo;
          y[chi] = hhi;
          if (mode == 0) {
            st_relaxed_u32(out_row + clo, tagw | __half_as_ushort(hlo));
            st_relaxed_u32(out_row + chi, tagw | __half_as_ushort(hhi));
          } else {
            // fused SiLU*mul with the arithmetic of aux.cu's silu_mul_kernel
            const float gf = __half2float(hlo), uf = __half2float(hhi);
            const __half a = __float2half_rn(gf / (1.f + __expf(-gf)) * uf);
            st_relaxed_u32(out_row + clo, tagw | __half_as_ushort(a));
            if (o->act_out != nullptr) o->act_out[clo] = a;
          }
        }
      }
      SP_STAMP(6);
      // (the next op's staging barriers separate these reads of part[] / wfirst[] from the next flushes)
    }
  }

  // ---- the last CTA to leave advances the tag base for the next run (every CTA read it before doing anything)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&state[1], 1) == nblk - 1) {
      state[1] = 0;
      state[0] = (base + n_ops) % 65535;
    }
  }
}

}  // namespace b200awq

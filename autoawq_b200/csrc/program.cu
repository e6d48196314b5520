// Decode program: the whole chain of M = 1 operator calls of a decode step (RMSNorm -> W4A16 linear -> ... ->
// SiLU*mul -> linear) recorded once and executed by ONE persistent kernel launch.
//
// Why (measured, B200, profiles/r01_gemv_v3_phase_timeline.log): a stand-alone GEMV launch spends ~5 us of its
// 7-20 us outside the weight stream - launch + ring fill (the first tile lands after ~5 us of loaded HBM latency),
// the split-K tail, the ticket round trip - and HBM idles through every one of those gaps, 128 times per decode
// step.  The packed weights never depend on the activations, so here the producer warp of every CTA walks the
// WHOLE op list and keeps its shared-memory ring full across op boundaries: while the consumers of op i reduce,
// publish and wait for the grid-wide completion of op i, the tiles of op i+1 are already landing.
//
// Structure (one CTA per SM, launched cooperatively so that all CTAs are co-resident):
//   * producer warp: as in the persistent GEMV (gemv.cu) - lane w feeds consumer warp w's private stages with
//     8 KB weight tiles (TMA 2-D, 128B swizzle) + the tile's group scales / zeros - but over all ops back to back;
//     the tensor maps live in the device-resident op table;
//   * consumers, per op: (1) wait until every column block of the previous op has been published (an acquiring
//     poll of done[op-1]; finalising CTAs release-add to it), (2) stage the op's activations in shared memory,
//     applying the recorded glue op on the fly - RMSNorm (each CTA recomputes the 4096-element norm from L2: 8 KB)
//     or SiLU*mul - with the same arithmetic as the stand-alone kernels (aux.cu), (3) the tile loop / per-group
//     fold / split-K push / ticket / finalise of the persistent GEMV, unchanged (gemv_tile.cuh).
//   CTA 0 also writes the transformed activations to the buffer the recorded glue op named, so every tensor of
//   the per-op path holds the same values after a program run.
//
// Reference call sequence this replaces: awq/modules/fused/block.py:117-170 (norm -> qkv -> ... -> o -> norm ->
// mlp) with awq/modules/fused/mlp.py:41-55 (gate/up GEMM, silu*mul, down GEMM), each a separate awq_ext call.
#include <cuda.h>

#include <cstring>
#include <vector>

#include "../../include/b200awq.h"
#include "common.cuh"
#include "gemv_tile.cuh"
#include "kernels.h"

namespace b200awq {

enum { kProCopy = 0, kProRmsnorm = 1, kProSilu = 2 };

struct __align__(128) ProgOp {
  CUtensorMap tmw;        // qweight [K, N/8] int32, box 32 words x 64 rows, 128B swizzle
  const __half* scales;
  const int32_t* qzeros;
  const __half* bias;
  __half* y;
  const __half* src;      // COPY: x; RMSNORM: the un-normalised row; SILU: gate|up [2K]
  const __half* norm_w;   // RMSNORM weight [K]
  __half* xout;           // where the recorded glue op wanted its result (written by CTA 0), or null
  long long ldsrc;
  int K, N, G, g_shift;
  int prologue;
  float eps;
  int ncb;                // N / 256 column blocks = completion count of this op
  int pad;
};
static_assert(sizeof(ProgOp) == 256, "ProgOp layout");

constexpr int kProgSPW = 2;
constexpr int kProgNS = kV3Warps * kProgSPW;
constexpr int kProgMT = 1;
constexpr size_t kProgFixedSmem = (size_t)kProgNS * (kV3TileBytes + kV3AuxBytes) +
                                  (size_t)(kV3Warps * kProgMT * kGvRedStride + kV3Warps * kProgMT * kV3TileCols) * 4 +
                                  2 * kProgNS * 8 + 128 + 64;
static_assert(kProgFixedSmem % 16 == 0, "xs must stay 16-byte aligned");

__device__ __forceinline__ float prog_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kV3Threads, 1)
    program_kernel(const ProgOp* __restrict__ ops, int n_ops, float* __restrict__ acc_ws, int* __restrict__ tickets,
                   int* __restrict__ done, int M) {
  constexpr int MT = kProgMT, SPW = kProgSPW, NS = kProgNS;
  extern __shared__ __align__(1024) uint8_t pg_smem[];
  uint8_t* ring = pg_smem;
  uint8_t* aux = pg_smem + (size_t)NS * kV3TileBytes;
  float* red = reinterpret_cast<float*>(aux + (size_t)NS * kV3AuxBytes);
  float* colacc = red + kV3Warps * MT * kGvRedStride;
  uint64_t* full = reinterpret_cast<uint64_t*>(colacc + kV3Warps * MT * kV3TileCols);
  uint64_t* empty = full + NS;
  int* flags = reinterpret_cast<int*>(empty + NS);
  int* warp_cb = flags + 16;
  int* warp_ntl = warp_cb + 8;
  float* wsum = reinterpret_cast<float*>(flags + 32);  // 8 floats (+ pad)
  __half* xs = reinterpret_cast<__half*>(pg_smem + kProgFixedSmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(pg_smem) & 1023u) != 0) __trap();
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < kV3Warps * MT * kV3TileCols; i += kV3Threads) colacc[i] = 0.f;
  __syncthreads();

  if (warp == 0) {
    // ============================================================ producer: the weight stream of ALL ops
    if (lane < kV3Warps) {
      const int w = lane;
      int stage_i = 0;
      uint32_t ph = 0;
      for (int op = 0; op < n_ops; ++op) {
        const ProgOp* o = ops + op;
        const int K = o->K, N = o->N, g_shift = o->g_shift;
        const __half* scales = o->scales;
        const int32_t* qzeros = o->qzeros;
        const int NW = N >> 3;
        const int TPC = K / kV3TileRows;
        const int T = (N / kV3TileCols) * TPC;
        const int t0 = (int)((int64_t)T * blockIdx.x / gridDim.x);
        const int t1 = (int)((int64_t)T * (blockIdx.x + 1) / gridDim.x);
        const int ntile = t1 - t0;
        const int a = t0 + (int)((int64_t)ntile * w / kV3Warps);
        const int bnd = t0 + (int)((int64_t)ntile * (w + 1) / kV3Warps);
        int cb = a / TPC, kt = a - cb * TPC;
        for (int t = a; t < bnd; ++t) {
          const int stage = w * SPW + stage_i;
          mbar_wait(&empty[stage], ph ^ 1);
          const int grp_abs = (kt * kV3TileRows) >> g_shift;
          uint8_t* st = ring + (size_t)stage * kV3TileBytes;
          uint8_t* sa = aux + (size_t)stage * kV3AuxBytes;
          mbar_arrive_expect_tx(&full[stage], kV3TileBytes + kV3AuxBytes);
          tma_load_2d(st, &o->tmw, &full[stage], cb * (kV3TileCols / 8), kt * kV3TileRows);
          bulk_load_1d(sa, scales + (int64_t)grp_abs * N + cb * kV3TileCols, kV3ScaleBytes, &full[stage]);
          bulk_load_1d(sa + kV3ScaleBytes, qzeros + (int64_t)grp_abs * NW + cb * (kV3TileCols / 8), kV3ZeroBytes,
                       &full[stage]);
          if (++kt == TPC) { kt = 0; ++cb; }
          if (++stage_i == SPW) { stage_i = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  // ================================================================ consumers
  const int cw = warp - 1;
  const int ct = tid - 32;
  const int g = lane >> 2, tig = lane & 3;
  const bool tok_ok = g < M;
  float* my_red = red + (size_t)cw * MT * kGvRedStride;
  float* my_col = colacc + (size_t)cw * MT * kV3TileCols;
  constexpr int NCT = kV3Warps * 32;

  int stage_i = 0;
  uint32_t ph = 0;
  for (int op = 0; op < n_ops; ++op) {
    const ProgOp* o = ops + op;
    const int K = o->K, N = o->N, G = o->G, g_shift = o->g_shift;
    const __half* bias = o->bias;
    __half* y = o->y;
    const int TPC = K / kV3TileRows;
    const int T = (N / kV3TileCols) * TPC;
    const int t0 = (int)((int64_t)T * blockIdx.x / gridDim.x);
    const int t1 = (int)((int64_t)T * (blockIdx.x + 1) / gridDim.x);
    const int ntile = t1 - t0;
    const int a_w = t0 + (int)((int64_t)ntile * cw / kV3Warps);
    const int b_w = t0 + (int)((int64_t)ntile * (cw + 1) / kV3Warps);

    // ---- (1) the previous op's outputs are this op's inputs: wait until all its column blocks are published
    if (op > 0) {
      if (ct == 0) {
        const int target = ops[op - 1].ncb;
        while (ld_acquire_s32(&done[op - 1]) < target) {
        }
      }
      named_bar_sync_gv(1, NCT);
    }

    // ---- (2) stage (and transform) the activations; arithmetic mirrors aux.cu exactly
    {
      const __half* src = o->src;
      __half* xout = (blockIdx.x == 0) ? o->xout : nullptr;
      const int pro = o->prologue;
      if (pro == kProCopy) {
        for (int i = ct * 8; i < K; i += NCT * 8) *reinterpret_cast<uint4*>(xs + i) = ldcg_u4(src + i);
      } else if (pro == kProSilu) {
        for (int i = ct * 8; i < K; i += NCT * 8) {
          const uint4 gv = ldcg_u4(src + i), uv = ldcg_u4(src + K + i);
          const __half* gh = reinterpret_cast<const __half*>(&gv);
          const __half* uh = reinterpret_cast<const __half*>(&uv);
          uint4 ov;
          __half* oh = reinterpret_cast<__half*>(&ov);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gf = __half2float(gh[j]), uf = __half2float(uh[j]);
            oh[j] = __float2half_rn(gf / (1.f + __expf(-gf)) * uf);
          }
          *reinterpret_cast<uint4*>(xs + i) = ov;
          if (xout != nullptr) *reinterpret_cast<uint4*>(xout + i) = ov;
        }
      } else {
        float ss = 0.f;
        for (int i = ct * 8; i < K; i += NCT * 8) {
          const uint4 v = ldcg_u4(src + i);
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            ss += f.x * f.x + f.y * f.y;
          }
          *reinterpret_cast<uint4*>(xs + i) = v;
        }
        ss = prog_warp_sum(ss);
        if (lane == 0) wsum[cw] = ss;
        named_bar_sync_gv(1, NCT);
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += wsum[i];
        const float rs = rsqrtf(tot / static_cast<float>(K) + o->eps);
        const __half* nw = o->norm_w;
        for (int i = ct * 8; i < K; i += NCT * 8) {   // the thread's own chunks again
          uint4 v = *reinterpret_cast<const uint4*>(xs + i);
          const uint4 wv = *reinterpret_cast<const uint4*>(nw + i);
          __half* vh = reinterpret_cast<__half*>(&v);
          const __half* wh = reinterpret_cast<const __half*>(&wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) vh[j] = __float2half_rn(__half2float(vh[j]) * rs * __half2float(wh[j]));
          *reinterpret_cast<uint4*>(xs + i) = v;
          if (xout != nullptr) *reinterpret_cast<uint4*>(xout + i) = v;
        }
      }
      named_bar_sync_gv(1, NCT);
    }

    // ---- (3) the persistent-GEMV tile loop over this warp's run of tiles
    auto load_x = [&](int t, int ktile, uint32_t (&xb)[4][2]) {
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) xb[bb][0] = xb[bb][1] = 0u;
      if (t < b_w && tok_ok) {
        const __half* px = xs + ktile * kV3TileRows + 2 * tig;   // M = 1: token g = 0
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          xb[bb][0] = *reinterpret_cast<const uint32_t*>(px + 16 * bb);
          xb[bb][1] = *reinterpret_cast<const uint32_t*>(px + 16 * bb + 8);
        }
      }
    };
    float acc[4][4][4];
    float xs_acc[4];
    auto zero_acc = [&]() {
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int tt = 0; tt < 4; ++tt)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[w][tt][r] = 0.f;
      xs_acc[0] = xs_acc[1] = xs_acc[2] = xs_acc[3] = 0.f;
    };
    zero_acc();

    int cur_cb = -1, ntl = 0;
    uint32_t xcur[4][2], xnext[4][2];
    int cb = a_w / TPC, kt = a_w - cb * TPC;
    load_x(a_w, kt, xcur);
    for (int t = a_w; t < b_w; ++t) {
      const int stage = cw * SPW + stage_i;
      if (cb != cur_cb) {
        if (cur_cb >= 0 && ntl > 0) {
          __syncwarp();
          const bool fin = v3_push_warp<MT>(my_col, cur_cb, ntl, TPC, lane, bias, y, acc_ws, tickets, M, N);
          if (fin) {
            __syncwarp();
            if (lane == 0) red_release_add_s32(&done[op], 1);
          }
        }
        cur_cb = cb;
        ntl = 0;
      }
      ++ntl;
      load_x(t + 1, (kt + 1 == TPC) ? 0 : kt + 1, xnext);
      mbar_wait(&full[stage], ph);
      const uint8_t* st = ring + (size_t)stage * kV3TileBytes;
      const uint8_t* sa = aux + (size_t)stage * kV3AuxBytes;
      v3_tile_mma(st, g, tig, xcur, acc, xs_acc);
      const bool group_end = g_shift < 31 ? ((((kt + 1) * kV3TileRows) & (G - 1)) == 0) : (kt + 1 == TPC);
      if (group_end || t + 1 == b_w) {
        v3_fold<MT>(sa, my_red, my_col, lane, g, tig, acc, xs_acc);
        zero_acc();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        xcur[bb][0] = xnext[bb][0];
        xcur[bb][1] = xnext[bb][1];
      }
      if (++kt == TPC) { kt = 0; ++cb; }
      if (++stage_i == SPW) { stage_i = 0; ph ^= 1; }
    }

    // ---- CTA-level reduction, tickets, finalisation (as gemv_v3_kernel), then publish
    if (lane == 0) {
      warp_cb[cw] = (ntl > 0) ? cur_cb : -1;
      warp_ntl[cw] = ntl;
    }
    named_bar_sync_gv(1, NCT);
    {
      int w0 = 0;
      while (w0 < kV3Warps) {
        const int cbg = warp_cb[w0];
        int w1 = w0 + 1;
        while (w1 < kV3Warps && warp_cb[w1] == cbg) ++w1;
        if (cbg >= 0)
          v3_add_cols<MT, NCT>(colacc + (size_t)w0 * MT * kV3TileCols, w1 - w0, MT * kV3TileCols, cbg, ct, acc_ws, M, N);
        w0 = w1;
      }
    }
    named_bar_sync_gv(1, NCT);
    if (ct < kV3Warps) {
      const int w = ct;
      const int cbg = warp_cb[w];
      int is_last = 0;
      if (cbg >= 0 && (w == 0 || warp_cb[w - 1] != cbg)) {
        int tiles = 0;
        for (int w1 = w; w1 < kV3Warps && warp_cb[w1] == cbg; ++w1) tiles += warp_ntl[w1];
        is_last = (atom_add_acq_rel(&tickets[cbg], tiles) + tiles == TPC);
      }
      flags[w] = is_last;
    }
    named_bar_sync_gv(1, NCT);
    int nfin = 0;
#pragma unroll 1
    for (int w = 0; w < kV3Warps; ++w)
      if (flags[w]) {
        v3_finalize<MT, NCT>(warp_cb[w], ct, bias, y, acc_ws, tickets, M, N);
        ++nfin;
      }
    if (nfin > 0) {   // CTA-uniform
      named_bar_sync_gv(1, NCT);
      if (ct == 0) red_release_add_s32(&done[op], nfin);
    }
    // (the barrier after the completion wait of the next op separates these shared-memory reads from its writes)
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct Program {
  ProgOp* d_ops = nullptr;
  int* d_done = nullptr;
  int n_ops = 0;
  int M = 0;
  int max_N = 0;
  size_t smem = 0;
  int device = 0;
};

static int prog_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = B200AWQ_SM_COUNT_FALLBACK;
  return n;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool overlaps(const void* a, size_t na, const void* b, size_t nb) {
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(a), b0 = reinterpret_cast<uintptr_t>(b);
  return a0 < b0 + nb && b0 < a0 + na;
}

// Folds the recorded call sequence into linear ops with an activation prologue.  Returns a B200AWQ_* code;
// *cuda_err carries the CUDA error behind B200AWQ_ECUDA.
//
// Hazard rules (the kernel orders ops only through "every column block of op i-1 is published"):
//   * a glue op (RMSNorm / SiLU*mul) is executed as the prologue of every later linear that reads its output
//     buffer; CTA 0 writes that buffer as a side effect, nobody inside the kernel may READ it;
//   * a linear must not write (y) what it reads (src) or what its own prologue publishes (xout);
//   * a buffer that a pending glue record depends on must not be overwritten before the record's last use.
int program_create(const b200awq_op_t* ops, int n, Program** out, cudaError_t* cuda_err) {
  *cuda_err = cudaSuccess;
  *out = nullptr;
  if (ops == nullptr || n <= 0) return B200AWQ_EINVAL;
  std::vector<ProgOp> table;
  struct Glue {
    int kind;
    const void* src;
    const void* w;
    void* out;
    int width;
    float eps;
    bool used;
    bool live;
  };
  std::vector<Glue> glues;
  const int grid = prog_sm_count();
  int max_K = 0, max_N = 0, M = -1;
  for (int i = 0; i < n; ++i) {
    const b200awq_op_t& op = ops[i];
    if (M < 0) M = op.M;
    if (op.M != M) return B200AWQ_EUNSUPPORTED;
    if (op.kind == B200AWQ_OP_RMSNORM || op.kind == B200AWQ_OP_SILU_AND_MUL) {
      if (op.x == nullptr || op.y == nullptr || op.K <= 0) return B200AWQ_EINVAL;
      if (op.kind == B200AWQ_OP_RMSNORM && op.weight == nullptr) return B200AWQ_EINVAL;
      if ((op.K % 8) != 0 || !aligned16(op.x) || !aligned16(op.y) || (op.weight != nullptr && !aligned16(op.weight)))
        return B200AWQ_EUNSUPPORTED;
      const size_t in_bytes = (size_t)(op.kind == B200AWQ_OP_SILU_AND_MUL ? 2 : 1) * op.K * 2;
      if (overlaps(op.y, (size_t)op.K * 2, op.x, in_bytes)) return B200AWQ_EUNSUPPORTED;  // in-place glue op
      for (Glue& gl : glues)
        if (gl.live && (overlaps(gl.out, (size_t)gl.width * 2, op.y, (size_t)op.K * 2) ||
                        overlaps(gl.src, (size_t)(gl.kind == kProSilu ? 2 : 1) * gl.width * 2, op.y, (size_t)op.K * 2))) {
          if (!gl.used) return B200AWQ_EUNSUPPORTED;
          gl.live = false;
        }
      // its input must not be a buffer only CTA 0 publishes
      for (const Glue& gl : glues)
        if (overlaps(gl.out, (size_t)gl.width * 2, op.x, in_bytes)) return B200AWQ_EUNSUPPORTED;
      glues.push_back(Glue{op.kind == B200AWQ_OP_RMSNORM ? kProRmsnorm : kProSilu, op.x, op.weight, op.y, op.K, op.eps,
                           false, true});
      continue;
    }
    if (op.kind != B200AWQ_OP_LINEAR_GEMM) return B200AWQ_EINVAL;
    if (op.x == nullptr || op.qweight == nullptr || op.scales == nullptr || op.qzeros == nullptr || op.y == nullptr ||
        op.K <= 0 || op.N <= 0 || op.group_size <= 0 || (op.K % op.group_size) != 0)
      return B200AWQ_EINVAL;
    GemmArgs a{op.x, op.ldx, static_cast<const int32_t*>(op.qweight), op.scales, static_cast<const int32_t*>(op.qzeros),
               op.bias, op.y, op.M, op.K, op.N, op.group_size};
    if (M != 1 || !gemv_v3_supported(a)) return B200AWQ_EUNSUPPORTED;
    if ((op.N / kV3TileCols) * (op.K / kV3TileRows) < grid) return B200AWQ_EUNSUPPORTED;  // every CTA owns tiles
    ProgOp p;
    std::memset(&p, 0, sizeof(p));
    cudaError_t e = make_tmap_2d(a.qweight, /*int32*/ 1, (uint64_t)(a.N / 8), (uint64_t)a.K, (uint64_t)(a.N / 8) * 4, 32,
                                 kV3TileRows, &p.tmw);
    if (e != cudaSuccess) {
      *cuda_err = e;
      return B200AWQ_ECUDA;
    }
    p.scales = static_cast<const __half*>(op.scales);
    p.qzeros = static_cast<const int32_t*>(op.qzeros);
    p.bias = static_cast<const __half*>(op.bias);
    p.y = static_cast<__half*>(op.y);
    p.K = op.K;
    p.N = op.N;
    p.G = op.group_size;
    p.g_shift = 31;
    if ((p.G & (p.G - 1)) == 0) {
      p.g_shift = 0;
      while ((1 << p.g_shift) < p.G) ++p.g_shift;
    }
    p.ncb = op.N / kV3TileCols;
    p.ldsrc = op.ldx;
    Glue* hit = nullptr;
    for (Glue& gl : glues)
      if (gl.live && gl.out == op.x && gl.width == op.K) hit = &gl;
    if (hit != nullptr) {
      p.prologue = hit->kind;
      p.src = static_cast<const __half*>(hit->src);
      p.norm_w = static_cast<const __half*>(hit->w);
      p.xout = hit->used ? nullptr : static_cast<__half*>(hit->out);   // published once, by its first consumer
      p.eps = hit->eps;
      hit->used = true;
    } else {
      p.prologue = kProCopy;
      p.src = static_cast<const __half*>(op.x);
      if (!aligned16(op.x)) return B200AWQ_EUNSUPPORTED;
      for (const Glue& gl : glues)   // reading a buffer only CTA 0 publishes (a dead or mismatching record)
        if (overlaps(gl.out, (size_t)gl.width * 2, op.x, (size_t)op.K * 2)) return B200AWQ_EUNSUPPORTED;
    }
    const size_t src_bytes = (size_t)(p.prologue == kProSilu ? 2 : 1) * op.K * 2;
    if (overlaps(p.y, (size_t)op.N * 2, p.src, src_bytes)) return B200AWQ_EUNSUPPORTED;
    if (p.xout != nullptr && overlaps(p.xout, (size_t)op.K * 2, p.y, (size_t)op.N * 2)) return B200AWQ_EUNSUPPORTED;
    // writing y over something a live glue record still needs ends that record
    for (Glue& gl : glues)
      if (gl.live && &gl != hit &&
          (overlaps(gl.out, (size_t)gl.width * 2, p.y, (size_t)op.N * 2) ||
           overlaps(gl.src, (size_t)(gl.kind == kProSilu ? 2 : 1) * gl.width * 2, p.y, (size_t)op.N * 2))) {
        if (!gl.used) return B200AWQ_EUNSUPPORTED;
        gl.live = false;
      }
    max_K = op.K > max_K ? op.K : max_K;
    max_N = op.N > max_N ? op.N : max_N;
    table.push_back(p);
  }
  for (const Glue& gl : glues)
    if (!gl.used) return B200AWQ_EUNSUPPORTED;   // a glue op nobody consumes would never run
  if (table.empty()) return B200AWQ_EUNSUPPORTED;
  const size_t smem = kProgFixedSmem + (size_t)(max_K + 8) * 2 * kProgMT;
  if (smem > (size_t)227 * 1024) return B200AWQ_EUNSUPPORTED;

  Program* pr = new Program();
  pr->n_ops = static_cast<int>(table.size());
  pr->M = M;
  pr->max_N = max_N;
  pr->smem = smem;
  cudaError_t e = cudaGetDevice(&pr->device);
  if (e == cudaSuccess) e = cudaMalloc(&pr->d_ops, table.size() * sizeof(ProgOp));
  if (e == cudaSuccess) e = cudaMalloc(&pr->d_done, table.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(pr->d_ops, table.data(), table.size() * sizeof(ProgOp), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  if (e != cudaSuccess) {
    cudaFree(pr->d_ops);
    cudaFree(pr->d_done);
    delete pr;
    *cuda_err = e;
    return B200AWQ_ECUDA;
  }
  *out = pr;
  return B200AWQ_OK;
}

int program_max_n(const Program* p) { return p->max_N; }
int program_m(const Program* p) { return p->M; }
int program_num_ops(const Program* p) { return p->n_ops; }

cudaError_t program_run(Program* p, float* acc_ws, int* tickets, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(p->d_done, 0, (size_t)p->n_ops * sizeof(int), st);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(prog_sm_count());
  cfg.blockDim = dim3(kV3Threads);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the completion counters are grid-wide waits
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const ProgOp* ops = p->d_ops;
  return cudaLaunchKernelEx(&cfg, program_kernel, ops, p->n_ops, acc_ws, tickets, p->d_done, p->M);
}

void program_destroy(Program* p) {
  if (p == nullptr) return;
  cudaFree(p->d_ops);
  cudaFree(p->d_done);
  delete p;
}

}  // namespace b200awq

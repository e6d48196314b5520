/* Plain-C restatement of the reference's W4A16 dequantisation + contraction.  TEST INFRASTRUCTURE ONLY
 * (checker for tests/, smoke() and bench.py's cpu_baseline; never linked into the product library).
 * Follows awq/utils/packing_utils.py:8-43,87-102 (unpack, AWQ de-interleave, (q - z) * s in fp16) and the
 * naive forward branch awq/modules/linear/gemm.py:71-77.  Independent of the numpy oracle on purpose:
 * tests/test_oracle_c.py requires the two to agree bit for bit.  Parity pinning: through the numpy oracle's
 * golden vectors (tests/golden/*.npz, generated from the real reference). */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

/* round-to-nearest-even float -> half */
static uint16_t float_to_half(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((x > 0x7F800000u) ? 0x200u : 0)); /* inf / nan */
  if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                                        /* overflow */
  if (x < 0x33000001u) return (uint16_t)sign;                                                      /* underflow to 0 */
  int e = (int)(x >> 23) - 127;
  uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
  int shift;
  uint16_t base;
  if (e < -14) { shift = 13 + (-14 - e); base = 0; }
  else { shift = 13; base = (uint16_t)((e + 15) << 10); man &= 0x7FFFFFu; }
  uint32_t q = man >> shift, rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1))) ++q;
  return (uint16_t)(sign | (base + q));
}

static const int kShift[8] = {0, 16, 4, 20, 8, 24, 12, 28}; /* 4 * AWQ_REVERSE_ORDER[j] */

/* W[K, N] fp16 = (nib - znib) * s, GEMM layout */
void oracle_dequantize_gemm(const int32_t* qweight, const int32_t* qzeros, const uint16_t* scales, uint16_t* out,
                            int K, int N, int G) {
  const int NW = N / 8;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k) {
    const int g = k / G;
    for (int n = 0; n < N; ++n) {
      const uint32_t w = (uint32_t)qweight[(int64_t)k * NW + n / 8], z = (uint32_t)qzeros[(int64_t)g * NW + n / 8];
      const int d = (int)((w >> kShift[n % 8]) & 0xF) - (int)((z >> kShift[n % 8]) & 0xF);
      /* int -> fp16 exact; fp16 * fp16 rounded once: the fp32 product of two halves is exact */
      out[(int64_t)k * N + n] = float_to_half((float)d * half_to_float(scales[(int64_t)g * N + n]));
    }
  }
}

/* Y[M, N] (double) = X[M, K] fp16 . W[K, N] fp16 */
void oracle_gemm_f64(const uint16_t* x, const uint16_t* w, double* y, int M, int K, int N) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n) {
    for (int m = 0; m < M; ++m) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k)
        acc += (double)half_to_float(x[(int64_t)m * K + k]) * (double)half_to_float(w[(int64_t)k * N + n]);
      y[(int64_t)m * N + n] = acc;
    }
  }
}

uint16_t oracle_float_to_half(float f) { return float_to_half(f); }
float oracle_half_to_float(uint16_t h) { return half_to_float(h); }

/* ---- MoE routing (awq/modules/fused/moe.py:92-171): independent C restatement of the two index-producing steps ---- */

/* moe_align_block_size (moe.py:92-134): slots grouped by expert in ascending slot order, each run padded with `numel`
 * to a multiple of block_size.  sorted_ids holds numel + num_experts * (block_size - 1) entries (pre-filled here with
 * numel), expert_ids numel + num_experts (pre-filled with -1).  Returns the padded length. */
int oracle_moe_align_block_size(const int32_t* topk_ids, int numel, int num_experts, int block_size,
                                int32_t* sorted_ids, int32_t* expert_ids) {
  int pos = 0, blk = 0;
  for (int i = 0; i < numel + num_experts * (block_size - 1); ++i) sorted_ids[i] = numel;
  for (int i = 0; i < numel + num_experts; ++i) expert_ids[i] = -1;
  for (int e = 0; e < num_experts; ++e) {
    int cnt = 0;
    for (int i = 0; i < numel; ++i)
      if (topk_ids[i] == e) sorted_ids[pos + cnt++] = i;
    if (cnt == 0) continue;
    const int padded = (cnt + block_size - 1) / block_size * block_size;
    for (int b = 0; b < padded / block_size; ++b) expert_ids[blk + b] = e;
    pos += padded;
    blk += padded / block_size;
  }
  return pos;
}

/* top-k of softmax(gating) per row, ties to the lower index; weights are the softmax probabilities (double
 * arithmetic, rounded to float once). */
void oracle_topk_softmax(const float* gating, int M, int E, int topk, float* weights, int32_t* ids) {
  for (int m = 0; m < M; ++m) {
    const float* g = gating + (size_t)m * E;
    double mx = g[0], sum = 0.0;
    for (int e = 1; e < E; ++e) mx = g[e] > mx ? g[e] : mx;
    for (int e = 0; e < E; ++e) sum += exp((double)g[e] - mx);
    for (int k = 0; k < topk; ++k) {
      int best = -1;
      for (int e = 0; e < E; ++e) {
        int taken = 0;
        for (int j = 0; j < k; ++j) taken |= ids[(size_t)m * topk + j] == e;
        if (taken) continue;
        if (best < 0 || g[e] > g[best]) best = e;
      }
      ids[(size_t)m * topk + k] = best;
      weights[(size_t)m * topk + k] = (float)(exp((double)g[best] - mx) / sum);
    }
  }
}

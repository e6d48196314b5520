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

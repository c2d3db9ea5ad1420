// fp_inv_gcd.cuh -- Fq inversion by Bernstein-Yang division steps ("safegcd") on signed 30-bit limbs.
//
// The reference inverts with a binary extended Euclid (bls12_381/fq.rs:849-902); the inverse of a field element is unique
// and every value leaves this library canonical, so ANY inversion algorithm returns the reference's bits.  Round 1 used a
// Fermat power a^(q-2): 380 squarings + ~105 products = 120 k MAC32 per lane, 5.4 % of a whole pairing.  Division steps
// need no multi-word arithmetic per step: 30 steps at a time are run on the low words of (f, g) alone while a 2x2 integer
// matrix records them; the matrix is then applied to the full (f, g) and -- modulo q, with an exact division by 2^30 --
// to the pair (d, e) that tracks the inverse.  ~27 batches for a 381-bit operand, each ~8 short loop trips + 130
// 32x32->64 multiply-accumulates: about a tenth of the Fermat chain.
//
// Variable-time form: every lane runs until ITS g is zero, a batch cancels up to four low bits of g per trip (w = -g / f
// mod 2^k, k <= 4) and skips runs of zero bits with one count-trailing-zeros.  Lanes of a warp
// diverge inside; nothing in here synchronises or shuffles, so the routine is safe in divergent code (curve kernels) and
// the lane-pair kernels re-converge at their next shuffle.
//
// Values: d, e in (-2q, q) as in the published algorithm (13 limbs of 30 bits, top limb signed); f, g likewise signed.
// e starts at R^2 mod q instead of 1, so that for a Montgomery operand A = a R the result A^-1 R^2 = a^-1 R is already
// the Montgomery form of the inverse -- no multiplication at the end.
// The arithmetic is plain C on int32 / int64 (compiled for the host by tests/cpp/test_fp_inv_gcd.cpp and checked there
// against Python's pow(a, -1, q) vectors); ptxas turns the 64-bit accumulations into signed IMAD.WIDE.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BLS_GCD_HD __host__ __device__ __forceinline__
#else
#define BLS_GCD_HD inline
#endif

namespace bls {
namespace gcd30 {

constexpr int N = 13;                       // 13 x 30 = 390 bits >= 381 + sign + the factor 2 of (-2q, q)
constexpr int32_t M30 = 0x3fffffff;
// q, R^2 mod q (R = 2^384) in 30-bit limbs, and q^-1 mod 2^30 (recomputed and checked by tests/test_host_logic.py)
#define BLS_GCD_Q30                                                                                                   \
  { 0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673, 0x12e13ce1, 0x2cd76477,     \
    0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111 }
#define BLS_GCD_R2_30                                                                                                 \
  { 0x1c341746, 0x137c7cd0, 0x1d104f1f, 0x1db9a982, 0x15b6d50a, 0x151db132, 0x183c08de, 0x222a64e7, 0x152d67eb,     \
    0x3a16d466, 0x3aa9a793, 0x3964b2b8, 0x0011988f }
constexpr uint32_t QINV30 = 0x00030003u;

BLS_GCD_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

// 12 x 32-bit words (a canonical value < q) -> 13 x 30-bit limbs
BLS_GCD_HD void from_words(int32_t (&r)[N], const uint32_t (&a)[12]) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int bit = 30 * i, w = bit >> 5, s = bit & 31;
    uint32_t lo = a[w] >> s;
    if (s > 2 && w + 1 < 12) lo |= a[w + 1] << (32 - s);
    r[i] = (int32_t)(lo & (uint32_t)M30);
  }
}
// 13 x 30-bit limbs of a value in [0, q) -> 12 x 32-bit words
BLS_GCD_HD void to_words(uint32_t (&a)[12], const int32_t (&r)[N]) {
#pragma unroll
  for (int j = 0; j < 12; j++) {
    const int bit = 32 * j, i = bit / 30, off = bit % 30;       // off <= 22: two limbs cover the word
    a[j] = ((uint32_t)r[i] >> off) | ((uint32_t)r[i + 1] << (30 - off));
  }
}

struct Trans { int32_t u, v, q, r; };

// 30 division steps on the low words; eta = -delta.  Returns the new eta and the matrix t with
//   t * [f, g] = 2^30 * [f', g']
BLS_GCD_HD int32_t divsteps_30_var(int32_t eta, uint32_t f0, uint32_t g0, Trans& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
  int i = 30;
  for (;;) {
    const int zeros = ctz32(g | (0xffffffffu << i));            // at most i: the sentinel bit
    g >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
    if (i == 0) break;
    if (eta < 0) {                                              // delta > 0 and g odd: (f, g) <- (g, -f)
      uint32_t tmp;
      eta = -eta;
      tmp = f; f = g; g = 0u - tmp;
      tmp = u; u = q; q = 0u - tmp;
      tmp = v; v = r; r = 0u - tmp;
    }
    // cancel the low min(eta + 1, i, 4) bits of g with a multiple of f (no more than eta + 1: the sign of delta flips
    // there; wider cancellations were measured to save no trips -- eta + 1 is what limits them)
    const int limit = (eta + 1) > i ? i : (eta + 1);
    const uint32_t m = (0xffffffffu >> (32 - limit)) & 15u;
    uint32_t w = f + (((f + 1u) & 4u) << 1);                    // 1 / f (mod 16) for odd f
    w = ((0u - w) * g) & m;
    g += f * w; q += u * w; r += v * w;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return eta;
}

// acc + a * b with a 32 x 32 -> 64 signed product (mad.wide.s32: one IMAD.WIDE)
// (inline PTX on the device: the compiler otherwise widens the operands it knows to be masked 64-bit values and emits
// 64 x 64-bit multiplies -- three multiply instructions instead of one)
BLS_GCD_HD int64_t mac(int64_t acc, int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
  int64_t r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(acc));
  return r;
#else
  return acc + (int64_t)a * (int64_t)b;
#endif
}

// [f, g] <- t * [f, g] / 2^30 (exact)
BLS_GCD_HD void update_fg(int32_t (&f)[N], int32_t (&g)[N], const Trans& t) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  int64_t cf = mac(mac(0, u, f[0]), v, g[0]);
  int64_t cg = mac(mac(0, q, f[0]), r, g[0]);
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < N; i++) {
    const int32_t fi = f[i], gi = g[i];
    cf = mac(mac(cf, u, fi), v, gi);
    cg = mac(mac(cg, q, fi), r, gi);
    f[i - 1] = (int32_t)cf & M30; cf >>= 30;
    g[i - 1] = (int32_t)cg & M30; cg >>= 30;
  }
  f[N - 1] = (int32_t)cf;
  g[N - 1] = (int32_t)cg;
}

// [d, e] <- t * [d, e] / 2^30 mod q; both stay in (-2q, q)
BLS_GCD_HD void update_de(int32_t (&d)[N], int32_t (&e)[N], const Trans& t) {
  const int32_t Q30[N] = BLS_GCD_Q30;
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d[N - 1] >> 31, se = e[N - 1] >> 31;
  int32_t md = (u & sd) + (v & se);
  int32_t me = (q & sd) + (r & se);
  int64_t cd = mac(mac(0, u, d[0]), v, e[0]);
  int64_t ce = mac(mac(0, q, d[0]), r, e[0]);
  // the multiples of q that make the low 30 bits zero
  md -= (int32_t)((QINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((QINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd = mac(cd, Q30[0], md);
  ce = mac(ce, Q30[0], me);
  cd >>= 30; ce >>= 30;
#pragma unroll
  for (int i = 1; i < N; i++) {
    const int32_t di = d[i], ei = e[i];
    cd = mac(mac(mac(cd, u, di), v, ei), Q30[i], md);
    ce = mac(mac(mac(ce, q, di), r, ei), Q30[i], me);
    d[i - 1] = (int32_t)cd & M30; cd >>= 30;
    e[i - 1] = (int32_t)ce & M30; ce >>= 30;
  }
  d[N - 1] = (int32_t)cd;
  e[N - 1] = (int32_t)ce;
}

// r in (-2q, q) -> [0, q), negated first when sign < 0
BLS_GCD_HD void normalize(int32_t (&r)[N], int32_t sign) {
  const int32_t Q30[N] = BLS_GCD_Q30;
  int32_t cond_add = r[N - 1] >> 31;
  const int32_t cond_neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < N; i++) {
    r[i] += Q30[i] & cond_add;
    r[i] = (r[i] ^ cond_neg) - cond_neg;
  }
#pragma unroll
  for (int i = 0; i < N - 1; i++) { r[i + 1] += r[i] >> 30; r[i] &= M30; }
  cond_add = r[N - 1] >> 31;
#pragma unroll
  for (int i = 0; i < N; i++) r[i] += Q30[i] & cond_add;
#pragma unroll
  for (int i = 0; i < N - 1; i++) { r[i + 1] += r[i] >> 30; r[i] &= M30; }
}

// out = a^-1 * 2^768 mod q for a canonical a < q (words); a == 0 gives 0.  Returns the number of batches (diagnostics).
BLS_GCD_HD int invert_words(uint32_t (&out)[12], const uint32_t (&a)[12]) {
  int32_t f[N] = BLS_GCD_Q30, e[N] = BLS_GCD_R2_30, d[N], g[N];
#pragma unroll
  for (int i = 0; i < N; i++) d[i] = 0;
  from_words(g, a);
  int32_t eta = -1;
  int batches = 0;
#pragma unroll 1
  for (;;) {
    int32_t nz = g[0];
#pragma unroll
    for (int i = 1; i < N; i++) nz |= g[i];
    if (nz == 0) break;
    Trans t;
    eta = divsteps_30_var(eta, (uint32_t)f[0], (uint32_t)g[0], t);
    update_de(d, e, t);
    update_fg(f, g, t);
    batches++;
  }
  normalize(d, f[N - 1]);                                     // g = 0: f = +-gcd = +-1 (or q for a = 0, where d = 0)
  to_words(out, d);
  return batches;
}

}  // namespace gcd30
}  // namespace bls

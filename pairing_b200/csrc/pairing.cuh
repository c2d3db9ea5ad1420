// pairing.cuh -- the BLS12-381 optimal-ate pairing as the reference computes it
// (bls12_381/mod.rs:40-358): G2 line-coefficient steps, `ell`, the Miller loop over
// BLS_X >> 1 = 0x6900800000008000 and the final exponentiation.
// The raw Miller value and the G2Prepared coefficients are API-visible in the reference, so the
// line-coefficient scaling of Algorithms 26/27 (eprint 2010/354) is kept exactly.
#pragma once
#include "curve.cuh"

namespace bls {

struct Coeffs { Fp2 c0, c1, c2; };

// mod.rs:176-245
__device__ __noinline__ void g2_doubling_step(Jac<Fp2>& r, Coeffs& out) {
  Fp2 tmp0 = fp2_sqr(r.x);
  Fp2 tmp1 = fp2_sqr(r.y);
  Fp2 tmp2 = fp2_sqr(tmp1);
  Fp2 tmp3 = fp2_dbl(fp2_sub(fp2_sub(fp2_sqr(fp2_add(tmp1, r.x)), tmp0), tmp2));
  Fp2 tmp4 = fp2_add(fp2_dbl(tmp0), tmp0);
  Fp2 tmp6 = fp2_add(r.x, tmp4);
  Fp2 tmp5 = fp2_sqr(tmp4);
  Fp2 zsq = fp2_sqr(r.z);
  r.x = fp2_sub(fp2_sub(tmp5, tmp3), tmp3);
  r.z = fp2_sub(fp2_sub(fp2_sqr(fp2_add(r.z, r.y)), tmp1), zsq);
  r.y = fp2_sub(fp2_mul(fp2_sub(tmp3, r.x), tmp4), fp2_dbl(fp2_dbl(fp2_dbl(tmp2))));
  out.c1 = fp2_neg(fp2_dbl(fp2_mul(tmp4, zsq)));
  out.c2 = fp2_sub(fp2_sub(fp2_sub(fp2_sqr(tmp6), tmp0), tmp5), fp2_dbl(fp2_dbl(tmp1)));
  out.c0 = fp2_dbl(fp2_mul(r.z, zsq));
}

// mod.rs:247-333
__device__ __noinline__ void g2_addition_step(Jac<Fp2>& r, const Fp2& qx, const Fp2& qy, Coeffs& out) {
  Fp2 zsq = fp2_sqr(r.z);
  Fp2 ysq = fp2_sqr(qy);
  Fp2 t0 = fp2_mul(zsq, qx);
  Fp2 t1 = fp2_mul(fp2_sub(fp2_sub(fp2_sqr(fp2_add(qy, r.z)), ysq), zsq), zsq);
  Fp2 t2 = fp2_sub(t0, r.x);
  Fp2 t3 = fp2_sqr(t2);
  Fp2 t4 = fp2_dbl(fp2_dbl(t3));
  Fp2 t5 = fp2_mul(t4, t2);
  Fp2 t6 = fp2_sub(fp2_sub(t1, r.y), r.y);
  Fp2 t9 = fp2_mul(t6, qx);
  Fp2 t7 = fp2_mul(t4, r.x);
  r.x = fp2_sub(fp2_sub(fp2_sub(fp2_sqr(t6), t5), t7), t7);
  r.z = fp2_sub(fp2_sub(fp2_sqr(fp2_add(r.z, t2)), zsq), t3);
  Fp2 t10 = fp2_add(qy, r.z);
  Fp2 t8 = fp2_mul(fp2_sub(t7, r.x), t6);
  r.y = fp2_sub(t8, fp2_dbl(fp2_mul(r.y, t5)));
  t10 = fp2_sub(fp2_sub(fp2_sqr(t10), ysq), fp2_sqr(r.z));
  out.c2 = fp2_sub(fp2_dbl(t9), t10);
  out.c0 = fp2_dbl(r.z);
  out.c1 = fp2_dbl(fp2_neg(t6));
}

// mod.rs:57-69
__device__ __forceinline__ void ell(Fp12& f, const Coeffs& c, const Fp& px, const Fp& py) {
  Fp2 c0 = fp2_mul_fp(c.c0, py);
  Fp2 c1 = fp2_mul_fp(c.c1, px);
  fp12_mul_by_014(f, c.c2, c1, c0);
}

// The loop schedule: bits of BLS_X >> 1 below the leading one, MSB first (mod.rs:72-78).
#define BLS_LOOP_BITS (BLS_X_ABS >> 1)
#define BLS_LOOP_TOP 61   /* bit 62 is the leading one */

// Single-pair Miller loop with the G2 steps computed on the fly (value-identical to
// prepare-then-loop: the coefficient sequence is consumed in generation order, mod.rs:345-349).
// `live` == false reproduces the reference's skipping of pairs with an infinity member: f = 1.
__device__ __forceinline__ void miller_loop_single(Fp12& f, const Fp& px, const Fp& py, const Fp2& qx, const Fp2& qy, bool live) {
  fp12_one(f);
  if (!live) return;   // conjugate(1) == 1
  Jac<Fp2> r; r.x = qx; r.y = qy; r.z = fp2_one();
  Coeffs c;
#pragma unroll 1
  for (int b = BLS_LOOP_TOP; b >= 0; b--) {
    g2_doubling_step(r, c);
    ell(f, c, px, py);
    if ((BLS_LOOP_BITS >> b) & 1ull) {
      g2_addition_step(r, qx, qy, c);
      ell(f, c, px, py);
    }
    fp12_sqr(f, f);
  }
  g2_doubling_step(r, c);
  ell(f, c, px, py);
  fp12_conjugate(f);   // BLS_X_IS_NEGATIVE
}

// exp_by_x (mod.rs:116-121): Field::pow(&[x]) (lib.rs:306-324) followed by a conjugation.
// Two value-preserving shortcuts (SURVEY.md 8c "latitude"): the leading `one * self` product of pow
// is a copy, and -- the operand being in the cyclotomic subgroup (it is only called after the easy
// part) -- every squaring is a Granger-Scott cyclotomic squaring.
__device__ __noinline__ void fp12_exp_by_x(Fp12& out, const Fp12& a, uint64_t x) {
  Fp12 res = a;
  const int top = 63 - __clzll((long long)x);
#pragma unroll 1
  for (int n = top - 1; n >= 0; n--) {
    fp12_cyclotomic_sqr(res, res);
    if ((x >> n) & 1ull) fp12_mul(res, res, a);
  }
  fp12_conjugate(res);
  out = res;
}

// mod.rs:104-160.  Returns false for a zero input (the reference's None); `out` is then zero.
__device__ __forceinline__ bool final_exponentiation(Fp12& out, const Fp12& in) {
  Fp12 f1 = in, f2, r;
  fp12_conjugate(f1);
  if (!fp12_inv(f2, in)) { out.c0 = fp6_zero(); out.c1 = fp6_zero(); return false; }
  fp12_mul(r, f1, f2);
  f2 = r;
  fp12_frobenius(r, r, 2);
  fp12_mul(r, r, f2);
  const uint64_t x = BLS_X_ABS;
  Fp12 y0, y1, y2, y3;
  fp12_sqr(y0, r);
  fp12_exp_by_x(y1, y0, x);
  fp12_exp_by_x(y2, y1, x >> 1);
  y3 = r; fp12_conjugate(y3);
  fp12_mul(y1, y1, y3);
  fp12_conjugate(y1);
  fp12_mul(y1, y1, y2);
  fp12_exp_by_x(y2, y1, x);
  fp12_exp_by_x(y3, y2, x);
  fp12_conjugate(y1);
  fp12_mul(y3, y3, y1);
  fp12_conjugate(y1);
  fp12_frobenius(y1, y1, 3);
  fp12_frobenius(y2, y2, 2);
  fp12_mul(y1, y1, y2);
  fp12_exp_by_x(y2, y3, x);
  fp12_mul(y2, y2, y0);
  fp12_mul(y2, y2, r);
  fp12_mul(y1, y1, y2);
  fp12_frobenius(y2, y3, 1);
  fp12_mul(y1, y1, y2);
  out = y1;
  return true;
}

}  // namespace bls

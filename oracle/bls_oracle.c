/* bls_oracle.c -- CPU restatement (plain C, 6 x u64 limbs, unsigned __int128) of the BLS12-381
 * pairing / wNAF path of the `pairing` crate v0.14.2 (dignifiedquire/pairing).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load this
 * library, and only as the checker / the timed CPU baseline -- never on the product path.
 *
 * The reference is Rust and this image has no Rust toolchain, so the reference itself cannot be
 * compiled here (no oracle/_ref).  This file follows the reference function by function, in
 * Montgomery form, with the same formulas, operation order and special cases, so that every
 * output -- including representative-dependent ones (Jacobian triples, raw Miller values,
 * G2Prepared coefficients) -- is bit-identical to the crate's.  It is pinned by the reference's
 * own known-answer tests (tests/test_oracle_kat.py): RELIC pairing vector, k*G vectors,
 * Montgomery constants, Fq/Fq2 KATs -- and cross-checked against the independent big-integer
 * model in bls_model.py.
 *
 * Citations are file:line into /root/reference/src.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "constants.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fq;
typedef struct { fq c0, c1; } fq2;
typedef struct { fq2 c0, c1, c2; } fq6;
typedef struct { fq6 c0, c1; } fq12;
typedef struct { fq x, y; uint64_t inf; } g1_affine;
typedef struct { fq x, y, z; } g1;
typedef struct { fq2 x, y; uint64_t inf; } g2_affine;
typedef struct { fq2 x, y, z; } g2;
typedef struct { uint64_t l[4]; } fr_repr;
typedef struct { fq2 c[68][3]; uint64_t inf; } g2_prepared;

/* ---------------------------------------------------------------- limb primitives, lib.rs:646-679 */
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t *carry) {
  u128 t = (u128)a + b + *carry;
  *carry = (uint64_t)(t >> 64);
  return (uint64_t)t;
}
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t *borrow) {
  u128 t = ((u128)1 << 64) + a - b - *borrow;
  *borrow = (t >> 64) == 0;
  return (uint64_t)t;
}
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t *carry) {
  u128 t = (u128)a + (u128)b * c + *carry;
  *carry = (uint64_t)(t >> 64);
  return (uint64_t)t;
}

/* ---------------------------------------------------------------- FqRepr helpers, fq.rs:510-697 */
static const fq *MOD(void) { return (const fq *)BLS_MODULUS; }
static const fq *ONE(void) { return (const fq *)BLS_R; }

static inline int repr_is_zero(const uint64_t *a, int n) {
  uint64_t o = 0;
  for (int i = 0; i < n; i++) o |= a[i];
  return o == 0;
}
static inline int repr_cmp(const uint64_t *a, const uint64_t *b, int n) { /* -1,0,1 */
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
static inline void repr_add_nocarry(uint64_t *a, const uint64_t *b, int n) {
  uint64_t c = 0;
  for (int i = 0; i < n; i++) a[i] = adc(a[i], b[i], &c);
}
static inline void repr_sub_noborrow(uint64_t *a, const uint64_t *b, int n) {
  uint64_t c = 0;
  for (int i = 0; i < n; i++) a[i] = sbb(a[i], b[i], &c);
}
static inline void repr_div2(uint64_t *a, int n) {
  uint64_t t = 0;
  for (int i = n - 1; i >= 0; i--) {
    uint64_t t2 = a[i] << 63;
    a[i] = (a[i] >> 1) | t;
    t = t2;
  }
}
static inline void repr_mul2(uint64_t *a, int n) {
  uint64_t last = 0;
  for (int i = 0; i < n; i++) {
    uint64_t tmp = a[i] >> 63;
    a[i] = (a[i] << 1) | last;
    last = tmp;
  }
}
static inline int repr_num_bits(const uint64_t *a, int n) { /* fr.rs:213-225 */
  int ret = n * 64;
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] == 0) { ret -= 64; continue; }
    ret -= __builtin_clzll(a[i]);
    break;
  }
  return ret;
}

/* ---------------------------------------------------------------- Fq, fq.rs:796-1123 */
static inline int fq_is_zero(const fq *a) { return repr_is_zero(a->l, 6); }
static inline int fq_eq(const fq *a, const fq *b) { return memcmp(a, b, sizeof(fq)) == 0; }
static inline int fq_is_valid(const fq *a) { return repr_cmp(a->l, BLS_MODULUS, 6) < 0; } /* fq.rs:1023 */
static inline void fq_reduce(fq *a) { if (!fq_is_valid(a)) repr_sub_noborrow(a->l, BLS_MODULUS, 6); } /* fq.rs:1030 */

static inline void fq_add(fq *a, const fq *b) { repr_add_nocarry(a->l, b->l, 6); fq_reduce(a); } /* fq.rs:813 */
static inline void fq_dbl(fq *a) { repr_mul2(a->l, 6); fq_reduce(a); }                         /* fq.rs:822 */
static inline void fq_sub(fq *a, const fq *b) {                                                /* fq.rs:831 */
  if (repr_cmp(b->l, a->l, 6) > 0) repr_add_nocarry(a->l, BLS_MODULUS, 6);
  repr_sub_noborrow(a->l, b->l, 6);
}
static inline void fq_neg(fq *a) {                                                             /* fq.rs:841 */
  if (!fq_is_zero(a)) {
    fq t = *MOD();
    repr_sub_noborrow(t.l, a->l, 6);
    *a = t;
  }
}

/* HAC 14.32 as written at fq.rs:1037-1122 */
static inline void fq_mont_reduce(fq *out, uint64_t r[12]) {
  uint64_t carry2 = 0;
  for (int i = 0; i < 6; i++) {
    uint64_t k = r[i] * BLS_INV64, carry = 0;
    (void)mac(r[i], k, BLS_MODULUS[0], &carry);
    for (int j = 1; j < 6; j++) r[i + j] = mac(r[i + j], k, BLS_MODULUS[j], &carry);
    r[i + 6] = adc(r[i + 6], carry2, &carry);
    carry2 = carry;
  }
  for (int i = 0; i < 6; i++) out->l[i] = r[6 + i];
  fq_reduce(out);
}

static inline void fq_mul(fq *a, const fq *b) { /* fq.rs:910-960 */
  uint64_t r[12] = {0};
  for (int i = 0; i < 6; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 6; j++) r[i + j] = mac(r[i + j], a->l[i], b->l[j], &carry);
    r[i + 6] = carry;
  }
  fq_mont_reduce(a, r);
}

static inline void fq_sqr(fq *a) { /* fq.rs:963-1016 */
  uint64_t r[12] = {0};
  for (int i = 0; i < 5; i++) {
    uint64_t carry = 0;
    for (int j = i + 1; j < 6; j++) r[i + j] = mac(r[i + j], a->l[i], a->l[j], &carry);
    r[i + 6] = carry;
  }
  r[11] = r[10] >> 63;
  for (int i = 10; i >= 2; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
  r[1] <<= 1;
  uint64_t carry = 0;
  for (int i = 0; i < 6; i++) {
    r[2 * i] = mac(r[2 * i], a->l[i], a->l[i], &carry);
    r[2 * i + 1] = adc(r[2 * i + 1], 0, &carry);
  }
  fq_mont_reduce(a, r);
}

/* Binary extended Euclid, fq.rs:849-902.  Returns 0 for a == 0 (None). */
static int fq_inv(fq *out, const fq *a) {
  if (fq_is_zero(a)) return 0;
  static const uint64_t one[6] = {1, 0, 0, 0, 0, 0};
  fq u = *a, v = *MOD(), b = *(const fq *)BLS_R2, c = {{0}};
  while (repr_cmp(u.l, one, 6) != 0 && repr_cmp(v.l, one, 6) != 0) {
    while ((u.l[0] & 1) == 0) {
      repr_div2(u.l, 6);
      if (b.l[0] & 1) repr_add_nocarry(b.l, BLS_MODULUS, 6);
      repr_div2(b.l, 6);
    }
    while ((v.l[0] & 1) == 0) {
      repr_div2(v.l, 6);
      if (c.l[0] & 1) repr_add_nocarry(c.l, BLS_MODULUS, 6);
      repr_div2(c.l, 6);
    }
    if (repr_cmp(v.l, u.l, 6) < 0) {
      repr_sub_noborrow(u.l, v.l, 6);
      fq_sub(&b, &c);
    } else {
      repr_sub_noborrow(v.l, u.l, 6);
      fq_sub(&c, &b);
    }
  }
  *out = (repr_cmp(u.l, one, 6) == 0) ? b : c;
  return 1;
}

static int fq_from_repr(fq *out, const fq *r) { /* fq.rs:747-756 */
  if (!fq_is_valid(r)) return 0;
  *out = *r;
  fq_mul(out, (const fq *)BLS_R2);
  return 1;
}
static void fq_into_repr(fq *out, const fq *a) { /* fq.rs:758-777 */
  uint64_t r[12] = {0};
  memcpy(r, a->l, 48);
  fq_mont_reduce(out, r);
}

/* ---------------------------------------------------------------- Fq2, fq2.rs:39-160 */
static inline int fq2_is_zero(const fq2 *a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static inline int fq2_eq(const fq2 *a, const fq2 *b) { return memcmp(a, b, sizeof(fq2)) == 0; }
static inline void fq2_add(fq2 *a, const fq2 *b) { fq_add(&a->c0, &b->c0); fq_add(&a->c1, &b->c1); }
static inline void fq2_sub(fq2 *a, const fq2 *b) { fq_sub(&a->c0, &b->c0); fq_sub(&a->c1, &b->c1); }
static inline void fq2_dbl(fq2 *a) { fq_dbl(&a->c0); fq_dbl(&a->c1); }
static inline void fq2_neg(fq2 *a) { fq_neg(&a->c0); fq_neg(&a->c1); }
static inline void fq2_mul_by_nonresidue(fq2 *a) { /* fq2.rs:41-45 */
  fq t0 = a->c0;
  fq_sub(&a->c0, &a->c1);
  fq_add(&a->c1, &t0);
}
static inline void fq2_sqr(fq2 *a) { /* fq2.rs:87-101 */
  fq ab = a->c0; fq_mul(&ab, &a->c1);
  fq c0c1 = a->c0; fq_add(&c0c1, &a->c1);
  fq c0 = a->c1; fq_neg(&c0); fq_add(&c0, &a->c0);
  fq_mul(&c0, &c0c1); fq_sub(&c0, &ab);
  a->c1 = ab; fq_add(&a->c1, &ab);
  fq_add(&c0, &ab);
  a->c0 = c0;
}
static inline void fq2_mul(fq2 *a, const fq2 *b) { /* fq2.rs:123-136 */
  fq aa = a->c0; fq_mul(&aa, &b->c0);
  fq bb = a->c1; fq_mul(&bb, &b->c1);
  fq o = b->c0; fq_add(&o, &b->c1);
  fq_add(&a->c1, &a->c0);
  fq_mul(&a->c1, &o);
  fq_sub(&a->c1, &aa);
  fq_sub(&a->c1, &bb);
  a->c0 = aa;
  fq_sub(&a->c0, &bb);
}
static int fq2_inv(fq2 *out, const fq2 *a) { /* fq2.rs:138-155 */
  fq t1 = a->c1; fq_sqr(&t1);
  fq t0 = a->c0; fq_sqr(&t0);
  fq_add(&t0, &t1);
  fq t;
  if (!fq_inv(&t, &t0)) return 0;
  fq2 tmp = *a;
  fq_mul(&tmp.c0, &t);
  fq_mul(&tmp.c1, &t);
  fq_neg(&tmp.c1);
  *out = tmp;
  return 1;
}
static inline void fq2_frobenius(fq2 *a, unsigned power) { /* fq2.rs:157-159 */
  fq_mul(&a->c1, (const fq *)BLS_FROB_FQ2_C1[power % 2]);
}

/* ---------------------------------------------------------------- Fq6, fq6.rs:30-302 */
static inline int fq6_is_zero(const fq6 *a) { return fq2_is_zero(&a->c0) && fq2_is_zero(&a->c1) && fq2_is_zero(&a->c2); }
static inline void fq6_add(fq6 *a, const fq6 *b) { fq2_add(&a->c0, &b->c0); fq2_add(&a->c1, &b->c1); fq2_add(&a->c2, &b->c2); }
static inline void fq6_sub(fq6 *a, const fq6 *b) { fq2_sub(&a->c0, &b->c0); fq2_sub(&a->c1, &b->c1); fq2_sub(&a->c2, &b->c2); }
static inline void fq6_neg(fq6 *a) { fq2_neg(&a->c0); fq2_neg(&a->c1); fq2_neg(&a->c2); }
static inline void fq6_mul_by_nonresidue(fq6 *a) { /* fq6.rs:32-38 */
  fq2 t = a->c0; a->c0 = a->c1; a->c1 = t;
  t = a->c0; a->c0 = a->c2; a->c2 = t;
  fq2_mul_by_nonresidue(&a->c0);
}
static void fq6_mul_by_1(fq6 *a, const fq2 *c1) { /* fq6.rs:40-66 */
  fq2 b_b = a->c1; fq2_mul(&b_b, c1);
  fq2 t1 = *c1;
  { fq2 tmp = a->c1; fq2_add(&tmp, &a->c2); fq2_mul(&t1, &tmp); fq2_sub(&t1, &b_b); fq2_mul_by_nonresidue(&t1); }
  fq2 t2 = *c1;
  { fq2 tmp = a->c0; fq2_add(&tmp, &a->c1); fq2_mul(&t2, &tmp); fq2_sub(&t2, &b_b); }
  a->c0 = t1; a->c1 = t2; a->c2 = b_b;
}
static void fq6_mul_by_01(fq6 *a, const fq2 *c0, const fq2 *c1) { /* fq6.rs:68-109 */
  fq2 a_a = a->c0, b_b = a->c1;
  fq2_mul(&a_a, c0); fq2_mul(&b_b, c1);
  fq2 t1 = *c1;
  { fq2 tmp = a->c1; fq2_add(&tmp, &a->c2); fq2_mul(&t1, &tmp); fq2_sub(&t1, &b_b); fq2_mul_by_nonresidue(&t1); fq2_add(&t1, &a_a); }
  fq2 t3 = *c0;
  { fq2 tmp = a->c0; fq2_add(&tmp, &a->c2); fq2_mul(&t3, &tmp); fq2_sub(&t3, &a_a); fq2_add(&t3, &b_b); }
  fq2 t2 = *c0; fq2_add(&t2, c1);
  { fq2 tmp = a->c0; fq2_add(&tmp, &a->c1); fq2_mul(&t2, &tmp); fq2_sub(&t2, &a_a); fq2_sub(&t2, &b_b); }
  a->c0 = t1; a->c1 = t2; a->c2 = t3;
}
static void fq6_frobenius(fq6 *a, unsigned power) { /* fq6.rs:157-164 */
  fq2_frobenius(&a->c0, power); fq2_frobenius(&a->c1, power); fq2_frobenius(&a->c2, power);
  fq2_mul(&a->c1, (const fq2 *)BLS_FROB_FQ6_C1[power % 6]);
  fq2_mul(&a->c2, (const fq2 *)BLS_FROB_FQ6_C2[power % 6]);
}
static void fq6_sqr(fq6 *a) { /* fq6.rs:166-197 */
  fq2 s0 = a->c0; fq2_sqr(&s0);
  fq2 ab = a->c0; fq2_mul(&ab, &a->c1);
  fq2 s1 = ab; fq2_dbl(&s1);
  fq2 s2 = a->c0; fq2_sub(&s2, &a->c1); fq2_add(&s2, &a->c2); fq2_sqr(&s2);
  fq2 bc = a->c1; fq2_mul(&bc, &a->c2);
  fq2 s3 = bc; fq2_dbl(&s3);
  fq2 s4 = a->c2; fq2_sqr(&s4);
  a->c0 = s3; fq2_mul_by_nonresidue(&a->c0); fq2_add(&a->c0, &s0);
  a->c1 = s4; fq2_mul_by_nonresidue(&a->c1); fq2_add(&a->c1, &s1);
  a->c2 = s1; fq2_add(&a->c2, &s2); fq2_add(&a->c2, &s3); fq2_sub(&a->c2, &s0); fq2_sub(&a->c2, &s4);
}
static void fq6_mul(fq6 *a, const fq6 *b) { /* fq6.rs:199-248 */
  fq2 a_a = a->c0, b_b = a->c1, c_c = a->c2;
  fq2_mul(&a_a, &b->c0); fq2_mul(&b_b, &b->c1); fq2_mul(&c_c, &b->c2);
  fq2 t1 = b->c1; fq2_add(&t1, &b->c2);
  { fq2 tmp = a->c1; fq2_add(&tmp, &a->c2); fq2_mul(&t1, &tmp); fq2_sub(&t1, &b_b); fq2_sub(&t1, &c_c); fq2_mul_by_nonresidue(&t1); fq2_add(&t1, &a_a); }
  fq2 t3 = b->c0; fq2_add(&t3, &b->c2);
  { fq2 tmp = a->c0; fq2_add(&tmp, &a->c2); fq2_mul(&t3, &tmp); fq2_sub(&t3, &a_a); fq2_add(&t3, &b_b); fq2_sub(&t3, &c_c); }
  fq2 t2 = b->c0; fq2_add(&t2, &b->c1);
  { fq2 tmp = a->c0; fq2_add(&tmp, &a->c1); fq2_mul(&t2, &tmp); fq2_sub(&t2, &a_a); fq2_sub(&t2, &b_b); fq2_mul_by_nonresidue(&c_c); fq2_add(&t2, &c_c); }
  a->c0 = t1; a->c1 = t2; a->c2 = t3;
}
static int fq6_inv(fq6 *out, const fq6 *a) { /* fq6.rs:250-301 */
  fq2 c0 = a->c2; fq2_mul_by_nonresidue(&c0); fq2_mul(&c0, &a->c1); fq2_neg(&c0);
  { fq2 c0s = a->c0; fq2_sqr(&c0s); fq2_add(&c0, &c0s); }
  fq2 c1 = a->c2; fq2_sqr(&c1); fq2_mul_by_nonresidue(&c1);
  { fq2 c01 = a->c0; fq2_mul(&c01, &a->c1); fq2_sub(&c1, &c01); }
  fq2 c2 = a->c1; fq2_sqr(&c2);
  { fq2 c02 = a->c0; fq2_mul(&c02, &a->c2); fq2_sub(&c2, &c02); }
  fq2 tmp1 = a->c2; fq2_mul(&tmp1, &c1);
  fq2 tmp2 = a->c1; fq2_mul(&tmp2, &c2);
  fq2_add(&tmp1, &tmp2); fq2_mul_by_nonresidue(&tmp1);
  tmp2 = a->c0; fq2_mul(&tmp2, &c0); fq2_add(&tmp1, &tmp2);
  fq2 t;
  if (!fq2_inv(&t, &tmp1)) return 0;
  fq6 r = {t, t, t};
  fq2_mul(&r.c0, &c0); fq2_mul(&r.c1, &c1); fq2_mul(&r.c2, &c2);
  *out = r;
  return 1;
}

/* ---------------------------------------------------------------- Fq12, fq12.rs:29-149 */
static inline int fq12_is_zero(const fq12 *a) { return fq6_is_zero(&a->c0) && fq6_is_zero(&a->c1); }
static void fq12_one(fq12 *a) { memset(a, 0, sizeof(*a)); a->c0.c0.c0 = *ONE(); }
static inline void fq12_conjugate(fq12 *a) { fq6_neg(&a->c1); } /* fq12.rs:30-32 */
static void fq12_mul_by_014(fq12 *a, const fq2 *c0, const fq2 *c1, const fq2 *c4) { /* fq12.rs:34-48 */
  fq6 aa = a->c0; fq6_mul_by_01(&aa, c0, c1);
  fq6 bb = a->c1; fq6_mul_by_1(&bb, c4);
  fq2 o = *c1; fq2_add(&o, c4);
  fq6_add(&a->c1, &a->c0);
  fq6_mul_by_01(&a->c1, c0, &o);
  fq6_sub(&a->c1, &aa);
  fq6_sub(&a->c1, &bb);
  a->c0 = bb;
  fq6_mul_by_nonresidue(&a->c0);
  fq6_add(&a->c0, &aa);
}
static void fq12_frobenius(fq12 *a, unsigned power) { /* fq12.rs:90-97 */
  fq6_frobenius(&a->c0, power); fq6_frobenius(&a->c1, power);
  const fq2 *k = (const fq2 *)BLS_FROB_FQ12_C1[power % 12];
  fq2_mul(&a->c1.c0, k); fq2_mul(&a->c1.c1, k); fq2_mul(&a->c1.c2, k);
}
static void fq12_sqr(fq12 *a) { /* fq12.rs:99-114 */
  fq6 ab = a->c0; fq6_mul(&ab, &a->c1);
  fq6 c0c1 = a->c0; fq6_add(&c0c1, &a->c1);
  fq6 c0 = a->c1; fq6_mul_by_nonresidue(&c0); fq6_add(&c0, &a->c0);
  fq6_mul(&c0, &c0c1); fq6_sub(&c0, &ab);
  a->c1 = ab; fq6_add(&a->c1, &ab);
  fq6_mul_by_nonresidue(&ab);
  fq6_sub(&c0, &ab);
  a->c0 = c0;
}
static void fq12_mul(fq12 *a, const fq12 *b) { /* fq12.rs:116-130 */
  fq6 aa = a->c0; fq6_mul(&aa, &b->c0);
  fq6 bb = a->c1; fq6_mul(&bb, &b->c1);
  fq6 o = b->c0; fq6_add(&o, &b->c1);
  fq6_add(&a->c1, &a->c0);
  fq6_mul(&a->c1, &o);
  fq6_sub(&a->c1, &aa);
  fq6_sub(&a->c1, &bb);
  a->c0 = bb;
  fq6_mul_by_nonresidue(&a->c0);
  fq6_add(&a->c0, &aa);
}
static int fq12_inv(fq12 *out, const fq12 *a) { /* fq12.rs:132-148 */
  fq6 c0s = a->c0; fq6_sqr(&c0s);
  fq6 c1s = a->c1; fq6_sqr(&c1s);
  fq6_mul_by_nonresidue(&c1s);
  fq6_sub(&c0s, &c1s);
  fq6 t;
  if (!fq6_inv(&t, &c0s)) return 0;
  fq12 tmp = {t, t};
  fq6_mul(&tmp.c0, &a->c0);
  fq6_mul(&tmp.c1, &a->c1);
  fq6_neg(&tmp.c1);
  *out = tmp;
  return 1;
}
/* Field::pow over one u64 limb, lib.rs:306-324 (BitIterator MSB first, lib.rs:583-610) */
static void fq12_pow_u64(fq12 *out, const fq12 *a, uint64_t e) {
  fq12 res; fq12_one(&res);
  int found_one = 0;
  for (int n = 63; n >= 0; n--) {
    int bit = (e >> n) & 1;
    if (found_one) fq12_sqr(&res); else found_one = bit;
    if (bit) fq12_mul(&res, a);
  }
  *out = res;
}

/* ---------------------------------------------------------------- curve groups: ec.rs:216-619 via an
 * include-twice template, the C analogue of the reference's curve_impl! macro. */
#define CURVE_G g1
#define CURVE_A g1_affine
#define CURVE_F fq
#define FN(n) g1_##n
#define F_(n) fq_##n
#define F_ONE (*ONE())
#include "curve_impl.inc"
#undef CURVE_G
#undef CURVE_A
#undef CURVE_F
#undef FN
#undef F_
#undef F_ONE

static const fq2 *FQ2_ONE_P(void) { static fq2 o; static int init = 0; if (!init) { memset(&o, 0, sizeof o); o.c0 = *ONE(); init = 1; } return &o; }
#define CURVE_G g2
#define CURVE_A g2_affine
#define CURVE_F fq2
#define FN(n) g2_##n
#define F_(n) fq2_##n
#define F_ONE (*FQ2_ONE_P())
#include "curve_impl.inc"

/* window heuristics: ec.rs:895-905 (G1), 1586-1596 (G2) */
static int g1_window_for_scalar(const fr_repr *k) { int nb = repr_num_bits(k->l, 4); return nb >= 130 ? 4 : (nb >= 34 ? 3 : 2); }
static int g2_window_for_scalar(const fr_repr *k) { int nb = repr_num_bits(k->l, 4); return nb >= 103 ? 4 : (nb >= 37 ? 3 : 2); }

/* wnaf_form, wnaf.rs:18-43.  Returns the digit count (<= 257). */
static int wnaf_form(int64_t *out, fr_repr c, int window) {
  int n = 0;
  while (!repr_is_zero(c.l, 4)) {
    int64_t u;
    if (c.l[0] & 1) {
      u = (int64_t)(c.l[0] % (1ull << (window + 1)));
      if (u > (1ll << window)) u -= 1ll << (window + 1);
      uint64_t t[4] = {0, 0, 0, 0};
      if (u > 0) { t[0] = (uint64_t)u; repr_sub_noborrow(c.l, t, 4); }
      else       { t[0] = (uint64_t)(-u); repr_add_nocarry(c.l, t, 4); }
    } else {
      u = 0;
    }
    out[n++] = u;
    repr_div2(c.l, 4);
  }
  return n;
}

/* ---------------------------------------------------------------- engine: bls12_381/mod.rs:40-358 */
static void doubling_step(g2 *r, fq2 out[3]) { /* mod.rs:176-245 */
  fq2 tmp0 = r->x; fq2_sqr(&tmp0);
  fq2 tmp1 = r->y; fq2_sqr(&tmp1);
  fq2 tmp2 = tmp1; fq2_sqr(&tmp2);
  fq2 tmp3 = tmp1; fq2_add(&tmp3, &r->x); fq2_sqr(&tmp3); fq2_sub(&tmp3, &tmp0); fq2_sub(&tmp3, &tmp2); fq2_dbl(&tmp3);
  fq2 tmp4 = tmp0; fq2_dbl(&tmp4); fq2_add(&tmp4, &tmp0);
  fq2 tmp6 = r->x; fq2_add(&tmp6, &tmp4);
  fq2 tmp5 = tmp4; fq2_sqr(&tmp5);
  fq2 zsquared = r->z; fq2_sqr(&zsquared);
  r->x = tmp5; fq2_sub(&r->x, &tmp3); fq2_sub(&r->x, &tmp3);
  fq2_add(&r->z, &r->y); fq2_sqr(&r->z); fq2_sub(&r->z, &tmp1); fq2_sub(&r->z, &zsquared);
  r->y = tmp3; fq2_sub(&r->y, &r->x); fq2_mul(&r->y, &tmp4);
  fq2_dbl(&tmp2); fq2_dbl(&tmp2); fq2_dbl(&tmp2);
  fq2_sub(&r->y, &tmp2);
  tmp3 = tmp4; fq2_mul(&tmp3, &zsquared); fq2_dbl(&tmp3); fq2_neg(&tmp3);
  fq2_sqr(&tmp6); fq2_sub(&tmp6, &tmp0); fq2_sub(&tmp6, &tmp5);
  fq2_dbl(&tmp1); fq2_dbl(&tmp1);
  fq2_sub(&tmp6, &tmp1);
  tmp0 = r->z; fq2_mul(&tmp0, &zsquared); fq2_dbl(&tmp0);
  out[0] = tmp0; out[1] = tmp3; out[2] = tmp6;
}
static void addition_step(g2 *r, const g2_affine *q, fq2 out[3]) { /* mod.rs:247-333 */
  fq2 zsquared = r->z; fq2_sqr(&zsquared);
  fq2 ysquared = q->y; fq2_sqr(&ysquared);
  fq2 t0 = zsquared; fq2_mul(&t0, &q->x);
  fq2 t1 = q->y; fq2_add(&t1, &r->z); fq2_sqr(&t1); fq2_sub(&t1, &ysquared); fq2_sub(&t1, &zsquared); fq2_mul(&t1, &zsquared);
  fq2 t2 = t0; fq2_sub(&t2, &r->x);
  fq2 t3 = t2; fq2_sqr(&t3);
  fq2 t4 = t3; fq2_dbl(&t4); fq2_dbl(&t4);
  fq2 t5 = t4; fq2_mul(&t5, &t2);
  fq2 t6 = t1; fq2_sub(&t6, &r->y); fq2_sub(&t6, &r->y);
  fq2 t9 = t6; fq2_mul(&t9, &q->x);
  fq2 t7 = t4; fq2_mul(&t7, &r->x);
  r->x = t6; fq2_sqr(&r->x); fq2_sub(&r->x, &t5); fq2_sub(&r->x, &t7); fq2_sub(&r->x, &t7);
  fq2_add(&r->z, &t2); fq2_sqr(&r->z); fq2_sub(&r->z, &zsquared); fq2_sub(&r->z, &t3);
  fq2 t10 = q->y; fq2_add(&t10, &r->z);
  fq2 t8 = t7; fq2_sub(&t8, &r->x); fq2_mul(&t8, &t6);
  t0 = r->y; fq2_mul(&t0, &t5); fq2_dbl(&t0);
  r->y = t8; fq2_sub(&r->y, &t0);
  fq2_sqr(&t10); fq2_sub(&t10, &ysquared);
  fq2 ztsquared = r->z; fq2_sqr(&ztsquared);
  fq2_sub(&t10, &ztsquared);
  fq2_dbl(&t9); fq2_sub(&t9, &t10);
  t10 = r->z; fq2_dbl(&t10);
  fq2_neg(&t6);
  t1 = t6; fq2_dbl(&t1);
  out[0] = t10; out[1] = t1; out[2] = t9;
}
static void g2_prepare(g2_prepared *out, const g2_affine *q) { /* mod.rs:168-358 */
  memset(out, 0, sizeof(*out));
  if (q->inf) { out->inf = 1; return; }
  int n = 0;
  g2 r; g2_from_affine(&r, q);
  int found_one = 0;
  for (int b = 63; b >= 0; b--) {
    int bit = ((BLS_X_ABS >> 1) >> b) & 1;
    if (!found_one) { found_one = bit; continue; }
    doubling_step(&r, out->c[n++]);
    if (bit) addition_step(&r, q, out->c[n++]);
  }
  doubling_step(&r, out->c[n++]);
}
static void ell(fq12 *f, const fq2 coeffs[3], const g1_affine *p) { /* mod.rs:57-69 */
  fq2 c0 = coeffs[0], c1 = coeffs[1];
  fq_mul(&c0.c0, &p->y); fq_mul(&c0.c1, &p->y);
  fq_mul(&c1.c0, &p->x); fq_mul(&c1.c1, &p->x);
  fq12_mul_by_014(f, &coeffs[2], &c1, &c0);
}
/* miller_loop over n (G1Affine, G2Prepared) pairs sharing one accumulator, mod.rs:40-102 */
static void miller_loop(fq12 *out, const g1_affine *const *ps, const g2_prepared *const *qs, size_t n) {
  size_t live = 0;
  const g1_affine **lp = (const g1_affine **)malloc(sizeof(void *) * (n ? n : 1));
  const g2_prepared **lq = (const g2_prepared **)malloc(sizeof(void *) * (n ? n : 1));
  for (size_t i = 0; i < n; i++)
    if (!ps[i]->inf && !qs[i]->inf) { lp[live] = ps[i]; lq[live] = qs[i]; live++; }
  fq12 f; fq12_one(&f);
  int idx = 0, found_one = 0;
  for (int b = 63; b >= 0; b--) {
    int bit = ((BLS_X_ABS >> 1) >> b) & 1;
    if (!found_one) { found_one = bit; continue; }
    for (size_t i = 0; i < live; i++) ell(&f, lq[i]->c[idx], lp[i]);
    idx++;
    if (bit) { for (size_t i = 0; i < live; i++) ell(&f, lq[i]->c[idx], lp[i]); idx++; }
    fq12_sqr(&f);
  }
  for (size_t i = 0; i < live; i++) ell(&f, lq[i]->c[idx], lp[i]);
  fq12_conjugate(&f); /* BLS_X_IS_NEGATIVE */
  *out = f;
  free(lp); free(lq);
}
static void exp_by_x(fq12 *f, uint64_t x) { /* mod.rs:116-121 */
  fq12 t; fq12_pow_u64(&t, f, x); fq12_conjugate(&t); *f = t;
}
static int final_exponentiation(fq12 *out, const fq12 *rin) { /* mod.rs:104-160 */
  fq12 f1 = *rin; fq12_conjugate(&f1);
  fq12 f2;
  if (!fq12_inv(&f2, rin)) return 0;
  fq12 r = f1; fq12_mul(&r, &f2);
  f2 = r;
  fq12_frobenius(&r, 2);
  fq12_mul(&r, &f2);
  uint64_t x = BLS_X_ABS;
  fq12 y0 = r; fq12_sqr(&y0);
  fq12 y1 = y0; exp_by_x(&y1, x);
  x >>= 1;
  fq12 y2 = y1; exp_by_x(&y2, x);
  x <<= 1;
  fq12 y3 = r; fq12_conjugate(&y3);
  fq12_mul(&y1, &y3);
  fq12_conjugate(&y1);
  fq12_mul(&y1, &y2);
  y2 = y1; exp_by_x(&y2, x);
  y3 = y2; exp_by_x(&y3, x);
  fq12_conjugate(&y1);
  fq12_mul(&y3, &y1);
  fq12_conjugate(&y1);
  fq12_frobenius(&y1, 3);
  fq12_frobenius(&y2, 2);
  fq12_mul(&y1, &y2);
  y2 = y3; exp_by_x(&y2, x);
  fq12_mul(&y2, &y0);
  fq12_mul(&y2, &r);
  fq12_mul(&y1, &y2);
  y2 = y3; fq12_frobenius(&y2, 1);
  fq12_mul(&y1, &y2);
  *out = y1;
  return 1;
}

/* ================================================================ exported C API (ctypes) ======= */
#define API __attribute__((visibility("default")))

typedef void (*range_fn)(void *ctx, size_t lo, size_t hi);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->lo, j->hi); return NULL; }
/* static contiguous partition of [0,n) over `threads` pthreads (the rayon-over-the-batch baseline) */
static void parallel_for(size_t n, int threads, range_fn fn, void *ctx) {
  if (threads <= 1 || n < 2) { fn(ctx, 0, n); return; }
  if ((size_t)threads > n) threads = (int)n;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t].fn = fn; jobs[t].ctx = ctx;
    jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* ---- field-level batch entry points (op codes shared with the CUDA test kernels) ---- */
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_SQR = 3, OP_NEG = 4, OP_DBL = 5, OP_INV = 6,
       OP_FROM_REPR = 7, OP_INTO_REPR = 8, OP_MUL_NONRES = 9, OP_FROB1 = 10, OP_FROB2 = 11, OP_FROB3 = 12,
       OP_CONJ = 13, OP_MUL_BY_014 = 14, OP_MUL_BY_01 = 15, OP_MUL_BY_1 = 16 };

/* a, b, out: n elements each (b may be NULL for unary ops); ok[n] gets 0 where the result is None */
API int oracle_fq_op(int op, const fq *a, const fq *b, fq *out, uint8_t *ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fq r = a[i]; int good = 1;
    switch (op) {
      case OP_ADD: fq_add(&r, &b[i]); break;
      case OP_SUB: fq_sub(&r, &b[i]); break;
      case OP_MUL: fq_mul(&r, &b[i]); break;
      case OP_SQR: fq_sqr(&r); break;
      case OP_NEG: fq_neg(&r); break;
      case OP_DBL: fq_dbl(&r); break;
      case OP_INV: good = fq_inv(&r, &a[i]); if (!good) memset(&r, 0, sizeof r); break;
      case OP_FROM_REPR: good = fq_from_repr(&r, &a[i]); if (!good) memset(&r, 0, sizeof r); break;
      case OP_INTO_REPR: fq_into_repr(&r, &a[i]); break;
      default: return -1;
    }
    out[i] = r; if (ok) ok[i] = (uint8_t)good;
  }
  return 0;
}
API int oracle_fq2_op(int op, const fq2 *a, const fq2 *b, fq2 *out, uint8_t *ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fq2 r = a[i]; int good = 1;
    switch (op) {
      case OP_ADD: fq2_add(&r, &b[i]); break;
      case OP_SUB: fq2_sub(&r, &b[i]); break;
      case OP_MUL: fq2_mul(&r, &b[i]); break;
      case OP_SQR: fq2_sqr(&r); break;
      case OP_NEG: fq2_neg(&r); break;
      case OP_DBL: fq2_dbl(&r); break;
      case OP_INV: good = fq2_inv(&r, &a[i]); if (!good) memset(&r, 0, sizeof r); break;
      case OP_MUL_NONRES: fq2_mul_by_nonresidue(&r); break;
      case OP_FROB1: fq2_frobenius(&r, 1); break;
      default: return -1;
    }
    out[i] = r; if (ok) ok[i] = (uint8_t)good;
  }
  return 0;
}
/* b is fq6 for binary ops; for MUL_BY_01 b[i].c0,c1 are the sparse operands, for MUL_BY_1 b[i].c1 */
API int oracle_fq6_op(int op, const fq6 *a, const fq6 *b, fq6 *out, uint8_t *ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fq6 r = a[i]; int good = 1;
    switch (op) {
      case OP_ADD: fq6_add(&r, &b[i]); break;
      case OP_SUB: fq6_sub(&r, &b[i]); break;
      case OP_MUL: fq6_mul(&r, &b[i]); break;
      case OP_SQR: fq6_sqr(&r); break;
      case OP_NEG: fq6_neg(&r); break;
      case OP_INV: good = fq6_inv(&r, &a[i]); if (!good) memset(&r, 0, sizeof r); break;
      case OP_MUL_NONRES: fq6_mul_by_nonresidue(&r); break;
      case OP_FROB1: fq6_frobenius(&r, 1); break;
      case OP_FROB2: fq6_frobenius(&r, 2); break;
      case OP_FROB3: fq6_frobenius(&r, 3); break;
      case OP_MUL_BY_01: fq6_mul_by_01(&r, &b[i].c0, &b[i].c1); break;
      case OP_MUL_BY_1: fq6_mul_by_1(&r, &b[i].c1); break;
      default: return -1;
    }
    out[i] = r; if (ok) ok[i] = (uint8_t)good;
  }
  return 0;
}
/* for MUL_BY_014 the sparse operands are b[i].c0.c0 (c0), b[i].c0.c1 (c1), b[i].c1.c1 (c4) */
API int oracle_fq12_op(int op, const fq12 *a, const fq12 *b, fq12 *out, uint8_t *ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fq12 r = a[i]; int good = 1;
    switch (op) {
      case OP_MUL: fq12_mul(&r, &b[i]); break;
      case OP_SQR: fq12_sqr(&r); break;
      case OP_INV: good = fq12_inv(&r, &a[i]); if (!good) memset(&r, 0, sizeof r); break;
      case OP_CONJ: fq12_conjugate(&r); break;
      case OP_FROB1: fq12_frobenius(&r, 1); break;
      case OP_FROB2: fq12_frobenius(&r, 2); break;
      case OP_FROB3: fq12_frobenius(&r, 3); break;
      case OP_MUL_BY_014: fq12_mul_by_014(&r, &b[i].c0.c0, &b[i].c0.c1, &b[i].c1.c1); break;
      default: return -1;
    }
    out[i] = r; if (ok) ok[i] = (uint8_t)good;
  }
  return 0;
}
API int oracle_fq12_pow_u64(const fq12 *a, uint64_t e, fq12 *out, size_t n) {
  for (size_t i = 0; i < n; i++) fq12_pow_u64(&out[i], &a[i], e);
  return 0;
}

/* ---- engine ---- */
typedef struct { const g1_affine *p; const g2_affine *q; const g2_prepared *qp; const fq12 *fin; fq12 *out;
                 g2_prepared *prep; uint8_t *ok; } eng_ctx;

static void prep_range(void *c, size_t lo, size_t hi) { eng_ctx *e = (eng_ctx *)c; for (size_t i = lo; i < hi; i++) g2_prepare(&e->prep[i], &e->q[i]); }
API int oracle_g2_prepare(const g2_affine *q, g2_prepared *out, size_t n, int threads) {
  eng_ctx c = {0}; c.q = q; c.prep = out; parallel_for(n, threads, prep_range, &c); return 0;
}
static void miller_range(void *c, size_t lo, size_t hi) {
  eng_ctx *e = (eng_ctx *)c;
  g2_prepared *tmp = (g2_prepared *)malloc(sizeof(g2_prepared));
  for (size_t i = lo; i < hi; i++) {
    const g1_affine *pp = &e->p[i];
    const g2_prepared *qq;
    if (e->qp) qq = &e->qp[i]; else { g2_prepare(tmp, &e->q[i]); qq = tmp; }
    miller_loop(&e->out[i], &pp, &qq, 1);
  }
  free(tmp);
}
/* n independent single-pair Miller loops (G1Affine::prepare + G2Affine::prepare + miller_loop) */
API int oracle_miller_loop(const g1_affine *p, const g2_affine *q, fq12 *out, size_t n, int threads) {
  eng_ctx c = {0}; c.p = p; c.q = q; c.out = out; parallel_for(n, threads, miller_range, &c); return 0;
}
/* same, from already-prepared G2 coefficients */
API int oracle_miller_loop_prepared(const g1_affine *p, const g2_prepared *qp, fq12 *out, size_t n, int threads) {
  eng_ctx c = {0}; c.p = p; c.qp = qp; c.out = out; parallel_for(n, threads, miller_range, &c); return 0;
}
/* ONE Miller loop over n pairs sharing the accumulator, exactly Engine::miller_loop(&[...]) */
API int oracle_multi_miller_loop(const g1_affine *p, const g2_affine *q, size_t n, fq12 *out) {
  g2_prepared *preps = (g2_prepared *)malloc(sizeof(g2_prepared) * (n ? n : 1));
  const g1_affine **ps = (const g1_affine **)malloc(sizeof(void *) * (n ? n : 1));
  const g2_prepared **qs = (const g2_prepared **)malloc(sizeof(void *) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) { g2_prepare(&preps[i], &q[i]); ps[i] = &p[i]; qs[i] = &preps[i]; }
  miller_loop(out, ps, qs, n);
  free(preps); free(ps); free(qs);
  return 0;
}
/* product over i of the single-pair Miller values, threads-parallel: equal to the above as a field value */
typedef struct { const g1_affine *p; const g2_affine *q; fq12 *partial; size_t n; int threads; } mm_ctx;
static void mm_range(void *c, size_t lo, size_t hi) {
  mm_ctx *m = (mm_ctx *)c;
  /* which thread am I: derive from lo */
  size_t t = 0; while (m->n * (t + 1) / m->threads <= lo && t + 1 < (size_t)m->threads) t++;
  fq12 acc; fq12_one(&acc);
  g2_prepared *tmp = (g2_prepared *)malloc(sizeof(g2_prepared));
  for (size_t i = lo; i < hi; i++) {
    const g1_affine *pp = &m->p[i]; const g2_prepared *qq = tmp; fq12 f;
    g2_prepare(tmp, &m->q[i]);
    miller_loop(&f, &pp, &qq, 1);
    fq12_mul(&acc, &f);
  }
  free(tmp);
  m->partial[t] = acc;
}
API int oracle_multi_miller_product(const g1_affine *p, const g2_affine *q, size_t n, fq12 *out, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  fq12 *partial = (fq12 *)malloc(sizeof(fq12) * threads);
  for (int t = 0; t < threads; t++) fq12_one(&partial[t]);
  mm_ctx m = {p, q, partial, n, threads};
  if (threads == 1) mm_range(&m, 0, n); else parallel_for(n, threads, mm_range, &m);
  fq12 acc; fq12_one(&acc);
  for (int t = 0; t < threads; t++) fq12_mul(&acc, &partial[t]);
  *out = acc;
  free(partial);
  return 0;
}
static void fe_range(void *c, size_t lo, size_t hi) {
  eng_ctx *e = (eng_ctx *)c;
  for (size_t i = lo; i < hi; i++) {
    int good = final_exponentiation(&e->out[i], &e->fin[i]);
    if (!good) memset(&e->out[i], 0, sizeof(fq12));
    if (e->ok) e->ok[i] = (uint8_t)good;
  }
}
API int oracle_final_exponentiation(const fq12 *in, fq12 *out, uint8_t *is_some, size_t n, int threads) {
  eng_ctx c = {0}; c.fin = in; c.out = out; c.ok = is_some; parallel_for(n, threads, fe_range, &c); return 0;
}
static void pairing_range(void *c, size_t lo, size_t hi) {
  eng_ctx *e = (eng_ctx *)c;
  g2_prepared *tmp = (g2_prepared *)malloc(sizeof(g2_prepared));
  for (size_t i = lo; i < hi; i++) {
    const g1_affine *pp = &e->p[i]; const g2_prepared *qq = tmp; fq12 f;
    g2_prepare(tmp, &e->q[i]);
    miller_loop(&f, &pp, &qq, 1);
    final_exponentiation(&e->out[i], &f); /* .unwrap(): Miller values of valid points are non-zero */
  }
  free(tmp);
}
/* Engine::pairing on affine inputs, lib.rs:101-109 */
API int oracle_pairing(const g1_affine *p, const g2_affine *q, fq12 *out, size_t n, int threads) {
  eng_ctx c = {0}; c.p = p; c.q = q; c.out = out; parallel_for(n, threads, pairing_range, &c); return 0;
}

/* ---- curve ops ---- */
enum { PT_DOUBLE = 0, PT_ADD = 1, PT_ADD_MIXED = 2, PT_NEGATE = 3, PT_MUL = 4, PT_WNAF = 5, PT_SUB = 6 };
typedef struct { int op; int window; const void *a; const void *b; const fr_repr *k; void *out; } pt_ctx;

static void g1_range(void *c, size_t lo, size_t hi) {
  pt_ctx *p = (pt_ctx *)c;
  const g1 *a = (const g1 *)p->a; g1 *out = (g1 *)p->out;
  for (size_t i = lo; i < hi; i++) {
    g1 r = a[i];
    switch (p->op) {
      case PT_DOUBLE: g1_double(&r); break;
      case PT_ADD: g1_add_assign(&r, &((const g1 *)p->b)[i]); break;
      case PT_SUB: g1_sub_assign(&r, &((const g1 *)p->b)[i]); break;
      case PT_ADD_MIXED: g1_add_assign_mixed(&r, &((const g1_affine *)p->b)[i]); break;
      case PT_NEGATE: g1_negate(&r); break;
      case PT_MUL: g1_mul_assign(&r, &p->k[i]); break;
      case PT_WNAF: g1_wnaf_mul(&r, &a[i], &p->k[i], p->window ? p->window : g1_window_for_scalar(&p->k[i])); break;
    }
    out[i] = r;
  }
}
API int oracle_g1_op(int op, const g1 *a, const void *b, const fr_repr *k, g1 *out, size_t n, int window, int threads) {
  pt_ctx c = {op, window, a, b, k, out}; parallel_for(n, threads, g1_range, &c); return 0;
}
static void g2_range(void *c, size_t lo, size_t hi) {
  pt_ctx *p = (pt_ctx *)c;
  const g2 *a = (const g2 *)p->a; g2 *out = (g2 *)p->out;
  for (size_t i = lo; i < hi; i++) {
    g2 r = a[i];
    switch (p->op) {
      case PT_DOUBLE: g2_double(&r); break;
      case PT_ADD: g2_add_assign(&r, &((const g2 *)p->b)[i]); break;
      case PT_SUB: g2_sub_assign(&r, &((const g2 *)p->b)[i]); break;
      case PT_ADD_MIXED: g2_add_assign_mixed(&r, &((const g2_affine *)p->b)[i]); break;
      case PT_NEGATE: g2_negate(&r); break;
      case PT_MUL: g2_mul_assign(&r, &p->k[i]); break;
      case PT_WNAF: g2_wnaf_mul(&r, &a[i], &p->k[i], p->window ? p->window : g2_window_for_scalar(&p->k[i])); break;
    }
    out[i] = r;
  }
}
API int oracle_g2_op(int op, const g2 *a, const void *b, const fr_repr *k, g2 *out, size_t n, int window, int threads) {
  pt_ctx c = {op, window, a, b, k, out}; parallel_for(n, threads, g2_range, &c); return 0;
}
/* CurveProjective::batch_normalization over the WHOLE slice, sequential as in the reference */
API int oracle_g1_batch_normalization(g1 *v, size_t n) { g1_batch_normalization(v, n); return 0; }
API int oracle_g2_batch_normalization(g2 *v, size_t n) { g2_batch_normalization(v, n); return 0; }
/* per-point From<projective> for affine */
API int oracle_g1_into_affine(const g1 *v, g1_affine *out, size_t n) { for (size_t i = 0; i < n; i++) g1_into_affine(&out[i], &v[i]); return 0; }
API int oracle_g2_into_affine(const g2 *v, g2_affine *out, size_t n) { for (size_t i = 0; i < n; i++) g2_into_affine(&out[i], &v[i]); return 0; }
API int oracle_g1_from_affine(const g1_affine *v, g1 *out, size_t n) { for (size_t i = 0; i < n; i++) g1_from_affine(&out[i], &v[i]); return 0; }
API int oracle_g2_from_affine(const g2_affine *v, g2 *out, size_t n) { for (size_t i = 0; i < n; i++) g2_from_affine(&out[i], &v[i]); return 0; }
/* wnaf_form digits for one scalar; returns the digit count */
API int oracle_wnaf_form(const fr_repr *k, int window, int64_t *digits) { return wnaf_form(digits, *k, window); }
API int oracle_g1_window_for_scalar(const fr_repr *k) { return g1_window_for_scalar(k); }
API int oracle_g2_window_for_scalar(const fr_repr *k) { return g2_window_for_scalar(k); }
API void oracle_generators(g1_affine *g1o, g2_affine *g2o) {
  memset(g1o, 0, sizeof *g1o); memset(g2o, 0, sizeof *g2o);
  g1o->x = *(const fq *)BLS_G1_X; g1o->y = *(const fq *)BLS_G1_Y;
  g2o->x.c0 = *(const fq *)BLS_G2_X0; g2o->x.c1 = *(const fq *)BLS_G2_X1;
  g2o->y.c0 = *(const fq *)BLS_G2_Y0; g2o->y.c1 = *(const fq *)BLS_G2_Y1;
}
API size_t oracle_sizeof(int what) {
  switch (what) { case 0: return sizeof(fq); case 1: return sizeof(fq2); case 2: return sizeof(fq6); case 3: return sizeof(fq12);
    case 4: return sizeof(g1_affine); case 5: return sizeof(g1); case 6: return sizeof(g2_affine); case 7: return sizeof(g2);
    case 8: return sizeof(fr_repr); case 9: return sizeof(g2_prepared); }
  return 0;
}

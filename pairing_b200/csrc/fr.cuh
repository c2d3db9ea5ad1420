// fr.cuh -- the scalar field Fr of BLS12-381 (255-bit r) in Montgomery form, 8 x 32-bit limbs (bls12_381/fr.rs:324-572).
// The step AFTER the path: callers combine scalars (c * d in the reference's bilinearity test, tests/engine.rs:117-119).
// Not a hot path: plain word-serial CIOS on 64-bit temporaries, results canonical (< r) after every operation.
#pragma once
#include <stdint.h>

namespace bls {

struct Fr { uint32_t v[8]; };

// r (fr.rs:5-12), R = 2^256 mod r (fr.rs:20-26), R^2 mod r (fr.rs:28-34), -r^-1 mod 2^32 (low half of INV, fr.rs:36)
__device__ __forceinline__ Fr fr_modulus() { return Fr{{0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}}; }
__device__ __forceinline__ Fr fr_one() { return Fr{{0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}}; }
__device__ __forceinline__ Fr fr_r2() { return Fr{{0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}}; }
#define BLS_FR_NINV 0xffffffffu
__device__ __forceinline__ Fr fr_zero() { return Fr{{0, 0, 0, 0, 0, 0, 0, 0}}; }

__device__ __forceinline__ bool fr_is_zero(const Fr& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}
__device__ __forceinline__ bool fr_eq(const Fr& a, const Fr& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}
// a >= b as 256-bit integers
__device__ __forceinline__ bool fr_geq(const Fr& a, const Fr& b) {
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] != b.v[i]) return a.v[i] > b.v[i];
  }
  return true;
}
__device__ __forceinline__ Fr fr_raw_sub(const Fr& a, const Fr& b) {
  Fr r; uint64_t borrow = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow; r.v[i] = (uint32_t)t; borrow = (t >> 32) & 1; }
  return r;
}
// fr.rs:513-518 reduce
__device__ __forceinline__ Fr fr_reduce(const Fr& a) { return fr_geq(a, fr_modulus()) ? fr_raw_sub(a, fr_modulus()) : a; }
// fr.rs:341-348 (a + b < 2r < 2^256: no carry out)
__device__ __forceinline__ Fr fr_add(const Fr& a, const Fr& b) {
  Fr r; uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)a.v[i] + b.v[i] + c; r.v[i] = (uint32_t)t; c = t >> 32; }
  return fr_reduce(r);
}
// fr.rs:359-367
__device__ __forceinline__ Fr fr_sub(const Fr& a, const Fr& b) {
  if (fr_geq(a, b)) return fr_raw_sub(a, b);
  Fr t; uint64_t c = 0;
  const Fr m = fr_modulus();
#pragma unroll
  for (int i = 0; i < 8; i++) { uint64_t s = (uint64_t)a.v[i] + m.v[i] + c; t.v[i] = (uint32_t)s; c = s >> 32; }
  return fr_raw_sub(t, b);      // (a + r) - b; a + r < 2^256
}
// fr.rs:369-375
__device__ __forceinline__ Fr fr_neg(const Fr& a) { return fr_is_zero(a) ? a : fr_raw_sub(fr_modulus(), a); }
// fr.rs:438-465 + 520-572: a * b * 2^-256 mod r
__device__ __noinline__ Fr fr_mul(Fr a, Fr b) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
  const Fr m = fr_modulus();
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c; t[j] = (uint32_t)s; c = s >> 32; }
    uint64_t s = (uint64_t)t[8] + c; t[8] = (uint32_t)s; t[9] = (uint32_t)(s >> 32);
    const uint32_t k = t[0] * BLS_FR_NINV;
    c = ((uint64_t)k * m.v[0] + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) { uint64_t u = (uint64_t)k * m.v[j] + t[j] + c; t[j - 1] = (uint32_t)u; c = u >> 32; }
    s = (uint64_t)t[8] + c; t[7] = (uint32_t)s; t[8] = t[9] + (uint32_t)(s >> 32);
  }
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  // t < 2r: one conditional subtraction (t[8] is 0 here because 2r < 2^256)
  return fr_reduce(r);
}
__device__ __forceinline__ Fr fr_sqr(const Fr& a) { return fr_mul(a, a); }
// fr.rs:377-431 computes the inverse by binary extended Euclid; a^(r-2) is the same field value.  false for zero.
__device__ __noinline__ bool fr_inv(Fr& out, const Fr& a) {
  if (fr_is_zero(a)) { out = fr_zero(); return false; }
  const Fr e = fr_raw_sub(fr_modulus(), Fr{{2, 0, 0, 0, 0, 0, 0, 0}});   // r - 2
  Fr res = fr_one();
  bool started = false;
#pragma unroll 1
  for (int i = 254; i >= 0; i--) {
    if (started) res = fr_sqr(res);
    if ((e.v[i >> 5] >> (i & 31)) & 1u) { res = started ? fr_mul(res, a) : a; started = true; }
  }
  out = res;
  return true;
}

}  // namespace bls

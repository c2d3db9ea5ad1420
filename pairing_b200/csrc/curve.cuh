// curve.cuh -- G1 (over Fq) and G2 (over Fq2) in Jacobian coordinates, wNAF scalar multiplication.
// One template over the coordinate field, like the reference's `curve_impl!` macro
// (bls12_381/ec.rs:1-621).  Jacobian (X, Y, Z) triples are representative-dependent, so the formula
// set, the special cases and the wNAF op sequence follow the reference exactly:
//   double            dbl-2009-l   ec.rs:296-354
//   add_assign        add-2007-bl  ec.rs:356-444 (copy when self = inf, no-op when other = inf,
//                                  double when equal, H = 0 falls through when P + (-P))
//   add_assign_mixed  madd-2007-bl ec.rs:446-526
//   wnaf_table / wnaf_form / wnaf_exp   wnaf.rs:4-71
#pragma once
#include "tower.cuh"

namespace bls {

// field-generic spellings
__device__ __forceinline__ Fp f_add(const Fp& a, const Fp& b) { return fp_add(a, b); }
__device__ __forceinline__ Fp f_sub(const Fp& a, const Fp& b) { return fp_sub(a, b); }
__device__ __forceinline__ Fp f_dbl(const Fp& a) { return fp_dbl(a); }
__device__ __forceinline__ Fp f_neg(const Fp& a) { return fp_neg(a); }
__device__ __forceinline__ Fp f_mul(const Fp& a, const Fp& b) { return fp_mul(a, b); }
__device__ __forceinline__ Fp f_sqr(const Fp& a) { return fp_sqr(a); }
__device__ __forceinline__ bool f_is_zero(const Fp& a) { return fp_is_zero(a); }
__device__ __forceinline__ bool f_eq(const Fp& a, const Fp& b) { return fp_eq(a, b); }
__device__ __forceinline__ bool f_inv(Fp& r, const Fp& a) { return fp_inv(r, a); }
__device__ __forceinline__ void f_set_one(Fp& a) { a = fp_one(); }
__device__ __forceinline__ void f_set_zero(Fp& a) { a = fp_zero(); }

__device__ __forceinline__ Fp2 f_add(const Fp2& a, const Fp2& b) { return fp2_add(a, b); }
__device__ __forceinline__ Fp2 f_sub(const Fp2& a, const Fp2& b) { return fp2_sub(a, b); }
__device__ __forceinline__ Fp2 f_dbl(const Fp2& a) { return fp2_dbl(a); }
__device__ __forceinline__ Fp2 f_neg(const Fp2& a) { return fp2_neg(a); }
__device__ __forceinline__ Fp2 f_mul(const Fp2& a, const Fp2& b) { return fp2_mul(a, b); }
__device__ __forceinline__ Fp2 f_sqr(const Fp2& a) { return fp2_sqr(a); }
__device__ __forceinline__ bool f_is_zero(const Fp2& a) { return fp2_is_zero(a); }
__device__ __forceinline__ bool f_eq(const Fp2& a, const Fp2& b) { return fp2_eq(a, b); }
__device__ __forceinline__ bool f_inv(Fp2& r, const Fp2& a) { return fp2_inv(r, a); }
__device__ __forceinline__ void f_set_one(Fp2& a) { a = fp2_one(); }
__device__ __forceinline__ void f_set_zero(Fp2& a) { a = fp2_zero(); }

template <class F> struct Jac { F x, y, z; };
template <class F> struct Aff { F x, y; bool inf; };

template <class F> __device__ __forceinline__ bool pt_is_zero(const Jac<F>& p) { return f_is_zero(p.z); }
// ec.rs:224-230: (0, 1, 0)
template <class F> __device__ __forceinline__ void pt_set_zero(Jac<F>& p) { f_set_zero(p.x); f_set_one(p.y); f_set_zero(p.z); }
template <class F> __device__ __forceinline__ bool pt_is_normalized(const Jac<F>& p) {
  F one; f_set_one(one);
  return pt_is_zero(p) || f_eq(p.z, one);
}

template <class F> __device__ __noinline__ void pt_double(Jac<F>& s) {
  if (pt_is_zero(s)) return;
  F a = f_sqr(s.x);
  F b = f_sqr(s.y);
  F c = f_sqr(b);
  F d = f_dbl(f_sub(f_sub(f_sqr(f_add(s.x, b)), a), c));
  F e = f_add(f_dbl(a), a);
  F f = f_sqr(e);
  s.z = f_dbl(f_mul(s.z, s.y));
  s.x = f_sub(f_sub(f, d), d);
  c = f_dbl(f_dbl(f_dbl(c)));
  s.y = f_sub(f_mul(f_sub(d, s.x), e), c);
}

// Returns true when the operands are equal and the caller has to double instead (ec.rs:394-396);
// the doubling is issued by the inline wrapper below, not from inside this function.
template <class F> __device__ __noinline__ bool pt_add_core(Jac<F>& s, const Jac<F>& o) {
  if (pt_is_zero(s)) { s = o; return false; }
  if (pt_is_zero(o)) return false;
  F z1z1 = f_sqr(s.z);
  F z2z2 = f_sqr(o.z);
  F u1 = f_mul(s.x, z2z2);
  F u2 = f_mul(o.x, z1z1);
  F s1 = f_mul(f_mul(s.y, o.z), z2z2);
  F s2 = f_mul(f_mul(o.y, s.z), z1z1);
  F h = f_sub(u2, u1);
  F sd = f_sub(s2, s1);
  if (f_is_zero(h) && f_is_zero(sd)) return true;      // u1 == u2 && s1 == s2: the caller doubles (ec.rs:394-396)
  F i = f_sqr(f_dbl(h));
  F j = f_mul(h, i);
  F r = f_dbl(sd);
  F v = f_mul(u1, i);
  s.x = f_sub(f_sub(f_sub(f_sqr(r), j), v), v);
  s.y = f_sub(f_mul(f_sub(v, s.x), r), f_dbl(f_mul(s1, j)));
  s.z = f_mul(f_sub(f_sub(f_sqr(f_add(s.z, o.z)), z1z1), z2z2), h);
  return false;
}
template <class F> __device__ __forceinline__ void pt_add(Jac<F>& s, const Jac<F>& o) {
  if (pt_add_core(s, o)) pt_double(s);
}

template <class F> __device__ __noinline__ bool pt_add_mixed_core(Jac<F>& s, const Aff<F>& o) {
  if (o.inf) return false;
  if (pt_is_zero(s)) { s.x = o.x; s.y = o.y; f_set_one(s.z); return false; }
  F z1z1 = f_sqr(s.z);
  F u2 = f_mul(o.x, z1z1);
  F s2 = f_mul(f_mul(o.y, s.z), z1z1);
  F h = f_sub(u2, s.x);
  F sd = f_sub(s2, s.y);
  if (f_is_zero(h) && f_is_zero(sd)) return true;      // same point: the caller doubles (ec.rs:471-473)
  F hh = f_sqr(h);
  F i = f_dbl(f_dbl(hh));
  F j = f_mul(h, i);
  F r = f_dbl(sd);
  F v = f_mul(s.x, i);
  F x3 = f_sub(f_sub(f_sub(f_sqr(r), j), v), v);
  F y3 = f_sub(f_mul(f_sub(v, x3), r), f_dbl(f_mul(j, s.y)));
  s.z = f_sub(f_sub(f_sqr(f_add(s.z, h)), z1z1), hh);
  s.x = x3; s.y = y3;
  return false;
}
template <class F> __device__ __forceinline__ void pt_add_mixed(Jac<F>& s, const Aff<F>& o) {
  if (pt_add_mixed_core(s, o)) pt_double(s);
}

// ec.rs:528-532
template <class F> __device__ __forceinline__ void pt_negate(Jac<F>& s) { if (!pt_is_zero(s)) s.y = f_neg(s.y); }

// ec.rs:586-619 From<projective> for affine
template <class F> __device__ __noinline__ void pt_into_affine(Aff<F>& out, const Jac<F>& p) {
  F one; f_set_one(one);
  if (pt_is_zero(p)) { f_set_zero(out.x); out.y = one; out.inf = true; return; }   // ec.rs:158-164
  out.inf = false;
  if (f_eq(p.z, one)) { out.x = p.x; out.y = p.y; return; }
  F zinv; f_inv(zinv, p.z);
  F zp = f_sqr(zinv);
  out.x = f_mul(p.x, zp);
  out.y = f_mul(p.y, f_mul(zp, zinv));
}

// ---- scalars (FrRepr, fr.rs:57-244): canonical 256-bit integers, 8 x u32 (struct Scalar, fp.cuh) ----

__device__ __forceinline__ int scalar_num_bits(const Scalar& k) {   // fr.rs:213-225
  for (int i = 7; i >= 0; i--)
    if (k.v[i]) return 32 * i + 32 - __clz(k.v[i]);
  return 0;
}
// ec.rs:895-905 / 1586-1596
__device__ __forceinline__ int g1_window_for_bits(int nb) { return nb >= 130 ? 4 : (nb >= 34 ? 3 : 2); }
__device__ __forceinline__ int g2_window_for_bits(int nb) { return nb >= 103 ? 4 : (nb >= 37 ? 3 : 2); }

#define BLS_MAX_WNAF_WINDOW 7          /* per-point tables in per-thread local memory */
#define BLS_MAX_WNAF_EXPLICIT_WINDOW 13 /* per-point tables in a global-memory scratch row (the reference's test sweep, tests/curve.rs:78) */
#define BLS_MAX_WNAF_FIXED_WINDOW 16   /* shared table: ec.rs:907-921 picks up to 16 (G1) / 15 (G2) */
#define BLS_MAX_WNAF_TABLE (1 << (BLS_MAX_WNAF_WINDOW - 1))

// wnaf_form, wnaf.rs:18-43.  Digits are odd in (-2^w, 2^w): int8_t holds them for w <= 7, int32_t for w <= 16.
template <class D> __device__ __forceinline__ int wnaf_form(D* digits, Scalar c, int window) {
  int n = 0;
  const uint32_t mask = (2u << window) - 1u;
  while (true) {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= c.v[i];
    if (!nz) break;
    int u = 0;
    if (c.v[0] & 1u) {
      u = (int)(c.v[0] & mask);
      if (u > (1 << window)) u -= (2 << window);
      // c -= u  (u may be negative): add the sign-extended negation with carry
      uint32_t lo = (uint32_t)(-u);
      uint32_t ext = (u > 0) ? 0xffffffffu : 0u;
      uint64_t t = (uint64_t)c.v[0] + lo;
      c.v[0] = (uint32_t)t;
#pragma unroll
      for (int i = 1; i < 8; i++) {
        t = (uint64_t)c.v[i] + ext + (t >> 32);
        c.v[i] = (uint32_t)t;
      }
    }
    digits[n++] = (D)u;
#pragma unroll
    for (int i = 0; i < 7; i++) c.v[i] = (c.v[i] >> 1) | (c.v[i + 1] << 31);
    c.v[7] >>= 1;
  }
  return n;
}

// wnaf_exp with the warp's lanes DECOUPLED: every lane walks its own (double, add) sequence -- for digit i
// from the top: a double once a non-zero digit has been seen, then an add/sub if digit i is non-zero -- but
// the warp only ever executes one of the two group operations at a time.  Executing both at every digit
// position (the SIMT reading of wnaf_exp) leaves 5/6 of the lanes idle in every addition (density of
// non-zero digits for w = 4: 1/6), i.e. 42 % lane efficiency.  Here lanes that have reached an addition wait
// until enough of their neighbours have too.  Each lane still performs exactly the reference's operation
// sequence on its own point, so the Jacobian triple is the reference's bit for bit.  No shuffles inside:
// divergence is safe for F = Fp / Fp2.
// K points per lane: a lane that has reached an addition on one of its points keeps doubling another one,
// so it is almost never idle (tools/wnaf_sched_sim.py: 76 % lane efficiency for K = 1, 91 % for K = 2,
// 95 % for K = 3).  Per point the operation sequence is still exactly the reference's.
template <int K> struct WnafState {
  int i[K];            // current digit position of point j (-1: finished)
  bool found[K];       // a non-zero digit has been consumed (wnaf.rs:52, found_one)
  bool doubled[K];     // the double of position i[j] has been done
};
// `Tab::get(j, e)` returns entry e of the table of point j (per-point tables in local memory, or ONE shared
// table in global memory for the fixed-base mode Wnaf::base(g, n).scalar(s), wnaf.rs:93-107, 169-178).
template <class F, int K, class Tab, class D> __device__ __forceinline__ void pt_wnaf_run_lazy(Jac<F> (&res)[K], const Tab& table, D (&digits)[K][260], WnafState<K>& st) {
#pragma unroll 1
  while (true) {
    // next group operation of every point of this lane: 0 = finished, 1 = double, 2 = add/sub of digit i
    int op[K], dig[K];
    bool has_add = false, has_dbl = false;
#pragma unroll
    for (int j = 0; j < K; j++) {
      op[j] = 0; dig[j] = 0;
#pragma unroll 1
      while (st.i[j] >= 0) {
        if (st.found[j] && !st.doubled[j]) { op[j] = 1; break; }
        dig[j] = digits[j][st.i[j]];
        if (dig[j] != 0) { op[j] = 2; break; }
        st.i[j]--; st.doubled[j] = false;
      }
      has_add |= op[j] == 2; has_dbl |= op[j] == 1;
    }
    const unsigned m_add = __ballot_sync(0xffffffffu, has_add), m_dbl = __ballot_sync(0xffffffffu, has_dbl);
    if ((m_add | m_dbl) == 0) break;
    const bool do_add = m_dbl == 0 || (K == 1 ? 3 * __popc(m_add) >= 10 * __popc(m_dbl) : 4 * __popc(m_add) >= 5 * __popc(m_dbl));
    // the point of this lane that wants the chosen operation and is furthest behind
    int pick = -1, best = -1;
#pragma unroll
    for (int j = 0; j < K; j++)
      if (op[j] == (do_add ? 2 : 1) && st.i[j] > best) { best = st.i[j]; pick = j; }
    if (pick >= 0) {
      if (do_add) {
        const int n = dig[pick];
        Jac<F> t = table.get(pick, (n < 0 ? -n : n) >> 1);
        if (n < 0) pt_negate(t);             // sub_assign: copy, negate, add (lib.rs:156-160)
        pt_add(res[pick], t);
        st.found[pick] = true; st.doubled[pick] = false; st.i[pick]--;
      } else {
        pt_double(res[pick]);
        st.doubled[pick] = true;
      }
    }
  }
}

// ---- warp-cooperative addition of a FIXED point, for the window table of the fixed-base mode (wnaf.rs:4-15) ----
// table[i+1] = table[i] + 2g is a chain of 2^(w-1) DEPENDENT additions (each entry's Jacobian representative depends on the
// previous one), so it cannot be spread over points; but inside one add-2007-bl there are up to three independent
// products per level.  Lanes of one warp hold the same point; lane r (mod 3) computes product r of each level and the
// results are broadcast by shuffles: 6 multiply levels instead of 16 sequential products.  z2^2 and z2^3 of the fixed
// operand are computed once.  Same field values as pt_add_core, hence the same canonical triple.
__device__ __forceinline__ Fp f_sel(bool a, const Fp& x, const Fp& y) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = a ? x.v[i] : y.v[i];
  return r;
}
__device__ __forceinline__ Fp2 f_sel(bool a, const Fp2& x, const Fp2& y) { return Fp2{f_sel(a, x.c0, y.c0), f_sel(a, x.c1, y.c1)}; }
__device__ __forceinline__ Fp f_bcast(const Fp& x, int src) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = __shfl_sync(0xffffffffu, x.v[i], src);
  return r;
}
__device__ __forceinline__ Fp2 f_bcast(const Fp2& x, int src) { return Fp2{f_bcast(x.c0, src), f_bcast(x.c1, src)}; }
template <class F> __device__ __forceinline__ F f_sel3(int role, const F& a, const F& b, const F& c) { return f_sel(role == 0, a, f_sel(role == 1, b, c)); }

// s += o for a fixed o with z2z2 = o.z^2 and z2c = o.z^3; every lane of the warp holds the same s and o.
template <class F> __device__ __forceinline__ void pt_add_fixed_coop(Jac<F>& s, const Jac<F>& o, const F& z2z2, const F& z2c) {
  if (pt_is_zero(s) || pt_is_zero(o)) { pt_add(s, o); return; }          // warp-uniform: all lanes hold the same values
  const int role = (threadIdx.x & 31) % 3;
  F m = f_mul(f_sel3(role, s.z, s.x, s.y), f_sel3(role, s.z, z2z2, z2c));
  const F z1z1 = f_bcast(m, 0), u1 = f_bcast(m, 1), s1 = f_bcast(m, 2);
  const F zs = f_add(s.z, o.z);
  m = f_mul(f_sel3(role, o.x, s.z, zs), f_sel3(role, z1z1, z1z1, zs));
  const F u2 = f_bcast(m, 0), z1c = f_bcast(m, 1), zz = f_bcast(m, 2);
  const F h = f_sub(u2, u1), h2 = f_dbl(h);
  m = f_mul(f_sel(role == 0, o.y, h2), f_sel(role == 0, z1c, h2));
  const F s2 = f_bcast(m, 0), i = f_bcast(m, 1);
  const F sd = f_sub(s2, s1);
  if (f_is_zero(h) && f_is_zero(sd)) { pt_double(s); return; }           // equal operands (ec.rs:394-396), warp-uniform
  m = f_mul(f_sel3(role, h, u1, f_sub(f_sub(zz, z1z1), z2z2)), f_sel3(role, i, i, h));
  const F j = f_bcast(m, 0), v = f_bcast(m, 1), z3 = f_bcast(m, 2);
  const F r = f_dbl(sd);
  m = f_mul(f_sel(role == 0, r, s1), f_sel(role == 0, r, j));
  const F r2 = f_bcast(m, 0), s1j = f_bcast(m, 1);
  s.x = f_sub(f_sub(f_sub(r2, j), v), v);
  s.y = f_sub(f_mul(f_sub(v, s.x), r), f_dbl(s1j));
  s.z = z3;
}

// double-and-add, ec.rs:534-553
template <class F> __device__ __forceinline__ void pt_mul(Jac<F>& s, const Scalar& k) {
  Jac<F> res;
  pt_set_zero(res);
  bool found_one = false;
#pragma unroll 1
  for (int n = 255; n >= 0; n--) {
    bool bit = (k.v[n >> 5] >> (n & 31)) & 1u;
    if (found_one) pt_double(res); else found_one = bit;
    if (bit) pt_add(res, s);
  }
  s = res;
}

}  // namespace bls

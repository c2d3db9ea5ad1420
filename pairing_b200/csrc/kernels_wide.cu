// kernels_wide.cu -- the latency path: ONE WARP per element (wide.cuh runs the micro-programs of wide_prog_gen.cuh).
//   k_wide_pairing      Engine::pairing for small batches (BASELINE configs[0]: benches/bls12_381/mod.rs:91-107)
//   k_wide_final_exp    Engine::final_exponentiation for small batches (mod.rs:104-160)
//   k_wide_miller       Engine::miller_loop for small batches and for the Miller values of a small multi-pairing product
//   k_pair_product_tail the tail of a multi-pairing product: the product of the per-block / per-device partial Miller
//                       values on lane pairs, then -- optionally -- ONE final exponentiation by warp 0
// Its own translation unit (see abi_common.cuh): nothing here can perturb ptxas' allocation of the throughput kernels.
#include "abi_common.cuh"
#include "wide.cuh"

#define WIDE_SLOT_BYTES (WIDE_SLOT_WORDS * 4)

// one warp per block: blocks spread over all SMs and sub-partitions even for a handful of elements
__global__ void __launch_bounds__(32) k_wide_final_exp(const uint64_t* in, uint64_t* out, uint8_t* is_some, size_t n) {
  extern __shared__ __align__(16) uint32_t wide_slots[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const int lane = threadIdx.x & 31, lp = lane >> 1, c = lane & 1;
  bool zero = true;
  if (lp < 6) {                                       // input k = Fq2 coefficient k of f, in slots 0..5
    const Fp v = ld_fp(in + FQ12_W * i + 12 * lp + 6 * c);
    zero = fp_is_zero(v);
    wide_st(wide_slots, lp, c, v);
  }
  const bool all_zero = __all_sync(0xffffffffu, zero);   // mod.rs:107-108: None for f == 0
  wide_load_consts(wide_slots, WIDE_FINAL_EXP_CONST, WIDE_FINAL_EXP_NCONST);
  __syncwarp();
  wide_run(WIDE_FINAL_EXP_CODE, WIDE_FINAL_EXP_NROUNDS, wide_slots, WIDE_FINAL_EXP_NSLOTS);
  if (lp < 6) {
    Fp v = wide_ld(wide_slots, WIDE_FINAL_EXP_OUT[lp], c);
    if (all_zero) v = fp_zero();
    st_fp(out + FQ12_W * i + 12 * lp + 6 * c, v);
  }
  if (lane == 0 && is_some) is_some[i] = !all_zero;
}

// inputs of the PAIRING / MILLER programs: px, py (as Fq2 values with a zero u-part), qx, qy in slots 0..3
__global__ void __launch_bounds__(32) k_wide_pairing(const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n) {
  extern __shared__ __align__(16) uint32_t wide_slots[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const int lane = threadIdx.x & 31, lp = lane >> 1, c = lane & 1;
  const uint64_t* pi = p + G1A_W * i;
  const uint64_t* qi = q + G2A_W * i;
  const bool live = pi[12] == 0 && qi[24] == 0;
  if (lp < 2) wide_st(wide_slots, lp, c, c ? fp_zero() : ld_fp(pi + 6 * lp));
  else if (lp < 4) wide_st(wide_slots, lp, c, ld_fp(qi + 12 * (lp - 2) + 6 * c));
  wide_load_consts(wide_slots, WIDE_PAIRING_CONST, WIDE_PAIRING_NCONST);
  __syncwarp();
  wide_run(WIDE_PAIRING_CODE, WIDE_PAIRING_NROUNDS, wide_slots, WIDE_PAIRING_NSLOTS);
  if (lp < 6) {
    Fp v = wide_ld(wide_slots, WIDE_PAIRING_OUT[lp], c);
    if (!live) v = (lp == 0 && c == 0) ? fp_one() : fp_zero();     // mod.rs:49-54: the pair is skipped, e = final_exponentiation(one) = one
    st_fp(out + FQ12_W * i + 12 * lp + 6 * c, v);
  }
}

// Engine::miller_loop for one pair per warp (mod.rs:40-102): the MILLER program, i.e. the PAIRING program without the final
// exponentiation.  Small batches of bls_miller_loop_dev, and the Miller values of a SMALL multi-pairing product (the signature-
// verification shape: a product of two or a few pairings), which k_pair_product_tail then folds and exponentiates.
__global__ void __launch_bounds__(32) k_wide_miller(const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n) {
  extern __shared__ __align__(16) uint32_t wide_slots[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const int lane = threadIdx.x & 31, lp = lane >> 1, c = lane & 1;
  const uint64_t* pi = p + G1A_W * i;
  const uint64_t* qi = q + G2A_W * i;
  const bool live = pi[12] == 0 && qi[24] == 0;
  if (lp < 2) wide_st(wide_slots, lp, c, c ? fp_zero() : ld_fp(pi + 6 * lp));
  else if (lp < 4) wide_st(wide_slots, lp, c, ld_fp(qi + 12 * (lp - 2) + 6 * c));
  wide_load_consts(wide_slots, WIDE_MILLER_CONST, WIDE_MILLER_NCONST);
  __syncwarp();
  wide_run(WIDE_MILLER_CODE, WIDE_MILLER_NROUNDS, wide_slots, WIDE_MILLER_NSLOTS);
  if (lp < 6) {
    Fp v = wide_ld(wide_slots, WIDE_MILLER_OUT[lp], c);
    if (!live) v = (lp == 0 && c == 0) ? fp_one() : fp_zero();     // mod.rs:49-54: the pair is skipped, f stays one
    st_fp(out + FQ12_W * i + 12 * lp + 6 * c, v);
  }
}

// Engine::pairing on PROJECTIVE inputs (the crate's bench_pairing_full: `Bls12::pairing(G1, G2)` converts both points with
// into_affine first, lib.rs:101-109 + ec.rs:586-619): the TO_AFFINE program (two inversions, eight products) runs in front of
// the PAIRING program on the same slot file.
__global__ void __launch_bounds__(32) k_wide_pairing_projective(const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n) {
  extern __shared__ __align__(16) uint32_t wide_slots[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const int lane = threadIdx.x & 31, lp = lane >> 1, c = lane & 1;
  const uint64_t* pi = p + G1_W * i;
  const uint64_t* qi = q + G2_W * i;
  bool zinf = false;                                   // z == 0: the point at infinity (ec.rs:238-240)
  if (lp < 3) {                                        // pX, pY, pZ in slots 0..2 (zero u-part)
    const Fp v = ld_fp(pi + 6 * lp);
    if (lp == 2) zinf = fp_is_zero(v);
    wide_st(wide_slots, lp, c, c ? fp_zero() : v);
  } else if (lp < 6) {                                 // qX, qY, qZ in slots 3..5
    const Fp v = ld_fp(qi + 12 * (lp - 3) + 6 * c);
    if (lp == 5) zinf = fp_is_zero(v);
    wide_st(wide_slots, lp, c, v);
  }
  // P is infinity iff lane 4 (pZ, c = 0) saw zero; Q iff both lanes of pair 5 did
  const unsigned z = __ballot_sync(0xffffffffu, zinf);
  const bool live = !((z >> 4) & 1u) && !(((z >> 10) & 3u) == 3u);
  __syncwarp();
  wide_run(WIDE_TO_AFFINE_CODE, WIDE_TO_AFFINE_NROUNDS, wide_slots, WIDE_TO_AFFINE_NSLOTS);
  Fp aff = fp_zero();
  if (lp < 4) aff = wide_ld(wide_slots, WIDE_TO_AFFINE_OUT[lp], c);      // px, py, qx, qy
  __syncwarp();
  if (lp < 4) wide_st(wide_slots, lp, c, aff);                           // the PAIRING program's inputs: slots 0..3
  wide_load_consts(wide_slots, WIDE_PAIRING_CONST, WIDE_PAIRING_NCONST);
  __syncwarp();
  wide_run(WIDE_PAIRING_CODE, WIDE_PAIRING_NROUNDS, wide_slots, WIDE_PAIRING_NSLOTS);
  if (lp < 6) {
    Fp v = wide_ld(wide_slots, WIDE_PAIRING_OUT[lp], c);
    if (!live) v = (lp == 0 && c == 0) ? fp_one() : fp_zero();
    st_fp(out + FQ12_W * i + 12 * lp + 6 * c, v);
  }
}

// Product of `count` Fq12 values on lane pairs by ONE block, then (final_exp != 0) the final exponentiation of the
// product by warp 0 on the wide engine.  Lane pair l multiplies values l, l + 64, ...; the 64 partial products are
// folded by a shared-memory tree.  Every lane of a warp executes every product (full-mask shuffles inside p2_mul):
// out-of-range factors are replaced by one.
#define TAIL_TPB 128
#define TAIL_LP (TAIL_TPB / 2)
__device__ __forceinline__ void tail_st(uint32_t* sm, int lp, int c, const P12& f) {
  const Fp* v = reinterpret_cast<const Fp*>(&f);
#pragma unroll
  for (int k = 0; k < 6; k++) wide_st(sm, lp * 6 + k, c, v[k]);
}
__device__ __forceinline__ void tail_ld(P12& f, const uint32_t* sm, int lp, int c) {
  Fp* v = reinterpret_cast<Fp*>(&f);
#pragma unroll
  for (int k = 0; k < 6; k++) v[k] = wide_ld(sm, lp * 6 + k, c);
}
__global__ void __launch_bounds__(TAIL_TPB, 1) k_pair_product_tail(const uint64_t* in, size_t count, uint64_t* out, int final_exp, uint8_t* is_some) {
  extern __shared__ __align__(16) uint32_t wide_slots[];     // tree: 32 x 6 slots; afterwards the wide engine's slot file
  const int lp = threadIdx.x >> 1, c = threadIdx.x & 1;
  P12 acc;
  p12_one(acc);
  const size_t trips = (count + TAIL_LP - 1) / TAIL_LP;
#pragma unroll 1
  for (size_t j = 0; j < trips; j++) {
    const size_t i = lp + j * TAIL_LP;
    P12 x;
    if (i < count) {
      const uint64_t* pi = in + FQ12_W * i;
      x.c0.c0.v = ld_fp(pi + 6 * c); x.c0.c1.v = ld_fp(pi + 12 + 6 * c); x.c0.c2.v = ld_fp(pi + 24 + 6 * c);
      x.c1.c0.v = ld_fp(pi + 36 + 6 * c); x.c1.c1.v = ld_fp(pi + 48 + 6 * c); x.c1.c2.v = ld_fp(pi + 60 + 6 * c);
    } else {
      p12_one(x);
    }
    if (j == 0) acc = x; else p12_mul(acc, acc, x);
  }
#pragma unroll 1
  for (int s = TAIL_LP / 2; s >= 1; s >>= 1) {      // at level s the lane pairs [s, 2s) publish, the lane pairs [0, s) multiply
    __syncthreads();
    if (lp >= s && lp < 2 * s) tail_st(wide_slots, lp - s, c, acc);
    __syncthreads();
    if (lp < (s < 16 ? 16 : s)) {        // whole warps only: below 16 lane pairs the rest of warp 0 multiplies along (results unused)
      P12 x;
      tail_ld(x, wide_slots, lp, c);
      p12_mul(acc, acc, x);
    }
  }
  // lane pair 0 holds the product
  if (!final_exp) {
    if (lp == 0) {
      uint64_t* o = out;
      st_fp(o + 6 * c, acc.c0.c0.v); st_fp(o + 12 + 6 * c, acc.c0.c1.v); st_fp(o + 24 + 6 * c, acc.c0.c2.v);
      st_fp(o + 36 + 6 * c, acc.c1.c0.v); st_fp(o + 48 + 6 * c, acc.c1.c1.v); st_fp(o + 60 + 6 * c, acc.c1.c2.v);
    }
    return;
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  bool zero = true;
  if (lp == 0) {
    const Fp* v = reinterpret_cast<const Fp*>(&acc);
    zero = fp_is_zero(v[0]) && fp_is_zero(v[1]) && fp_is_zero(v[2]) && fp_is_zero(v[3]) && fp_is_zero(v[4]) && fp_is_zero(v[5]);
#pragma unroll
    for (int k = 0; k < 6; k++) wide_st(wide_slots, k, c, v[k]);
  }
  const bool all_zero = __all_sync(0xffffffffu, zero);
  wide_load_consts(wide_slots, WIDE_FINAL_EXP_CONST, WIDE_FINAL_EXP_NCONST);
  __syncwarp();
  wide_run(WIDE_FINAL_EXP_CODE, WIDE_FINAL_EXP_NROUNDS, wide_slots, WIDE_FINAL_EXP_NSLOTS);
  if (lp < 6) {
    Fp v = wide_ld(wide_slots, WIDE_FINAL_EXP_OUT[lp], c);
    if (all_zero) v = fp_zero();
    st_fp(out + 12 * lp + 6 * c, v);
  }
  if (threadIdx.x == 0 && is_some) is_some[0] = !all_zero;
}

static const size_t TAIL_SMEM = (size_t)(TAIL_LP / 2 * 6 > WIDE_FINAL_EXP_NSLOTS + 1 ? TAIL_LP / 2 * 6 : WIDE_FINAL_EXP_NSLOTS + 1) * WIDE_SLOT_BYTES;

extern "C" {
int bls_internal_wide_final_exp(bls_ctx* ctx, const bls_fq12* in, bls_fq12* out, uint8_t* is_some, size_t n, cudaStream_t s) {
  k_wide_final_exp<<<(unsigned)n, 32, (WIDE_FINAL_EXP_NSLOTS + 1) * WIDE_SLOT_BYTES, s>>>((const uint64_t*)in, (uint64_t*)out, is_some, n);
  LAUNCH_CHECK();
  return BLS_OK;
}
int bls_internal_wide_pairing(bls_ctx* ctx, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n, cudaStream_t s) {
  k_wide_pairing<<<(unsigned)n, 32, (WIDE_PAIRING_NSLOTS + 1) * WIDE_SLOT_BYTES, s>>>((const uint64_t*)p, (const uint64_t*)q, (uint64_t*)out, n);
  LAUNCH_CHECK();
  return BLS_OK;
}
int bls_internal_wide_miller(bls_ctx* ctx, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n, cudaStream_t s) {
  k_wide_miller<<<(unsigned)n, 32, (WIDE_MILLER_NSLOTS + 1) * WIDE_SLOT_BYTES, s>>>((const uint64_t*)p, (const uint64_t*)q, (uint64_t*)out, n);
  LAUNCH_CHECK();
  return BLS_OK;
}
int bls_internal_wide_pairing_projective(bls_ctx* ctx, const bls_g1* p, const bls_g2* q, bls_fq12* out, size_t n, cudaStream_t s) {
  k_wide_pairing_projective<<<(unsigned)n, 32, (WIDE_PAIRING_NSLOTS + 1) * WIDE_SLOT_BYTES, s>>>((const uint64_t*)p, (const uint64_t*)q, (uint64_t*)out, n);
  LAUNCH_CHECK();
  return BLS_OK;
}
int bls_internal_product_tail(bls_ctx* ctx, const bls_fq12* in, size_t count, bls_fq12* out1, int final_exp, uint8_t* is_some, cudaStream_t s) {
  k_pair_product_tail<<<1, TAIL_TPB, TAIL_SMEM, s>>>((const uint64_t*)in, count, (uint64_t*)out1, final_exp, is_some);
  LAUNCH_CHECK();
  return BLS_OK;
}
}  // extern "C"

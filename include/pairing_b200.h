/* pairing_b200.h -- C ABI of the B200-native batched BLS12-381 pairing / wNAF engine.
 *
 * This is the drop-in boundary beneath the `pairing` crate's trait surface (v0.14.2).  The
 * reference has no FFI of its own: its boundary is the Rust traits `Engine` (src/lib.rs:34-110),
 * `CurveProjective` (src/lib.rs:114-181), `CurveAffine` (src/lib.rs:185-234) and `Wnaf`
 * (src/wnaf.rs:75-179).  Each entry point below names the reference routine it replaces; the Rust
 * shim that binds them is shown in INTEGRATION.md and rust/src/ffi.rs.
 *
 * Data layout (all little-endian, identical byte-for-byte to the crate's in-memory values):
 *   Fq      6 x u64 limbs, Montgomery form (x * 2^384 mod q), always < q   (bls12_381/fq.rs:510,699)
 *   Fq2     c0 | c1                                                        (fq2.rs:8-12)
 *   Fq6     c0 | c1 | c2                                                   (fq6.rs:8-12)
 *   Fq12    c0 | c1                                                        (fq12.rs:8-12)
 *   FrRepr  4 x u64 limbs, canonical integer, NOT Montgomery               (fr.rs:58)
 * Ownership: the caller owns every buffer; nothing is retained after a call returns.
 * Errors: every function returns 0 on success or a negative bls_status; nothing unwinds across
 * the ABI.  There is no CPU fallback: without a CUDA device bls_ctx_create fails.
 * Threading: a bls_ctx may be shared by several host threads -- every entry point holds the context's lock while it runs on the
 * host, so concurrent calls serialise (one ordered stream per device); use one context per thread for concurrency.
 */
#ifndef PAIRING_B200_H
#define PAIRING_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[6]; } bls_fq;
typedef struct { bls_fq c0, c1; } bls_fq2;
typedef struct { bls_fq2 c0, c1, c2; } bls_fq6;
typedef struct { bls_fq6 c0, c1; } bls_fq12;                       /* 576 B */
typedef struct { bls_fq x, y; uint64_t infinity; } bls_g1_affine;  /* 104 B; ec.rs:13-18 */
typedef struct { bls_fq x, y, z; } bls_g1;                         /* 144 B; Jacobian, z == 0 <=> infinity; ec.rs:31-36 */
typedef struct { bls_fq2 x, y; uint64_t infinity; } bls_g2_affine; /* 200 B */
typedef struct { bls_fq2 x, y, z; } bls_g2;                        /* 288 B */
typedef struct { uint64_t l[4]; } bls_fr_repr;                     /* 32 B */
typedef struct { uint64_t l[4]; } bls_fr;                          /* Fr(FrRepr): Montgomery form (x * 2^256 mod r), < r; fr.rs:54-58 */
/* G2Prepared (ec.rs:1615-1619): 68 = 63 doubling + 5 addition coefficient triples in loop order */
typedef struct { bls_fq2 coeffs[68][3]; uint64_t infinity; } bls_g2_prepared; /* 19 592 B */

typedef struct bls_ctx bls_ctx;

typedef enum {
  BLS_OK = 0,
  BLS_ERR_INVALID_ARGUMENT = -1,
  BLS_ERR_NO_DEVICE = -2,
  BLS_ERR_CUDA = -3,
  BLS_ERR_OUT_OF_MEMORY = -4,
  BLS_ERR_UNSUPPORTED = -5
} bls_status;

/* One context = one CUDA device + one ordered stream + reusable device scratch.  `device` is a CUDA
 * ordinal.  Multi-GPU callers create one context per device (one process per GPU under
 * torch.distributed / NCCL, or several contexts in one process) and shard batches themselves;
 * see bls_fq12_product for the only cross-device reduction the path has. */
bls_ctx* bls_ctx_create(int device, int* err);
void bls_ctx_destroy(bls_ctx* ctx);
const char* bls_strerror(int status);
/* text of the last CUDA error seen by this context (empty string if none) */
const char* bls_ctx_last_error(const bls_ctx* ctx);
int bls_ctx_device(const bls_ctx* ctx);
int bls_ctx_sm_count(const bls_ctx* ctx);
/* number of kernel launches issued through this context so far (bench.py's gpu_launches) */
uint64_t bls_ctx_launch_count(const bls_ctx* ctx);
/* The host-buffer entry points stage through device buffers from a per-context memory pool that keeps its memory between
 * calls (no cudaMalloc / cudaFree per call after the first); bls_ctx_trim hands it back, keeping at most keep_bytes. */
int bls_ctx_trim(bls_ctx*, size_t keep_bytes);
/* bls_pairing_* / bls_final_exponentiation_* pick between two kernels by batch size: up to these many elements one WARP
 * works on each element (latency path: ~2 ms for one pairing or for a thousand, the crate's bench_pairing_full shape),
 * above them one lane pair does (throughput path: 8.8 ms of latency, 1.49 M pairings/s).  Same bits either way.
 * 0 disables the latency path.  Defaults: 2560 / 2560 (where the two curves cross on a B200). */
int bls_ctx_set_latency_path_limits(bls_ctx*, size_t max_pairings, size_t max_final_exps);

/* ------------------------------------------------------------------ pairing engine (host buffers) */

/* G2Affine::prepare / G2Prepared::from_affine, bls12_381/mod.rs:168-358, for n points. */
int bls_g2_prepare_batch(bls_ctx*, const bls_g2_affine* q, bls_g2_prepared* out, size_t n);
/* n independent Engine::miller_loop(&[(&p_i.prepare(), &q_i.prepare())]), mod.rs:40-102.
 * Pairs with an infinity member give Fq12::one() (mod.rs:49-54). */
int bls_miller_loop_batch(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n);
/* same, from already prepared G2 coefficients (the reference's actual miller_loop signature) */
int bls_miller_loop_prepared_batch(bls_ctx*, const bls_g1_affine* p, const bls_g2_prepared* q, bls_fq12* out, size_t n);
/* n independent Miller loops / pairings e(P_i, Q) against ONE prepared G2 point: the reference's
 * `let q = Q.prepare(); for p_i { Engine::miller_loop(&[(&p_i.prepare(), &q)]) }` (the reason G2Prepared exists,
 * ec.rs:1615-1619).  The coefficients are staged once per block in shared memory by a TMA bulk copy. */
int bls_miller_loop_shared_q_batch(bls_ctx*, const bls_g1_affine* p, const bls_g2_prepared* q1, bls_fq12* out, size_t n);
int bls_pairing_shared_q_batch(bls_ctx*, const bls_g1_affine* p, const bls_g2_prepared* q1, bls_fq12* out, size_t n);
/* ONE Engine::miller_loop over n pairs: the product of the n Miller values (mod.rs:80-95). */
int bls_multi_miller_loop(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1);
int bls_multi_miller_loop_prepared(bls_ctx*, const bls_g1_affine* p, const bls_g2_prepared* q, size_t n, bls_fq12* out1);
/* Engine::final_exponentiation(&Engine::miller_loop(pairs)) in ONE call -- the batch-verification shape (BASELINE
 * configs[2]): the Miller kernel leaves one partial product per block, one tail kernel folds them and runs the single
 * final exponentiation on the warp-cooperative engine.  *is_some = 0 marks the reference's None (product 0). */
int bls_pairing_product(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1, uint8_t* is_some);
/* Engine::final_exponentiation, mod.rs:104-160; is_some[i] = 0 marks the reference's None (input 0),
 * in which case out[i] is all-zero.  is_some may be NULL. */
int bls_final_exponentiation_batch(bls_ctx*, const bls_fq12* in, bls_fq12* out, uint8_t* is_some, size_t n);
/* Engine::pairing on affine inputs, lib.rs:101-109 (Miller loop + final exponentiation). */
int bls_pairing_batch(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n);
/* Engine::pairing(p, q) with PROJECTIVE arguments (`G1: Into<G1Affine>`, `G2: Into<G2Affine>`, lib.rs:101-109) -- how the crate's
 * own bench_pairing_full calls it (benches/bls12_381/mod.rs:91-107): the two into_affine conversions (ec.rs:586-619) are fused in
 * front of the pairing.  A pair with an infinity member gives Fq12::one(). */
int bls_pairing_projective_batch(bls_ctx*, const bls_g1* p, const bls_g2* q, bls_fq12* out, size_t n);
/* product of n Fq12 values (Fq12::mul_assign, fq12.rs:116-130): merges per-device partial Miller
 * products before the single final exponentiation of a sharded multi_miller_loop. */
int bls_fq12_product(bls_ctx*, const bls_fq12* in, size_t n, bls_fq12* out1);

/* Field::pow on Fq12 with an FrRepr exponent (lib.rs:306-324): GT exponentiation e(P,Q)^k, the step after
 * the path in the reference's bilinearity test (tests/engine.rs:93-126). */
int bls_fq12_pow_batch(bls_ctx*, const bls_fq12* a, const bls_fr_repr* k, bls_fq12* out, size_t n);

/* ------------------------------------------------------------------ curve groups (host buffers) */

/* Wnaf::new().scalar(k_i).base(g_i): window from the scalar (ec.rs:895-905 / 1586-1596),
 * wnaf_table + wnaf_form + wnaf_exp (wnaf.rs:4-71).  Output Jacobian triples are bit-identical
 * to the reference's. */
int bls_g1_wnaf_mul_batch(bls_ctx*, const bls_g1* bases, const bls_fr_repr* k, bls_g1* out, size_t n);
int bls_g2_wnaf_mul_batch(bls_ctx*, const bls_g2* bases, const bls_fr_repr* k, bls_g2* out, size_t n);
/* same with an explicit window 2..13 (the crate-internal wnaf_table/wnaf_form/wnaf_exp triple, as swept by
 * src/tests/curve.rs:78; windows above 7 keep their 2^(w-1)-entry tables in a per-call global-memory scratch) */
int bls_g1_wnaf_mul_window_batch(bls_ctx*, const bls_g1* bases, const bls_fr_repr* k, bls_g1* out, size_t n, int window);
int bls_g2_wnaf_mul_window_batch(bls_ctx*, const bls_g2* bases, const bls_fr_repr* k, bls_g2* out, size_t n, int window);
/* Fixed-base mode Wnaf::new().base(g, num_scalars) then .scalar(k_i) for every scalar (wnaf.rs:93-107,
 * 169-178): ONE window table of 2^(window-1) entries (wnaf_table, wnaf.rs:4-15) shared by all scalars;
 * `window` = recommended_wnaf_for_num_scalars(n) (ec.rs:907-921 / 1598-1612), 2..16.
 * bls_g*_wnaf_table returns the table itself (2^(window-1) Jacobian points, bit-identical to the crate's). */
int bls_g1_wnaf_fixed_base_batch(bls_ctx*, const bls_g1* base, int window, const bls_fr_repr* k, bls_g1* out, size_t n);
int bls_g2_wnaf_fixed_base_batch(bls_ctx*, const bls_g2* base, int window, const bls_fr_repr* k, bls_g2* out, size_t n);
int bls_g1_wnaf_table(bls_ctx*, const bls_g1* base, int window, bls_g1* table);
int bls_g2_wnaf_table(bls_ctx*, const bls_g2* base, int window, bls_g2* table);
/* CurveProjective::mul_assign (double-and-add), ec.rs:534-553 */
int bls_g1_mul_batch(bls_ctx*, const bls_g1* bases, const bls_fr_repr* k, bls_g1* out, size_t n);
int bls_g2_mul_batch(bls_ctx*, const bls_g2* bases, const bls_fr_repr* k, bls_g2* out, size_t n);
/* CurveAffine::mul, ec.rs:174-177: $affine::mul_bits (MSB-first over all 256 bits, mixed additions) -- a different
 * Jacobian representative of the same point as mul_assign on the projective form */
int bls_g1_affine_mul_batch(bls_ctx*, const bls_g1_affine* a, const bls_fr_repr* k, bls_g1* out, size_t n);
int bls_g2_affine_mul_batch(bls_ctx*, const bls_g2_affine* a, const bls_fr_repr* k, bls_g2* out, size_t n);
/* CurveProjective::batch_normalization, ec.rs:246-294, in place */
int bls_g1_batch_normalization(bls_ctx*, bls_g1* inout, size_t n);
int bls_g2_batch_normalization(bls_ctx*, bls_g2* inout, size_t n);
/* From<projective> for affine / CurveProjective::into_affine, ec.rs:586-619 */
int bls_g1_into_affine_batch(bls_ctx*, const bls_g1* in, bls_g1_affine* out, size_t n);
int bls_g2_into_affine_batch(bls_ctx*, const bls_g2* in, bls_g2_affine* out, size_t n);
/* element-wise group law: op = BLS_PT_*.  b is bls_g1/bls_g2 (ADD, SUB), bls_g*_affine (ADD_MIXED)
 * or NULL (DOUBLE, NEGATE).  ec.rs:296-532, lib.rs:156-160 */
enum { BLS_PT_DOUBLE = 0, BLS_PT_ADD = 1, BLS_PT_ADD_MIXED = 2, BLS_PT_NEGATE = 3, BLS_PT_SUB = 6 };
int bls_g1_op_batch(bls_ctx*, int op, const bls_g1* a, const void* b, bls_g1* out, size_t n);
int bls_g2_op_batch(bls_ctx*, int op, const bls_g2* a, const void* b, bls_g2* out, size_t n);

/* ------------------------------------------------------------------ point encodings (host buffers)
 * EncodedPoint::into_affine (checked != 0: on-curve and r-order-subgroup checks) / into_affine_unchecked and
 * EncodedPoint::from_affine for G1Uncompressed (96 B) / G1Compressed (48 B) (ec.rs:645-868) and
 * G2Uncompressed (192 B) / G2Compressed (96 B) (ec.rs:1292-1540).  `bytes` holds n encodings back to back.
 * status[i] mirrors GroupDecodingError (src/lib.rs:468-497); a rejected element decodes to the zero point. */
enum {
  BLS_DEC_OK = 0, BLS_DEC_UNEXPECTED_COMPRESSION_MODE = 1, BLS_DEC_UNEXPECTED_INFORMATION = 2,
  BLS_DEC_NOT_ON_CURVE = 3, BLS_DEC_NOT_IN_SUBGROUP = 4,
  BLS_DEC_COORDINATE = 16 /* + coordinate: G1 0 = x, 1 = y; G2 0 = x (c0), 1 = x (c1), 2 = y (c0), 3 = y (c1) */
};
int bls_g1_decode_batch(bls_ctx*, const uint8_t* bytes, int compressed, int checked, bls_g1_affine* out, uint8_t* status, size_t n);
int bls_g2_decode_batch(bls_ctx*, const uint8_t* bytes, int compressed, int checked, bls_g2_affine* out, uint8_t* status, size_t n);
int bls_g1_encode_batch(bls_ctx*, const bls_g1_affine* in, int compressed, uint8_t* bytes, size_t n);
int bls_g2_encode_batch(bls_ctx*, const bls_g2_affine* in, int compressed, uint8_t* bytes, size_t n);

/* G::rand with the randomness supplied by the caller (ec.rs:199-214): $affine::get_point_from_x (ec.rs:102-123; is_some = 0
 * when x^3 + b has no square root) and $affine::scale_by_cofactor (mul_bits with the cofactor, ec.rs:86-94, 871-875,
 * 1564-1578: Jacobian output, infinity possible -- the reference retries). */
int bls_g1_point_from_x_batch(bls_ctx*, const bls_fq* x, const uint8_t* greatest, bls_g1_affine* out, uint8_t* is_some, size_t n);
int bls_g2_point_from_x_batch(bls_ctx*, const bls_fq2* x, const uint8_t* greatest, bls_g2_affine* out, uint8_t* is_some, size_t n);
int bls_g1_scale_by_cofactor_batch(bls_ctx*, const bls_g1_affine* in, bls_g1* out, size_t n);
int bls_g2_scale_by_cofactor_batch(bls_ctx*, const bls_g2_affine* in, bls_g2* out, size_t n);

/* ------------------------------------------------------------------ field tower (host buffers) */
/* Element-wise field operations, the `Field` trait methods of src/lib.rs:267-325 on
 * Fq (degree 1), Fq2 (2), Fq6 (6), Fq12 (12).  `b` is an array of the same element type or NULL.
 * ok[i] = 0 marks None (inverse of 0, from_repr of a non-canonical value); may be NULL. */
enum {
  BLS_OP_ADD = 0, BLS_OP_SUB = 1, BLS_OP_MUL = 2, BLS_OP_SQR = 3, BLS_OP_NEG = 4, BLS_OP_DBL = 5,
  BLS_OP_INV = 6, BLS_OP_FROM_REPR = 7, BLS_OP_INTO_REPR = 8, BLS_OP_MUL_NONRES = 9,
  BLS_OP_FROB1 = 10, BLS_OP_FROB2 = 11, BLS_OP_FROB3 = 12, BLS_OP_CONJ = 13,
  BLS_OP_MUL_BY_014 = 14, /* Fq12 only: sparse operands b.c0.c0, b.c0.c1, b.c1.c1 (fq12.rs:34-48) */
  BLS_OP_MUL_BY_01 = 15,  /* Fq6 only: b.c0, b.c1 (fq6.rs:68-109) */
  BLS_OP_MUL_BY_1 = 16,   /* Fq6 only: b.c1 (fq6.rs:40-66) */
  BLS_OP_SQRT = 17,       /* Fq, Fq2: SqrtField::sqrt (fq.rs:1147-1170, fq2.rs:167-221); ok = 0 for a non-residue */
  /* lane-pair tower only (bls_pair_field_op_batch): */
  BLS_OP_MUL_BY_LINE_PAIR = 18,  /* Fq12: a * l * m for two sparse lines packed in b as (l.c0, l.c1, l.c4, m.c0, m.c1, m.c4) */
  BLS_OP_CYCLOTOMIC_SQR = 19     /* Fq12: Granger-Scott squaring; equals `square` for elements of the cyclotomic subgroup */
};
int bls_field_op_batch(bls_ctx*, int degree, int op, const void* a, const void* b, void* out, uint8_t* ok, size_t n);
/* The same element-wise operations (degree 2, 6, 12) on the LANE-PAIR tower the pairing kernels run (two lanes per element,
 * lazily reduced dual products): a separate implementation of fq2.rs / fq6.rs / fq12.rs, so it gets its own parity entry.
 * Operands may be any representative in [0, 2q]; results are canonical. */
int bls_pair_field_op_batch(bls_ctx*, int degree, int op, const void* a, const void* b, void* out, uint8_t* ok, size_t n);
int bls_pair_field_op_dev(bls_ctx*, int degree, int op, const void* a, const void* b, void* out, uint8_t* ok, size_t n, void* stream);

/* The scalar field Fr (fr.rs:324-572), element-wise: op is BLS_OP_ADD / SUB / MUL / SQR / NEG / DBL / INV /
 * FROM_REPR (bls_fr_repr -> bls_fr; ok = 0 for values >= r) / INTO_REPR (bls_fr -> bls_fr_repr).  The step after the
 * path: callers combine scalars (c * d in the reference's bilinearity test, tests/engine.rs:117-119). */
int bls_fr_op_batch(bls_ctx*, int op, const bls_fr* a, const bls_fr* b, bls_fr* out, uint8_t* ok, size_t n);

/* ------------------------------------------------------------------ device-pointer variants
 * Same semantics; every pointer is a device pointer on the context's device, the work is enqueued
 * on `stream` (a cudaStream_t, NULL = the context's own stream) and NOT synchronised.  `scratch`
 * arguments are caller-provided device buffers of the stated size (so no allocation happens on the
 * timed path). */
int bls_g2_prepare_dev(bls_ctx*, const bls_g2_affine* q, bls_g2_prepared* out, size_t n, void* stream);
int bls_miller_loop_dev(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n, void* stream);
int bls_miller_loop_prepared_dev(bls_ctx*, const bls_g1_affine* p, const bls_g2_prepared* q, bls_fq12* out, size_t n, void* stream);
int bls_final_exponentiation_dev(bls_ctx*, const bls_fq12* in, bls_fq12* out, uint8_t* is_some, size_t n, void* stream);
int bls_pairing_dev(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n, void* stream);
int bls_pairing_projective_dev(bls_ctx*, const bls_g1* p, const bls_g2* q, bls_fq12* out, size_t n, void* stream);
/* q1: ONE prepared point in device memory, 16-byte aligned; final_exp != 0 adds the final exponentiation */
int bls_miller_loop_shared_q_dev(bls_ctx*, const bls_g1_affine* p, const bls_g2_prepared* q1, bls_fq12* out, size_t n, int final_exp, void* stream);
int bls_fq12_pow_dev(bls_ctx*, const bls_fq12* a, const bls_fr_repr* k, bls_fq12* out, size_t n, void* stream);
/* scratch: bls_multi_miller_scratch_bytes(ctx, n) bytes */
size_t bls_multi_miller_scratch_bytes(const bls_ctx*, size_t n);
int bls_multi_miller_loop_dev(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1, void* scratch, void* stream);
int bls_pairing_product_dev(bls_ctx*, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1, uint8_t* is_some, void* scratch, void* stream);
/* product of n Fq12 values by ONE block (n up to a few thousand: the per-block or per-device partials of a sharded
 * multi_miller_loop), then -- final_exp != 0 -- Engine::final_exponentiation of the product (mod.rs:104-160) on the
 * warp-cooperative engine.  No scratch.  is_some (device pointer, may be NULL) as in bls_final_exponentiation_dev. */
int bls_fq12_product_tail_dev(bls_ctx*, const bls_fq12* in, size_t n, bls_fq12* out1, int final_exp, uint8_t* is_some, void* stream);
size_t bls_fq12_product_scratch_bytes(const bls_ctx*, size_t n);
int bls_fq12_product_dev(bls_ctx*, const bls_fq12* in, size_t n, bls_fq12* out1, void* scratch, void* stream);
int bls_g1_wnaf_mul_dev(bls_ctx*, const bls_g1* bases, const bls_fr_repr* k, bls_g1* out, size_t n, int window, void* stream);
int bls_g2_wnaf_mul_dev(bls_ctx*, const bls_g2* bases, const bls_fr_repr* k, bls_g2* out, size_t n, int window, void* stream);
/* fixed-base mode on the device: `table` holds 2^(window-1) points, built by bls_g*_wnaf_table_dev */
int bls_g1_wnaf_table_dev(bls_ctx*, const bls_g1* base, int window, bls_g1* table, void* stream);
int bls_g2_wnaf_table_dev(bls_ctx*, const bls_g2* base, int window, bls_g2* table, void* stream);
int bls_g1_wnaf_fixed_base_dev(bls_ctx*, const bls_g1* table, int window, const bls_fr_repr* k, bls_g1* out, size_t n, void* stream);
int bls_g2_wnaf_fixed_base_dev(bls_ctx*, const bls_g2* table, int window, const bls_fr_repr* k, bls_g2* out, size_t n, void* stream);
/* scratch: bls_batch_normalization_scratch_bytes(ctx, degree, n) bytes; degree 1 = G1, 2 = G2 */
size_t bls_batch_normalization_scratch_bytes(const bls_ctx*, int degree, size_t n);
int bls_g1_batch_normalization_dev(bls_ctx*, bls_g1* inout, size_t n, void* scratch, void* stream);
int bls_g2_batch_normalization_dev(bls_ctx*, bls_g2* inout, size_t n, void* scratch, void* stream);

/* ------------------------------------------------------------------ several GPUs behind ONE call
 * Engine::miller_loop takes ALL pairs of a product in one call (mod.rs:40-102), and a Rust / C caller is one process.
 * A bls_mgpu binds n devices of one node.  An n-pair product is split into contiguous shards, one host thread per
 * device; every device reduces its shard to ONE 576-byte partial product, the partials travel to the first device by
 * peer copies (NVLink; staged through the host where peer access is unavailable), and the first device folds them and
 * runs the single final exponentiation.  Independent batches (pairings, wNAF multiplications) are sharded the same
 * way with no exchange step.  devices == NULL selects devices 0 .. n_devices-1. */
typedef struct bls_mgpu bls_mgpu;
bls_mgpu* bls_mgpu_create(const int* devices, int n_devices, int* err);
void bls_mgpu_destroy(bls_mgpu*);
int bls_mgpu_device_count(const bls_mgpu*);
bls_ctx* bls_mgpu_ctx(bls_mgpu*, int i);   /* the single-device context of shard i (owned by the bls_mgpu) */
int bls_mgpu_multi_miller_loop(bls_mgpu*, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1);
int bls_mgpu_pairing_product(bls_mgpu*, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1, uint8_t* is_some);
int bls_mgpu_pairing_batch(bls_mgpu*, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n);
int bls_mgpu_g1_wnaf_mul_batch(bls_mgpu*, const bls_g1* bases, const bls_fr_repr* k, bls_g1* out, size_t n);
int bls_mgpu_g2_wnaf_mul_batch(bls_mgpu*, const bls_g2* bases, const bls_fr_repr* k, bls_g2* out, size_t n);
/* host wall-clock phases of the last product call, milliseconds: ms3[0] = shards (H2D + Miller kernel + per-device
 * fold + peer copy, slowest device), ms3[1] = fold of the partials + final exponentiation + D2H, ms3[2] = whole call */
int bls_mgpu_last_phase_ms(const bls_mgpu*, double* ms3);

/* ------------------------------------------------------------------ measurement
 * Register-resident integer-multiply microbenchmark: the roofline denominator for this path
 * (MEASURED_PEAKS.json has no integer figure).  variant 0 = chains of dependent IMAD.WIDE.U32
 * (32x32->64, the only multiply in the loop -- checked on the SASS by tests/test_abi.py),
 * 1 = back-to-back fp_mul (300 MAC32 each), 2 = 32-bit IMAD, 3 = carry-linked
 * IMAD.WIDE.U32.X rows only (mad.lo.cc / madc.hi.cc chains, no reduction), 4 = back-to-back dedicated squarings
 * (fq.rs:963-1016; 234 MAC32 each).
 * Returns multiply-accumulates per second in *macs_per_s and the kernel time in *ms. */
int bls_imad_peak(bls_ctx*, int variant, int iters, double* macs_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* PAIRING_B200_H */

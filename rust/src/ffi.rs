//! Raw bindings of include/pairing_b200.h.  POD mirrors are `#[repr(C)]`; the crate's own `Fq`,
//! `G1Affine`, ... are NOT (`pub(crate)` fields, no repr), so `lib.rs` marshals explicitly.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Copy, Clone)] pub struct bls_fq { pub l: [u64; 6] }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_fq2 { pub c0: bls_fq, pub c1: bls_fq }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_fq6 { pub c0: bls_fq2, pub c1: bls_fq2, pub c2: bls_fq2 }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_fq12 { pub c0: bls_fq6, pub c1: bls_fq6 }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_g1_affine { pub x: bls_fq, pub y: bls_fq, pub infinity: u64 }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_g1 { pub x: bls_fq, pub y: bls_fq, pub z: bls_fq }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_g2_affine { pub x: bls_fq2, pub y: bls_fq2, pub infinity: u64 }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_g2 { pub x: bls_fq2, pub y: bls_fq2, pub z: bls_fq2 }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_fr_repr { pub l: [u64; 4] }
#[repr(C)] #[derive(Copy, Clone)] pub struct bls_g2_prepared { pub coeffs: [[bls_fq2; 3]; 68], pub infinity: u64 }
pub enum bls_ctx {}
/// several devices of one node behind one call (include/pairing_b200.h, "several GPUs")
pub enum bls_mgpu {}

pub const BLS_OK: c_int = 0;

extern "C" {
    pub fn bls_ctx_create(device: c_int, err: *mut c_int) -> *mut bls_ctx;
    pub fn bls_ctx_destroy(ctx: *mut bls_ctx);
    pub fn bls_strerror(status: c_int) -> *const c_char;
    pub fn bls_ctx_last_error(ctx: *const bls_ctx) -> *const c_char;

    pub fn bls_g2_prepare_batch(ctx: *mut bls_ctx, q: *const bls_g2_affine, out: *mut bls_g2_prepared, n: usize) -> c_int;
    pub fn bls_miller_loop_batch(ctx: *mut bls_ctx, p: *const bls_g1_affine, q: *const bls_g2_affine, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_miller_loop_prepared_batch(ctx: *mut bls_ctx, p: *const bls_g1_affine, q: *const bls_g2_prepared, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_miller_loop_shared_q_batch(ctx: *mut bls_ctx, p: *const bls_g1_affine, q1: *const bls_g2_prepared, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_pairing_shared_q_batch(ctx: *mut bls_ctx, p: *const bls_g1_affine, q1: *const bls_g2_prepared, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_multi_miller_loop(ctx: *mut bls_ctx, p: *const bls_g1_affine, q: *const bls_g2_affine, n: usize, out1: *mut bls_fq12) -> c_int;
    pub fn bls_multi_miller_loop_prepared(ctx: *mut bls_ctx, p: *const bls_g1_affine, q: *const bls_g2_prepared, n: usize, out1: *mut bls_fq12) -> c_int;
    pub fn bls_pairing_projective_batch(ctx: *mut bls_ctx, p: *const bls_g1, q: *const bls_g2, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_pairing_product(ctx: *mut bls_ctx, p: *const bls_g1_affine, q: *const bls_g2_affine, n: usize, out1: *mut bls_fq12, is_some: *mut u8) -> c_int;
    pub fn bls_mgpu_create(devices: *const c_int, n_devices: c_int, err: *mut c_int) -> *mut bls_mgpu;
    pub fn bls_mgpu_destroy(m: *mut bls_mgpu);
    pub fn bls_mgpu_device_count(m: *const bls_mgpu) -> c_int;
    pub fn bls_mgpu_multi_miller_loop(m: *mut bls_mgpu, p: *const bls_g1_affine, q: *const bls_g2_affine, n: usize, out1: *mut bls_fq12) -> c_int;
    pub fn bls_mgpu_pairing_product(m: *mut bls_mgpu, p: *const bls_g1_affine, q: *const bls_g2_affine, n: usize, out1: *mut bls_fq12, is_some: *mut u8) -> c_int;
    pub fn bls_mgpu_pairing_batch(m: *mut bls_mgpu, p: *const bls_g1_affine, q: *const bls_g2_affine, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_mgpu_g1_wnaf_mul_batch(m: *mut bls_mgpu, bases: *const bls_g1, k: *const bls_fr_repr, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_mgpu_g2_wnaf_mul_batch(m: *mut bls_mgpu, bases: *const bls_g2, k: *const bls_fr_repr, out: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_final_exponentiation_batch(ctx: *mut bls_ctx, input: *const bls_fq12, out: *mut bls_fq12, is_some: *mut u8, n: usize) -> c_int;
    pub fn bls_pairing_batch(ctx: *mut bls_ctx, p: *const bls_g1_affine, q: *const bls_g2_affine, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_g1_decode_batch(ctx: *mut bls_ctx, bytes: *const u8, compressed: c_int, checked: c_int, out: *mut bls_g1_affine, status: *mut u8, n: usize) -> c_int;
    pub fn bls_g2_decode_batch(ctx: *mut bls_ctx, bytes: *const u8, compressed: c_int, checked: c_int, out: *mut bls_g2_affine, status: *mut u8, n: usize) -> c_int;
    pub fn bls_g1_encode_batch(ctx: *mut bls_ctx, input: *const bls_g1_affine, compressed: c_int, bytes: *mut u8, n: usize) -> c_int;
    pub fn bls_g2_encode_batch(ctx: *mut bls_ctx, input: *const bls_g2_affine, compressed: c_int, bytes: *mut u8, n: usize) -> c_int;
    pub fn bls_g1_affine_mul_batch(ctx: *mut bls_ctx, a: *const bls_g1_affine, k: *const bls_fr_repr, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_affine_mul_batch(ctx: *mut bls_ctx, a: *const bls_g2_affine, k: *const bls_fr_repr, out: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_g1_point_from_x_batch(ctx: *mut bls_ctx, x: *const bls_fq, greatest: *const u8, out: *mut bls_g1_affine, is_some: *mut u8, n: usize) -> c_int;
    pub fn bls_g2_point_from_x_batch(ctx: *mut bls_ctx, x: *const bls_fq2, greatest: *const u8, out: *mut bls_g2_affine, is_some: *mut u8, n: usize) -> c_int;
    pub fn bls_g1_scale_by_cofactor_batch(ctx: *mut bls_ctx, input: *const bls_g1_affine, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_scale_by_cofactor_batch(ctx: *mut bls_ctx, input: *const bls_g2_affine, out: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_fr_op_batch(ctx: *mut bls_ctx, op: c_int, a: *const bls_fr_repr, b: *const bls_fr_repr, out: *mut bls_fr_repr, ok: *mut u8, n: usize) -> c_int;
    pub fn bls_fq12_pow_batch(ctx: *mut bls_ctx, a: *const bls_fq12, k: *const bls_fr_repr, out: *mut bls_fq12, n: usize) -> c_int;
    pub fn bls_fq12_product(ctx: *mut bls_ctx, input: *const bls_fq12, n: usize, out1: *mut bls_fq12) -> c_int;

    pub fn bls_g1_wnaf_mul_batch(ctx: *mut bls_ctx, bases: *const bls_g1, k: *const bls_fr_repr, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_wnaf_mul_batch(ctx: *mut bls_ctx, bases: *const bls_g2, k: *const bls_fr_repr, out: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_g1_wnaf_mul_window_batch(ctx: *mut bls_ctx, bases: *const bls_g1, k: *const bls_fr_repr, out: *mut bls_g1, n: usize, window: c_int) -> c_int;
    pub fn bls_g2_wnaf_mul_window_batch(ctx: *mut bls_ctx, bases: *const bls_g2, k: *const bls_fr_repr, out: *mut bls_g2, n: usize, window: c_int) -> c_int;
    pub fn bls_g1_wnaf_fixed_base_batch(ctx: *mut bls_ctx, base: *const bls_g1, window: c_int, k: *const bls_fr_repr, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_wnaf_fixed_base_batch(ctx: *mut bls_ctx, base: *const bls_g2, window: c_int, k: *const bls_fr_repr, out: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_g1_wnaf_table(ctx: *mut bls_ctx, base: *const bls_g1, window: c_int, table: *mut bls_g1) -> c_int;
    pub fn bls_g2_wnaf_table(ctx: *mut bls_ctx, base: *const bls_g2, window: c_int, table: *mut bls_g2) -> c_int;
    pub fn bls_g1_mul_batch(ctx: *mut bls_ctx, bases: *const bls_g1, k: *const bls_fr_repr, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_mul_batch(ctx: *mut bls_ctx, bases: *const bls_g2, k: *const bls_fr_repr, out: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_g1_batch_normalization(ctx: *mut bls_ctx, inout: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_batch_normalization(ctx: *mut bls_ctx, inout: *mut bls_g2, n: usize) -> c_int;
    pub fn bls_g1_into_affine_batch(ctx: *mut bls_ctx, input: *const bls_g1, out: *mut bls_g1_affine, n: usize) -> c_int;
    pub fn bls_g2_into_affine_batch(ctx: *mut bls_ctx, input: *const bls_g2, out: *mut bls_g2_affine, n: usize) -> c_int;
    pub fn bls_g1_op_batch(ctx: *mut bls_ctx, op: c_int, a: *const bls_g1, b: *const c_void, out: *mut bls_g1, n: usize) -> c_int;
    pub fn bls_g2_op_batch(ctx: *mut bls_ctx, op: c_int, a: *const bls_g2, b: *const c_void, out: *mut bls_g2, n: usize) -> c_int;
}

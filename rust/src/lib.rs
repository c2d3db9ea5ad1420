//! Slice-level GPU entry points next to the `pairing` crate's scalar trait methods.
//!
//! This file lives INSIDE a fork of the crate, as the module `crate::bls12_381::gpu` (`mod gpu;` in
//! src/bls12_381/mod.rs, `ffi.rs` next to it as `gpu/ffi.rs`): the point fields are `pub(crate)`
//! (src/bls12_381/ec.rs:15-17, 33-35), so marshalling needs crate visibility, and the two accessors it
//! uses on `Fq` / `Fr` -- `mont_limbs()` and `from_mont_limbs()`, raw access to the Montgomery limbs --
//! are the two-line patch shown in INTEGRATION.md (the upstream fields are private to fq.rs / fr.rs).
//! Marshalling through the public API instead (`into_repr` / `from_repr`) would cost a Montgomery
//! conversion per coordinate in each direction.  The
//! scalar trait methods (`Engine::pairing`, `CurveProjective::mul_assign`, ...) stay as they are;
//! callers with batches use the functions below.  One `Gpu` = one `bls_ctx`; it is `Send` and is
//! wrapped in a `Mutex` because calls on a context must be serialised.
//!
//! Authored, not compiled here (no Rust toolchain in the build image) -- see INTEGRATION.md.
pub mod ffi;

use self::ffi::*;
use crate::bls12_381::{Fq12, FrRepr, G1Affine, G2Affine, G1, G2};
use std::sync::Mutex;

#[derive(Debug)]
pub struct GpuError(pub i32, pub String);

pub struct Gpu { ctx: Mutex<*mut bls_ctx> }
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}

impl Gpu {
    /// One context per CUDA device; there is no CPU fallback (fails without a device).
    pub fn new(device: i32) -> Result<Gpu, GpuError> {
        let mut err = 0;
        let ctx = unsafe { bls_ctx_create(device, &mut err) };
        if ctx.is_null() { return Err(GpuError(err, strerror(err))); }
        Ok(Gpu { ctx: Mutex::new(ctx) })
    }

    /// `Bls12::pairing(p_i, q_i)` for every i (src/lib.rs:101-109).
    pub fn pairing_batch(&self, p: &[G1Affine], q: &[G2Affine]) -> Result<Vec<Fq12>, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq: Vec<bls_g2_affine> = q.iter().map(marshal::g2_affine).collect();
        let mut out = vec![marshal::FQ12_ZERO; p.len()];
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_pairing_batch(*ctx, pp.as_ptr(), qq.as_ptr(), out.as_mut_ptr(), p.len()) }, *ctx)?;
        Ok(out.iter().map(marshal::fq12_back).collect())
    }

    /// `Bls12::pairing(p_i, q_i)` with projective arguments (`Into<G1Affine>`, `Into<G2Affine>`): conversions fused in front.
    pub fn pairing_projective_batch(&self, p: &[G1], q: &[G2]) -> Result<Vec<Fq12>, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1> = p.iter().map(marshal::g1).collect();
        let qq: Vec<bls_g2> = q.iter().map(marshal::g2).collect();
        let mut out = vec![marshal::FQ12_ZERO; p.len()];
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_pairing_projective_batch(*ctx, pp.as_ptr(), qq.as_ptr(), out.as_mut_ptr(), p.len()) }, *ctx)?;
        Ok(out.iter().map(marshal::fq12_back).collect())
    }

    /// `Bls12::miller_loop(&[(&p_0.prepare(), &q_0.prepare()), ...])`: one shared accumulator (mod.rs:40-102).
    pub fn multi_miller_loop(&self, p: &[G1Affine], q: &[G2Affine]) -> Result<Fq12, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq: Vec<bls_g2_affine> = q.iter().map(marshal::g2_affine).collect();
        let mut out = marshal::FQ12_ZERO;
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_multi_miller_loop(*ctx, pp.as_ptr(), qq.as_ptr(), p.len(), &mut out) }, *ctx)?;
        Ok(marshal::fq12_back(&out))
    }

    /// `Bls12::final_exponentiation(&Bls12::miller_loop(pairs))` in one call: the batch-verification shape.
    pub fn pairing_product(&self, p: &[G1Affine], q: &[G2Affine]) -> Result<Option<Fq12>, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq: Vec<bls_g2_affine> = q.iter().map(marshal::g2_affine).collect();
        let (mut out, mut some) = (marshal::FQ12_ZERO, 0u8);
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_pairing_product(*ctx, pp.as_ptr(), qq.as_ptr(), p.len(), &mut out, &mut some) }, *ctx)?;
        Ok(if some != 0 { Some(marshal::fq12_back(&out)) } else { None })
    }

    /// `Bls12::final_exponentiation` per element; `None` where the input is zero (mod.rs:104-160).
    pub fn final_exponentiation_batch(&self, f: &[Fq12]) -> Result<Vec<Option<Fq12>>, GpuError> {
        let ff: Vec<bls_fq12> = f.iter().map(marshal::fq12).collect();
        let mut out = vec![marshal::FQ12_ZERO; f.len()];
        let mut some = vec![0u8; f.len()];
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_final_exponentiation_batch(*ctx, ff.as_ptr(), out.as_mut_ptr(), some.as_mut_ptr(), f.len()) }, *ctx)?;
        Ok(out.iter().zip(some).map(|(v, s)| if s != 0 { Some(marshal::fq12_back(v)) } else { None }).collect())
    }

    /// `Wnaf::new().scalar(k_i).base(g_i)` per element (wnaf.rs:111-128, 158-164); Jacobian outputs are
    /// the reference's exact (X, Y, Z) triples.
    pub fn g1_wnaf_mul_batch(&self, bases: &[G1], k: &[FrRepr]) -> Result<Vec<G1>, GpuError> {
        assert_eq!(bases.len(), k.len());
        let bb: Vec<bls_g1> = bases.iter().map(marshal::g1).collect();
        let kk: Vec<bls_fr_repr> = k.iter().map(|r| bls_fr_repr { l: r.0 }).collect();
        let mut out = bb.clone();
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_g1_wnaf_mul_batch(*ctx, bb.as_ptr(), kk.as_ptr(), out.as_mut_ptr(), bb.len()) }, *ctx)?;
        Ok(out.iter().map(marshal::g1_back).collect())
    }

    /// `G1::batch_normalization(&mut v)` (ec.rs:246-294).
    pub fn g1_batch_normalization(&self, v: &mut [G1]) -> Result<(), GpuError> {
        let mut vv: Vec<bls_g1> = v.iter().map(marshal::g1).collect();
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_g1_batch_normalization(*ctx, vv.as_mut_ptr(), vv.len()) }, *ctx)?;
        for (dst, src) in v.iter_mut().zip(vv.iter()) { *dst = marshal::g1_back(src); }
        Ok(())
    }
    /// `Wnaf::new().base(g, k.len()).scalar(k_i)` for every scalar: one shared window table (wnaf.rs:93-107, 169-178).
    pub fn g1_wnaf_fixed_base(&self, base: &G1, k: &[FrRepr]) -> Result<Vec<G1>, GpuError> {
        use pairing::CurveProjective;
        let window = G1::recommended_wnaf_for_num_scalars(k.len()) as i32;
        let b = marshal::g1(base);
        let kk: Vec<bls_fr_repr> = k.iter().map(|r| bls_fr_repr { l: r.0 }).collect();
        let mut out = vec![b; k.len()];
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_g1_wnaf_fixed_base_batch(*ctx, &b, window, kk.as_ptr(), out.as_mut_ptr(), kk.len()) }, *ctx)?;
        Ok(out.iter().map(marshal::g1_back).collect())
    }

    /// `let q = Q.prepare(); ps.iter().map(|p| Bls12::pairing(p, Q))` with the one `G2Prepared` staged on the device.
    pub fn pairing_shared_q(&self, p: &[G1Affine], q: &G2Affine) -> Result<Vec<Fq12>, GpuError> {
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq = marshal::g2_affine(q);
        let mut prepared: Box<bls_g2_prepared> = Box::new(unsafe { std::mem::zeroed() });
        let mut out = vec![marshal::FQ12_ZERO; p.len()];
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_g2_prepare_batch(*ctx, &qq, &mut *prepared, 1) }, *ctx)?;
        check(unsafe { bls_pairing_shared_q_batch(*ctx, pp.as_ptr(), &*prepared, out.as_mut_ptr(), pp.len()) }, *ctx)?;
        Ok(out.iter().map(marshal::fq12_back).collect())
    }

    /// `G1Compressed::into_affine` for every 48-byte encoding: `Err(code)` carries the reference's `GroupDecodingError`
    /// as the status byte of include/pairing_b200.h (BLS_DEC_*).
    pub fn g1_decode_compressed(&self, bytes: &[u8]) -> Result<Vec<Result<G1Affine, u8>>, GpuError> {
        assert_eq!(bytes.len() % 48, 0);
        let n = bytes.len() / 48;
        let mut out = vec![bls_g1_affine { x: marshal::FQ_ZERO, y: marshal::FQ_ZERO, infinity: 0 }; n];
        let mut status = vec![0u8; n];
        let ctx = self.ctx.lock().unwrap();
        check(unsafe { bls_g1_decode_batch(*ctx, bytes.as_ptr(), 1, 1, out.as_mut_ptr(), status.as_mut_ptr(), n) }, *ctx)?;
        Ok(out.iter().zip(status).map(|(a, s)| if s == 0 { Ok(marshal::g1_affine_back(a)) } else { Err(s) }).collect())
    }
    // g2_wnaf_mul_batch, g2_batch_normalization, g2_prepare_batch, miller_loop_batch, the other encodings: same pattern.
}

impl Drop for Gpu {
    fn drop(&mut self) { unsafe { bls_ctx_destroy(*self.ctx.lock().unwrap()) } }
}

fn strerror(e: i32) -> String {
    unsafe { std::ffi::CStr::from_ptr(bls_strerror(e)).to_string_lossy().into_owned() }
}
fn check(rc: i32, ctx: *mut bls_ctx) -> Result<(), GpuError> {
    if rc == BLS_OK { return Ok(()); }
    let detail = unsafe { std::ffi::CStr::from_ptr(bls_ctx_last_error(ctx)).to_string_lossy().into_owned() };
    Err(GpuError(rc, format!("{} [{}]", strerror(rc), detail)))
}

/// Explicit field-by-field copies between the crate's types and the `#[repr(C)]` mirrors.  `Fq` is
/// `Fq(FqRepr([u64; 6]))` in Montgomery form (fq.rs:699-700), which is exactly `bls_fq`.
mod marshal {
    use super::ffi::*;
    use pairing::bls12_381::*;
    pub const FQ_ZERO: bls_fq = bls_fq { l: [0; 6] };
    pub const FQ2_ZERO: bls_fq2 = bls_fq2 { c0: FQ_ZERO, c1: FQ_ZERO };
    pub const FQ6_ZERO: bls_fq6 = bls_fq6 { c0: FQ2_ZERO, c1: FQ2_ZERO, c2: FQ2_ZERO };
    pub const FQ12_ZERO: bls_fq12 = bls_fq12 { c0: FQ6_ZERO, c1: FQ6_ZERO };
    // In the fork these use the crate-private accessors `Fq::mont_limbs()` / `Fq::from_mont_limbs()`
    // (two one-line additions to fq.rs) so that no Montgomery conversion happens at the boundary.
    pub fn fq(x: &Fq) -> bls_fq { bls_fq { l: x.mont_limbs() } }
    pub fn fq_back(x: &bls_fq) -> Fq { Fq::from_mont_limbs(x.l) }
    pub fn fq2(x: &Fq2) -> bls_fq2 { bls_fq2 { c0: fq(&x.c0), c1: fq(&x.c1) } }
    pub fn fq2_back(x: &bls_fq2) -> Fq2 { Fq2 { c0: fq_back(&x.c0), c1: fq_back(&x.c1) } }
    pub fn fq6(x: &Fq6) -> bls_fq6 { bls_fq6 { c0: fq2(&x.c0), c1: fq2(&x.c1), c2: fq2(&x.c2) } }
    pub fn fq6_back(x: &bls_fq6) -> Fq6 { Fq6 { c0: fq2_back(&x.c0), c1: fq2_back(&x.c1), c2: fq2_back(&x.c2) } }
    pub fn fq12(x: &Fq12) -> bls_fq12 { bls_fq12 { c0: fq6(&x.c0), c1: fq6(&x.c1) } }
    pub fn fq12_back(x: &bls_fq12) -> Fq12 { Fq12 { c0: fq6_back(&x.c0), c1: fq6_back(&x.c1) } }
    pub fn g1_affine(p: &G1Affine) -> bls_g1_affine { bls_g1_affine { x: fq(&p.x), y: fq(&p.y), infinity: p.infinity as u64 } }
    pub fn g2_affine(p: &G2Affine) -> bls_g2_affine { bls_g2_affine { x: fq2(&p.x), y: fq2(&p.y), infinity: p.infinity as u64 } }
    pub fn g1_affine_back(p: &bls_g1_affine) -> G1Affine { G1Affine { x: fq_back(&p.x), y: fq_back(&p.y), infinity: p.infinity != 0 } }
    pub fn g1(p: &G1) -> bls_g1 { bls_g1 { x: fq(&p.x), y: fq(&p.y), z: fq(&p.z) } }
    pub fn g1_back(p: &bls_g1) -> G1 { G1 { x: fq_back(&p.x), y: fq_back(&p.y), z: fq_back(&p.z) } }
}


/// Several GPUs of one node behind one call: `Engine::miller_loop` takes all pairs of a product at once
/// (src/bls12_381/mod.rs:40-102), so the sharding, the 576-byte exchange and the single final
/// exponentiation happen inside the library (pairing_b200/csrc/mgpu.cu).
pub struct MultiGpu { m: Mutex<*mut bls_mgpu> }
unsafe impl Send for MultiGpu {}
unsafe impl Sync for MultiGpu {}

impl MultiGpu {
    pub fn new(devices: &[i32]) -> Result<MultiGpu, GpuError> {
        let mut err = 0;
        let m = unsafe { bls_mgpu_create(devices.as_ptr(), devices.len() as i32, &mut err) };
        if m.is_null() { return Err(GpuError(err, strerror(err))); }
        Ok(MultiGpu { m: Mutex::new(m) })
    }

    /// One `miller_loop` over all pairs, sharded over the devices.
    pub fn multi_miller_loop(&self, p: &[G1Affine], q: &[G2Affine]) -> Result<Fq12, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq: Vec<bls_g2_affine> = q.iter().map(marshal::g2_affine).collect();
        let mut out = marshal::FQ12_ZERO;
        let m = self.m.lock().unwrap();
        let rc = unsafe { bls_mgpu_multi_miller_loop(*m, pp.as_ptr(), qq.as_ptr(), p.len(), &mut out) };
        if rc != 0 { return Err(GpuError(rc, strerror(rc))); }
        Ok(marshal::fq12_back(&out))
    }

    /// `final_exponentiation(miller_loop(pairs))` over all devices.
    pub fn pairing_product(&self, p: &[G1Affine], q: &[G2Affine]) -> Result<Option<Fq12>, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq: Vec<bls_g2_affine> = q.iter().map(marshal::g2_affine).collect();
        let (mut out, mut some) = (marshal::FQ12_ZERO, 0u8);
        let m = self.m.lock().unwrap();
        let rc = unsafe { bls_mgpu_pairing_product(*m, pp.as_ptr(), qq.as_ptr(), p.len(), &mut out, &mut some) };
        if rc != 0 { return Err(GpuError(rc, strerror(rc))); }
        Ok(if some != 0 { Some(marshal::fq12_back(&out)) } else { None })
    }

    /// Independent pairings, sharded over the devices (no exchange step).
    pub fn pairing_batch(&self, p: &[G1Affine], q: &[G2Affine]) -> Result<Vec<Fq12>, GpuError> {
        assert_eq!(p.len(), q.len());
        let pp: Vec<bls_g1_affine> = p.iter().map(marshal::g1_affine).collect();
        let qq: Vec<bls_g2_affine> = q.iter().map(marshal::g2_affine).collect();
        let mut out = vec![marshal::FQ12_ZERO; p.len()];
        let m = self.m.lock().unwrap();
        let rc = unsafe { bls_mgpu_pairing_batch(*m, pp.as_ptr(), qq.as_ptr(), out.as_mut_ptr(), p.len()) };
        if rc != 0 { return Err(GpuError(rc, strerror(rc))); }
        Ok(out.iter().map(marshal::fq12_back).collect())
    }

    pub fn g1_wnaf_mul_batch(&self, bases: &[G1], k: &[FrRepr]) -> Result<Vec<G1>, GpuError> {
        assert_eq!(bases.len(), k.len());
        let bb: Vec<bls_g1> = bases.iter().map(marshal::g1).collect();
        let kk: Vec<bls_fr_repr> = k.iter().map(|r| bls_fr_repr { l: r.0 }).collect();
        let mut out = bb.clone();
        let m = self.m.lock().unwrap();
        let rc = unsafe { bls_mgpu_g1_wnaf_mul_batch(*m, bb.as_ptr(), kk.as_ptr(), out.as_mut_ptr(), bb.len()) };
        if rc != 0 { return Err(GpuError(rc, strerror(rc))); }
        Ok(out.iter().map(marshal::g1_back).collect())
    }

    pub fn g2_wnaf_mul_batch(&self, bases: &[G2], k: &[FrRepr]) -> Result<Vec<G2>, GpuError> {
        assert_eq!(bases.len(), k.len());
        let bb: Vec<bls_g2> = bases.iter().map(marshal::g2).collect();
        let kk: Vec<bls_fr_repr> = k.iter().map(|r| bls_fr_repr { l: r.0 }).collect();
        let mut out = bb.clone();
        let m = self.m.lock().unwrap();
        let rc = unsafe { bls_mgpu_g2_wnaf_mul_batch(*m, bb.as_ptr(), kk.as_ptr(), out.as_mut_ptr(), bb.len()) };
        if rc != 0 { return Err(GpuError(rc, strerror(rc))); }
        Ok(out.iter().map(marshal::g2_back).collect())
    }
}

impl Drop for MultiGpu {
    fn drop(&mut self) { unsafe { bls_mgpu_destroy(*self.m.lock().unwrap()) } }
}

"""GPU tests at BASELINE.json's sizes (2^16 pairings, 2^18 scalar multiplications), on device-resident data
through the `*_dev` C-ABI entry points (pairing_b200.device.DeviceEngine).  At these sizes the oracle
checks a stride sample bit-exactly, and size-independent properties of the domain cover every element:

  * pairing == final_exponentiation(miller_loop)                       (lib.rs:101-109)
  * multi_miller_loop == product of the single Miller values            (tests/engine.rs:50-91)
  * bilinearity e([a]P, Q) == e(P, [a]Q)                                (tests/engine.rs:93-126)
  * wNAF == double-and-add after batch normalisation                    (tests/curve.rs:68-92)
  * batch_normalization is idempotent and == per-point into_affine      (tests/curve.rs:347-388)
"""
import numpy as np
import pytest
import torch

import bench
import oracle_lib as o
from pairing_b200 import _native as nat

pytestmark = pytest.mark.gpu
TH = o.default_threads()
N_PAIR = 1 << 16
N_WNAF = 1 << 18


def _np(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.fixture(scope="module")
def eng():
    from pairing_b200.device import DeviceEngine
    return DeviceEngine(device=0)


@pytest.fixture(scope="module")
def data(eng):
    """(pa, qa, g1_jac, scalars): N_WNAF subgroup points made by the engine's own wNAF + normalisation kernels,
    first validated against the oracle on a sample (so that the properties below start from known-good points)."""
    pa, qa, g1_jac, ks = bench.make_inputs(eng, N_WNAF, 0x7E57, torch, np)
    g1, g2 = o.generators()
    idx = np.arange(0, N_WNAF, N_WNAF // 64)
    k1 = _np(bench_scalars(eng, N_WNAF, 0x7E57 ^ 0x1111))[idx]
    want = o.g1_into_affine(o.g1_op("wnaf", np.repeat(o.g1_from_affine(g1), len(idx), 0), k=k1, threads=TH))
    assert np.array_equal(_np(pa)[idx], want)
    k2 = _np(bench_scalars(eng, N_WNAF, 0x7E57 ^ 0x2222))[idx]
    want = o.g2_into_affine(o.g2_op("wnaf", np.repeat(o.g2_from_affine(g2), len(idx), 0), k=k2, threads=TH))
    assert np.array_equal(_np(qa)[idx], want)
    return pa, qa, g1_jac, ks


def bench_scalars(eng, n, seed):
    """the scalar stream bench.make_inputs uses (SplitMix64, < 2^254)"""
    with np.errstate(over="ignore"):
        i = np.arange(1, 4 * n + 1, dtype=np.uint64)
        z = np.uint64(seed & (2**64 - 1)) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    k = z.reshape(n, 4)
    k[:, 3] &= np.uint64((1 << 62) - 1)
    k[:, 0] |= np.uint64(2)
    return torch.from_numpy(k.view(np.int64)).to(eng.device)


def test_pairing_full_batch(eng, data):
    pa, qa = data[0][:N_PAIR].clone(), data[1][:N_PAIR].clone()     # clone: the fixture is shared
    pa[5, 12] = 1                                     # an infinity P and an infinity Q inside the batch
    qa[9, 24] = 1
    gt = eng.pairing(pa, qa)
    ml = eng.miller_loop_batch(pa, qa)
    fe, ok = eng.final_exponentiation(ml)
    assert bool(ok.all()) and torch.equal(gt, fe)
    idx = np.r_[np.arange(0, N_PAIR, N_PAIR // 128), 5, 9, N_PAIR - 1]
    assert np.array_equal(_np(gt)[idx], o.pairing(_np(pa)[idx], _np(qa)[idx], TH))
    head = slice(0, 8192)                             # and every output of the first 2^13 pairs, bit for bit
    assert np.array_equal(_np(gt)[head], o.pairing(_np(pa)[head], _np(qa)[head], TH))
    one = np.zeros(72, dtype=np.uint64); one[:6] = _np(data[2])[0, 12:18]
    assert np.array_equal(_np(gt)[5], one) and np.array_equal(_np(gt)[9], one)
    # multi-Miller over the whole batch == product of the single Miller values (two different kernels)
    mm = eng.multi_miller_loop(pa, qa)
    assert torch.equal(mm, eng.fq12_product(ml))
    # the same through the host-buffer entry point (what the Rust shim calls)
    sub = slice(0, 1000)
    assert np.array_equal(eng.ctx.multi_miller_loop(_np(pa)[sub], _np(qa)[sub]), _np(eng.fq12_product(ml[sub].contiguous())))


def test_multi_miller_ragged_trip_counts(eng, data):
    """The multi-pairing kernels multiply the lines of two pairs before touching the accumulator; with T lane pairs a
    lane pair runs ceil(n / T) trips, so n picks the shape: an odd trip count (a single line left over), a last trip
    that only some lane pairs fill, infinity members on both slots.  == product of the single Miller values, for the
    steps-on-the-fly kernel (device buffers) and the G2Prepared kernel (host buffers)."""
    T = eng.ctx.sm_count * 2 * 64                     # lane pairs of a full grid (kernels_pair.cu: mm_threads)
    for n in (T + 1, 2 * T - 3, 2 * T + 64, 3 * T - 1000):
        pa, qa = data[0][:n].clone(), data[1][:n].clone()
        pa[3, 12] = 1; qa[T + 3 if n > T + 3 else 0, 24] = 1; pa[n - 1, 12] = 1
        ml = eng.miller_loop_batch(pa, qa)
        assert torch.equal(eng.multi_miller_loop(pa, qa), eng.fq12_product(ml)), n
    n = T + 70                                        # prepared coefficients: 19 592 B per pair, keep it small
    pa, qa = data[0][:n].clone(), data[1][:n].clone()
    pa[1, 12] = 1; qa[T + 1, 24] = 1
    want = _np(eng.fq12_product(eng.miller_loop_batch(pa, qa)))
    prep = eng.g2_prepare(qa)
    assert np.array_equal(eng.ctx.multi_miller_loop_prepared(_np(pa), _np(prep)), want)


def test_bilinearity_full_batch(eng, data):
    """e([a]P, Q) == e(P, [a]Q) for 2^16 independent (P, Q, a): wNAF (G1 and G2), batch normalisation, pairing."""
    pa, qa, g1_jac, ks = data
    n = N_PAIR
    a = ks[:n].contiguous()
    one = g1_jac[:1, 12:18]
    pj = g1_jac[:n].contiguous()
    qj = torch.zeros((n, nat.W_G2), dtype=torch.int64, device=eng.device)
    qj[:, :24] = qa[:n, :24]; qj[:, 24:30] = one
    ap = eng.g1_batch_normalization_(eng.g1_wnaf_mul(pj, a))
    aq = eng.g2_batch_normalization_(eng.g2_wnaf_mul(qj, a))
    lhs = eng.pairing(eng.jacobian_to_affine_rows(ap, 6), qa[:n].contiguous())
    rhs = eng.pairing(pa[:n].contiguous(), eng.jacobian_to_affine_rows(aq, 12))
    assert torch.equal(lhs, rhs)
    assert not torch.equal(lhs[0], lhs[1])            # not degenerate


def test_wnaf_equals_mul_assign_full_batch(eng, data):
    """Wnaf output and double-and-add output are different Jacobian triples of the same point: equal after
    batch normalisation; the wNAF triples themselves match the oracle bit-exactly on a sample."""
    pa, qa, g1_jac, ks = data
    n = N_WNAF
    two = torch.zeros((n, 4), dtype=torch.int64, device=eng.device); two[:, 0] = 3
    bases = eng.g1_wnaf_mul(g1_jac[:n].contiguous(), two, 2)          # non-normalised (Z != 1) bases = 3 * P_i
    k = ks[:n].clone()
    edges = [0, 1, 2, 3, 1 << 33, (1 << 34) - 1, 1 << 129, (1 << 130) - 1, 1 << 130]    # window thresholds, ec.rs:895-905
    ek = np.array([[(e >> (64 * j)) & (2**64 - 1) for j in range(4)] for e in edges], dtype=np.uint64)
    k[:len(edges)] = torch.from_numpy(ek.view(np.int64)).to(eng.device)
    bases[40] = 0; bases[40, 6:12] = g1_jac[0, 12:18]                   # an infinity base (0, 1, 0)
    w = eng.g1_wnaf_mul(bases, k)
    idx = np.r_[np.arange(0, 64), np.arange(64, n, n // 256)]
    assert np.array_equal(_np(w)[idx], o.g1_op("wnaf", _np(bases)[idx], k=_np(k)[idx], threads=TH))
    sub = slice(0, 1 << 14)                                             # double-and-add kernel on a 2^14 slice
    da = torch.from_numpy(eng.ctx.g1_mul(_np(bases[sub]), _np(k[sub])).view(np.int64)).to(eng.device)
    wn = eng.g1_batch_normalization_(w[sub].clone())
    dn = eng.g1_batch_normalization_(da)
    live = (wn[:, 12:18] != 0).any(dim=1)
    assert torch.equal(wn[live], dn[live]) and torch.equal(live, (dn[:, 12:18] != 0).any(dim=1))
    # normalisation: idempotent, and equal to the per-point conversion
    assert torch.equal(eng.g1_batch_normalization_(wn.clone()), wn)
    aff = eng.ctx.g1_into_affine(_np(w[:4096]))
    assert np.array_equal(aff[:, :12], _np(wn[:4096])[:, :12])


def test_g2_wnaf_and_prepare_sample(eng, data):
    pa, qa, g1_jac, ks = data
    n = 1 << 15
    one = g1_jac[:1, 12:18]
    qj = torch.zeros((n, nat.W_G2), dtype=torch.int64, device=eng.device)
    qj[:, :24] = qa[:n, :24]; qj[:, 24:30] = one
    three = torch.zeros((n, 4), dtype=torch.int64, device=eng.device); three[:, 0] = 3
    bases = eng.g2_wnaf_mul(qj, three, 2)
    w = eng.g2_wnaf_mul(bases, ks[:n].contiguous())
    idx = np.arange(0, n, n // 128)
    assert np.array_equal(_np(w)[idx], o.g2_op("wnaf", _np(bases)[idx], k=_np(ks[:n])[idx], threads=TH))
    wn = eng.g2_batch_normalization_(w.clone())
    prep = eng.g2_prepare(eng.jacobian_to_affine_rows(wn, 12))
    assert np.array_equal(_np(prep)[idx[:16]], o.g2_prepare(o.g2_into_affine(_np(w)[idx[:16]]), TH))
    # Miller loop from the stored coefficients == Miller loop with the steps on the fly
    p = pa[:n].contiguous()
    assert torch.equal(eng.miller_loop_prepared_batch(p, prep), eng.miller_loop_batch(p, eng.jacobian_to_affine_rows(wn, 12)))


def test_fixed_base_and_gt_pow_device_paths(eng, data):
    """Wnaf::base(g, n).scalar(s) for 2^18 scalars against ONE shared table (window 14 for G1 at this n) == the per-point
    wNAF after normalisation; e(P, Q)^k device path == host path; both sampled against the oracle."""
    pa, qa, g1_jac, ks = data
    from pairing_b200 import engine
    n = N_WNAF
    w = engine.G1.recommended_wnaf_for_num_scalars(n)
    assert w == 16                                                          # n > 62569 (ec.rs:907-921)
    base = g1_jac[7:8].contiguous()
    three = torch.zeros((1, 4), dtype=torch.int64, device=eng.device); three[0, 0] = 3
    base = eng.g1_wnaf_mul(base, three, 2)                                   # non-normalised base
    table = eng.wnaf_table(base, w)
    fixed = eng.wnaf_fixed_base(table, w, ks[:n].contiguous())
    idx = np.arange(0, n, n // 32)
    want = o.g1_op("wnaf", np.repeat(_np(base), len(idx), 0), k=_np(ks[:n])[idx], window=w, threads=TH)
    assert np.array_equal(_np(fixed)[idx], want)
    per_point = eng.g1_wnaf_mul(base.repeat(n, 1).contiguous(), ks[:n].contiguous())       # window 4 per scalar
    assert not torch.equal(fixed, per_point)                                 # different Jacobian representatives ...
    assert torch.equal(eng.g1_batch_normalization_(fixed.clone()), eng.g1_batch_normalization_(per_point))   # ... of the same points
    m_ = 4096
    gt = eng.pairing(pa[:m_].contiguous(), qa[:m_].contiguous())
    pw = eng.fq12_pow(gt, ks[:m_].contiguous())
    assert np.array_equal(_np(pw)[:64], eng.ctx.fq12_pow(_np(gt)[:64], _np(ks[:m_])[:64]))
    # e([k]P, Q) == e(P, Q)^k on the whole slice
    pj = g1_jac[:m_].contiguous()
    kp = eng.g1_batch_normalization_(eng.g1_wnaf_mul(pj, ks[:m_].contiguous()))
    assert torch.equal(eng.pairing(eng.jacobian_to_affine_rows(kp, 6), qa[:m_].contiguous()), pw)

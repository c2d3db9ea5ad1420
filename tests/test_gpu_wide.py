"""GPU parity of the latency path (kernels_wide.cu: one WARP per element), of the multi-pairing tail (block-level
partial products + k_pair_product_tail) and of the multi-device entry points (mgpu.cu), all through the C ABI and
against the CPU oracle.  The same inputs also run with the latency path disabled, so that the lane-pair throughput
kernels stay covered at small n."""
import numpy as np
import pytest

import bls_model as m
import datagen as dg
import oracle_lib as o

pytestmark = pytest.mark.gpu
TH = o.default_threads()


def eq(a, b):
    assert a.shape == b.shape and np.array_equal(a, b), "first mismatch at row %s" % (np.argwhere((a != b).any(axis=1))[:3].tolist(),)


@pytest.fixture(params=["wide", "lanepair"])
def path_ctx(ctx, request):
    """the session context with the latency path on (default limits) or off"""
    if request.param == "lanepair":
        ctx.set_latency_path_limits(0, 0)
    else:
        ctx.set_latency_path_limits(2560, 2560)
    yield ctx
    ctx.set_latency_path_limits(2560, 2560)


def test_single_pairing_relic_kat(path_ctx):
    """BASELINE configs[0] / bls12_381/tests/mod.rs:5-53 on both kernels: e(g1, g2) is the RELIC value"""
    import os
    g1, g2 = o.generators()
    kat = open(os.path.join(os.path.dirname(__file__), "golden", "relic_pairing_g1g2.bin"), "rb").read()
    assert path_ctx.pairing(g1, g2).tobytes() == kat


@pytest.mark.parametrize("n", [1, 2, 7, 33, 129, 1000])
def test_pairing_small_batches(path_ctx, n):
    """the crate's bench_pairing_full shape (benches/bls12_381/mod.rs:91-107: 1000 pairings) and ragged sizes,
    with infinity members (mod.rs:49-54 -> one)"""
    p = dg.g1_affine_points(n, 301, infinity_at=(3,) if n > 4 else ())
    q = dg.g2_affine_points(n, 302, infinity_at=(5, 6) if n > 8 else ())
    eq(path_ctx.pairing(p, q), o.pairing(p, q, TH))


@pytest.mark.parametrize("n", [1, 5, 130, 1000])
def test_pairing_on_projective_inputs(path_ctx, n):
    """Engine::pairing with `Into<G1Affine>` / `Into<G2Affine>` arguments, as the crate's bench_pairing_full calls it: the two
    into_affine conversions fused in front (non-normalised Z, Z == one and infinity members)"""
    pj = dg.g1_points(n, 320, infinity_at=(3,) if n > 4 else ())
    qj = dg.g2_points(n, 321, infinity_at=(4,) if n > 4 else ())
    if n > 8:                                          # a few already-normalised points (Z == one): the into_affine shortcut, ec.rs:592-596
        pj[6:8] = o.g1_batch_normalization(pj[6:8]); qj[7:9] = o.g2_batch_normalization(qj[7:9])
    want = o.pairing(o.g1_into_affine(pj), o.g2_into_affine(qj), TH)
    eq(path_ctx.pairing_projective(pj, qj), want)


def test_final_exponentiation_small_batches(path_ctx):
    f = dg.rand_field(70, 12, 303)
    f[4] = 0                       # None (mod.rs:107-108)
    f[5:40] = o.miller_loop(dg.g1_affine_points(35, 304), dg.g2_affine_points(35, 305), TH)
    # operands whose easy part lands on one (or on another element with a zero c1.c0 coefficient): the compressed
    # squarings of exp_by_x cannot be decompressed there and the lane-pair kernel has to take its uncompressed path --
    # inside warps whose other lane pairs stay on the compressed one
    one = np.zeros(72, dtype=np.uint64); one[:6] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)
    f[41] = one
    f[42, 36:] = 0                 # an element of Fq6
    f[43, 12:] = 0                 # an element of Fq2
    f[44] = 0; f[44, :6] = np.array(m.limbs64(m.Q - m.MONT_R), dtype=np.uint64)       # minus one
    f[45, :36] = 0                 # c0 = 0: f = c1 w, f^(q^6 - 1) = -1
    want, wok = o.final_exponentiation(f, TH)
    assert np.array_equal(want[41], one) and np.array_equal(want[42], one) and np.array_equal(want[45], one)
    got, gok = path_ctx.final_exponentiation(f)
    eq(got, want)
    assert np.array_equal(gok, wok) and gok[4] == 0
    got1, ok1 = path_ctx.final_exponentiation(f[7:8])          # a batch of one
    eq(got1, want[7:8])


@pytest.mark.parametrize("n", [0, 1, 2, 7, 129, 1000, 20000])
def test_pairing_product(path_ctx, n):
    ctx = path_ctx          # up to 1024 pairs the Miller loops run one WARP per pair (k_wide_miller) when the latency path is on
    """final_exponentiation(miller_loop(n pairs)) in one call: block-level partial products, tail kernel, final
    exponentiation on the warp-cooperative engine (mod.rs:40-160)"""
    base = 512
    p = dg.g1_affine_points(min(max(n, 1), base), 306, infinity_at=(2,) if n > 4 else ())
    q = dg.g2_affine_points(min(max(n, 1), base), 307, infinity_at=(4,) if n > 4 else ())
    reps = (n + base - 1) // base if n else 1
    p, q = np.tile(p, (reps, 1))[:n], np.tile(q, (reps, 1))[:n]
    mm = o.multi_miller_product(p, q, TH)
    want, wok = o.final_exponentiation(mm, 1)
    got, ok = ctx.pairing_product(p, q)
    eq(got, want)
    assert ok == bool(wok[0])
    eq(ctx.multi_miller_loop(p, q), mm)


@pytest.mark.parametrize("n", [1, 3, 200])
def test_miller_loop_small_batches(path_ctx, n):
    """Engine::miller_loop for n independent pairs (mod.rs:40-102) on both paths: one warp per pair (k_wide_miller) and one
    lane pair per pair (k_pair_miller<false>); pairs with a point at infinity give one"""
    p = dg.g1_affine_points(max(n, 8), 311, infinity_at=(1,))[:n]
    q = dg.g2_affine_points(max(n, 8), 312, infinity_at=(2,))[:n]
    eq(path_ctx.miller_loop(p, q), o.miller_loop(p, q, TH))


@pytest.mark.parametrize("count", [1, 2, 31, 32, 33, 64, 65, 300, 1000])
def test_product_tail(ctx, count):
    """k_pair_product_tail: strided products + shared-memory tree over `count` factors, with and without the final exponentiation"""
    import torch
    from pairing_b200.device import DeviceEngine
    eng = DeviceEngine(ctx=ctx)
    f = o.miller_loop(dg.g1_affine_points(8, 308), dg.g2_affine_points(8, 309), TH)
    f = np.tile(f, ((count + 7) // 8, 1))[:count].copy()
    f[count // 2] = dg.rand_field(1, 12, 310 + count)[0]
    want = f[:1].copy()
    for i in range(1, count):
        want, _ = o.fq12_op("mul", want, f[i:i + 1])
    d = torch.from_numpy(f.view(np.int64)).to(eng.device)
    got, _ = eng.fq12_product_tail(d, final_exp=False)
    eq(got.cpu().numpy().view(np.uint64), want)
    got, ok = eng.fq12_product_tail(d, final_exp=True)
    wfe, wok = o.final_exponentiation(want, 1)
    eq(got.cpu().numpy().view(np.uint64), wfe)
    assert int(ok.item()) == int(wok[0])


def test_product_tail_zero_is_none(ctx):
    import torch
    from pairing_b200.device import DeviceEngine
    eng = DeviceEngine(ctx=ctx)
    f = dg.rand_field(5, 12, 311)
    f[3] = 0
    got, ok = eng.fq12_product_tail(torch.from_numpy(f.view(np.int64)).to(eng.device), final_exp=True)
    assert int(ok.item()) == 0 and not got.cpu().numpy().any()


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_multi_device_entry_points(ctx, ndev):
    """bls_mgpu_*: ONE n-pair product / batch sharded over ndev devices inside the library, bit-equal to the oracle"""
    if ndev > _device_count():
        pytest.skip("needs %d GPUs" % ndev)
    import pairing_b200._native as nat
    with nat.MultiGpu(ndev) as mg:
        assert mg.device_count == ndev
        for n in (0, 1, ndev - 1, 5, 700, 9000):
            base = 256
            p = dg.g1_affine_points(base, 312, infinity_at=(2,))
            q = dg.g2_affine_points(base, 313, infinity_at=(4,))
            reps = max(1, (n + base - 1) // base)
            pn, qn = np.tile(p, (reps, 1))[:n], np.tile(q, (reps, 1))[:n]
            mm = o.multi_miller_product(pn, qn, TH)
            eq(mg.multi_miller_loop(pn, qn), mm)
            want, wok = o.final_exponentiation(mm, 1)
            got, ok = mg.pairing_product(pn, qn)
            eq(got, want)
            assert ok == bool(wok[0])
        ph = mg.last_phase_ms()
        assert ph["total_ms"] > 0
        n = 301
        p, q = dg.g1_affine_points(n, 314, infinity_at=(0,)), dg.g2_affine_points(n, 315)
        eq(mg.pairing(p, q), o.pairing(p, q, TH))
        bases, k = dg.g1_points(n, 316), dg.rand_scalars(n, 317)
        eq(mg.g1_wnaf_mul(bases, k), o.g1_op("wnaf", bases, k=k, threads=TH))
        bases2 = dg.g2_points(64, 318)
        eq(mg.g2_wnaf_mul(bases2, k[:64]), o.g2_op("wnaf", bases2, k=k[:64], threads=TH))


def test_entry_points_leave_the_callers_device_alone(ctx):
    """a context on another device must not switch the process's current CUDA device"""
    import torch
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pairing_b200._native as nat
    torch.cuda.set_device(0)
    with nat.Context(1) as c1:
        g1, g2 = o.generators()
        c1.pairing(g1, g2)
        assert torch.cuda.current_device() == 0


# ---- the lane-pair tower (pair_tower.cuh) op by op, on edge operands -----------------------
def _edge_rows(degree, seed):
    """rows whose Fq coefficients cycle through edge values -- 0, 1, q-1, R, (q+-1)/2, and NON-canonical representatives
    in (q, 2q] (the relaxed range the lane-pair code works in) -- followed by random rows"""
    q = m.Q
    vals = [0, 1, 2, q - 1, q - 2, m.MONT_R, (q - 1) // 2, (q + 1) // 2, 1 << 380, q, q + 1, 2 * q - 1, 2 * q, q + m.MONT_R]
    rows = []
    for r in range(len(vals) * 2):
        rows.append([x for k in range(degree) for x in m.limbs64(vals[(r + 5 * k) % len(vals)])])
    e = np.array(rows, dtype=np.uint64)
    return np.concatenate([e, dg.rand_field(96, degree, seed)])


def _canon(a, degree):
    """reduce every Fq coefficient mod q (what the oracle, which expects canonical inputs, is given)"""
    out = a.copy()
    for r in range(a.shape[0]):
        for k in range(degree):
            v = m.from_limbs64(a[r, 6 * k:6 * k + 6])
            if v >= m.Q:
                out[r, 6 * k:6 * k + 6] = np.array(m.limbs64(v % m.Q), dtype=np.uint64)
    return out


@pytest.mark.parametrize("degree,op", [(2, x) for x in ("add", "sub", "mul", "sqr", "neg", "dbl", "inv", "mul_nonres", "frob1")] +
                         [(6, x) for x in ("add", "sub", "mul", "sqr", "neg", "inv", "mul_nonres", "frob1", "frob2", "frob3", "mul_by_01", "mul_by_1")] +
                         [(12, x) for x in ("mul", "sqr", "inv", "conj", "frob1", "frob2", "frob3", "mul_by_014")])
def test_lane_pair_tower_ops(ctx, degree, op):
    a, b = _edge_rows(degree, 400 + degree), _edge_rows(degree, 500 + degree)[::-1].copy()
    binary = op in ("add", "sub", "mul", "mul_by_01", "mul_by_1", "mul_by_014")
    fn = {2: o.fq2_op, 6: o.fq6_op, 12: o.fq12_op}[degree]
    want, wok = fn(op, _canon(a, degree), _canon(b, degree) if binary else None)
    got, gok = ctx.pair_field_op(degree, op, a, b if binary else None)
    eq(got, want)
    assert np.array_equal(gok, wok)


def test_lane_pair_line_pair_product(ctx):
    """p12_mul_by_line_pair (the multi-pairing kernel's sparse x sparse product) == two mul_by_014 (fq12.rs:34-48)"""
    f = _edge_rows(12, 601)
    lm = _edge_rows(12, 602)[::-1].copy()
    fc, lc = _canon(f, 12), _canon(lm, 12)
    def sparse(rows, k):      # (c0, c1, c4) of line k as the b operand of the oracle's mul_by_014: b.c0.c0, b.c0.c1, b.c1.c1
        b = np.zeros_like(rows)
        b[:, 0:12] = rows[:, 36 * k:36 * k + 12]; b[:, 12:24] = rows[:, 36 * k + 12:36 * k + 24]; b[:, 48:60] = rows[:, 36 * k + 24:36 * k + 36]
        return b
    want, _ = o.fq12_op("mul_by_014", fc, sparse(lc, 0))
    want, _ = o.fq12_op("mul_by_014", want, sparse(lc, 1))
    got, _ = ctx.pair_field_op(12, "mul_by_line_pair", f, lm)
    eq(got, want)


def test_lane_pair_cyclotomic_square(ctx):
    """Granger-Scott squaring == Fq12::square on the cyclotomic subgroup (outputs of the final exponentiation and their powers)"""
    f = dg.rand_field(40, 12, 603)
    g, ok = o.final_exponentiation(f, TH)
    assert ok.all()
    want, _ = o.fq12_op("sqr", g)
    got, _ = ctx.pair_field_op(12, "cyclotomic_sqr", g)
    eq(got, want)


def test_one_context_shared_by_several_host_threads(ctx):
    """Every entry point holds the context's lock while it runs on the host (abi_common.cuh: USE_DEVICE), so calls from
    several threads on ONE bls_ctx serialise instead of interleaving their staging buffers and streams: four threads issue
    different calls at once (ctypes releases the GIL) and every result equals the one computed alone."""
    import threading
    p, q = dg.g1_affine_points(3000, 901), dg.g2_affine_points(3000, 902)        # above the latency-path limit: chunked pipeline + lane-pair kernel
    f = o.miller_loop(p[:40], q[:40], TH)
    jobs = {
        "pairing": lambda: ctx.pairing(p, q),
        "pairing_small": lambda: ctx.pairing(p[:33], q[:33]),
        "final_exp": lambda: ctx.final_exponentiation(f)[0],
        "product": lambda: ctx.pairing_product(p[:500], q[:500])[0],
    }
    want = {k: fn() for k, fn in jobs.items()}
    got, errs = {}, []
    def run(k, fn):
        try:
            for _ in range(3):
                got[k] = fn()
        except Exception as e:       # noqa: BLE001 -- reported below
            errs.append((k, e))
    threads = [threading.Thread(target=run, args=kv) for kv in jobs.items()]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not errs, errs
    for k in jobs:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k

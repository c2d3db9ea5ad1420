"""GPU parity: every C-ABI entry point against the CPU oracle on the same seeded inputs, bit-exact.
(Integer/byte work: the bar is equality of every output byte.)"""
import numpy as np
import pytest

import bls_model as m
import datagen as dg
import oracle_lib as o

pytestmark = pytest.mark.gpu
TH = o.default_threads()


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    if not np.array_equal(a, b):
        bad = np.nonzero((a != b).any(axis=1))[0]
        raise AssertionError("%d of %d rows differ; first bad row %d" % (len(bad), a.shape[0], bad[0]))


def _edge_fq():
    vals = [0, 1, 2, m.Q - 1, m.Q - 2, m.MONT_R, (m.Q - 1) // 2, (m.Q + 1) // 2, 1 << 380, (1 << 381) - 1 - (1 << 380)]
    vals = [v % m.Q for v in vals]
    return np.array([m.limbs64(v) for v in vals], dtype=np.uint64)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "sqr", "neg", "dbl", "inv", "from_repr", "into_repr"])
def test_fq_ops(ctx, op):
    a = np.concatenate([_edge_fq(), dg.rand_fq(4096, 1)])
    b = np.concatenate([_edge_fq()[::-1], dg.rand_fq(4096, 2)])
    if op == "from_repr":   # include non-canonical inputs: q, q+1, 2^384-1
        extra = np.array([m.limbs64(m.Q), m.limbs64(m.Q + 1), m.limbs64((1 << 384) - 1)], dtype=np.uint64)
        a = np.concatenate([a, extra]); b = np.concatenate([b, extra])
    binary = op in ("add", "sub", "mul")
    want, wok = o.fq_op(op, a, b if binary else None)
    got, gok = ctx.field_op(1, op, a, b if binary else None)
    eq(got, want)
    assert np.array_equal(gok, wok)


def test_fq_inverse_division_steps(ctx):
    """The Fq inversion of every kernel (pairing_b200/csrc/fp_inv_gcd.cuh: division steps on 30-bit limbs; the reference:
    fq.rs:849-902) against pow(a, -1, q) computed here: operands of every bit length (the batch loop ends at a different
    step for each lane of a warp), +-small values, and 2^15 random ones -- on the thread-per-element path and on lane pairs."""
    import random
    rng = random.Random(0x1F)
    vals = [0, 1, 2, 3, m.Q - 1, m.Q - 2, m.Q - 3, (m.Q + 1) // 2, m.MONT_R, m.MONT_RINV]
    vals += [rng.randrange(1, 1 << k) % m.Q for k in range(1, 382) for _ in range(4)]
    vals += [m.Q - rng.randrange(1, 1 << k) for k in range(1, 64)]
    vals += [rng.randrange(m.Q) for _ in range(1 << 15)]
    a = np.array([m.limbs64(m.to_mont(v)) for v in vals], dtype=np.uint64)
    want = np.array([m.limbs64(m.to_mont(pow(v, -1, m.Q)) if v else 0) for v in vals], dtype=np.uint64)
    got, ok = ctx.field_op(1, "inv", a)
    eq(got, want)
    assert np.array_equal(ok, np.array([1 if v else 0 for v in vals], dtype=ok.dtype))
    # lane pairs: x + 0 u inverts to x^-1 + 0 u through p2_inv (both lanes of a pair run the inversion of the norm x^2)
    a2 = np.zeros((len(vals), 12), dtype=np.uint64); a2[:, :6] = a
    w2 = np.zeros_like(a2); w2[:, :6] = want
    got2, ok2 = ctx.pair_field_op(2, "inv", a2)
    eq(got2, w2)
    assert np.array_equal(ok2, ok)


def test_fq_add_sub_cross_edges(ctx):
    e = _edge_fq()
    a = np.repeat(e, len(e), 0); b = np.tile(e, (len(e), 1))
    for op in ("add", "sub", "mul"):
        eq(ctx.field_op(1, op, a, b)[0], o.fq_op(op, a, b)[0])


@pytest.mark.parametrize("op", ["add", "sub", "mul", "sqr", "neg", "dbl", "inv", "mul_nonres", "frob1"])
def test_fq2_ops(ctx, op):
    a = dg.rand_field(2048, 2, 3); b = dg.rand_field(2048, 2, 4)
    a[0] = 0; a[1, 6:] = 0; a[2, :6] = 0
    binary = op in ("add", "sub", "mul")
    want, wok = o.fq2_op(op, a, b if binary else None)
    got, gok = ctx.field_op(2, op, a, b if binary else None)
    eq(got, want)
    assert np.array_equal(gok, wok)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "sqr", "neg", "inv", "mul_nonres", "frob1", "frob2", "frob3", "mul_by_01", "mul_by_1"])
def test_fq6_ops(ctx, op):
    a = dg.rand_field(512, 6, 5); b = dg.rand_field(512, 6, 6)
    a[0] = 0
    binary = op in ("add", "sub", "mul", "mul_by_01", "mul_by_1")
    want, wok = o.fq6_op(op, a, b if binary else None)
    got, gok = ctx.field_op(6, op, a, b if binary else None)
    eq(got, want)
    assert np.array_equal(gok, wok)


@pytest.mark.parametrize("op", ["mul", "sqr", "inv", "conj", "frob1", "frob2", "frob3", "mul_by_014"])
def test_fq12_ops(ctx, op):
    a = dg.rand_field(256, 12, 7); b = dg.rand_field(256, 12, 8)
    a[0] = 0
    binary = op in ("mul", "mul_by_014")
    want, wok = o.fq12_op(op, a, b if binary else None)
    got, gok = ctx.field_op(12, op, a, b if binary else None)
    eq(got, want)
    assert np.array_equal(gok, wok)


def test_pairing_relic_kat(ctx):
    """src/bls12_381/tests/mod.rs:5-53: e(G1::one(), G2::one()) against the RELIC value."""
    g1, g2 = o.generators()
    got = ctx.pairing(g1, g2)
    want = open(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "relic_pairing_g1g2.bin"), "rb").read()
    assert got.tobytes() == want


def test_g2_prepare(ctx):
    q = dg.g2_affine_points(96, 11, infinity_at=(5,))
    eq(ctx.g2_prepare(q), o.g2_prepare(q, TH))


def test_miller_loop_and_prepared(ctx):
    n = 200
    p = dg.g1_affine_points(n, 12, infinity_at=(3,))
    q = dg.g2_affine_points(n, 13, infinity_at=(7,))
    want = o.miller_loop(p, q, TH)
    eq(ctx.miller_loop(p, q), want)
    eq(ctx.miller_loop_prepared(p, o.g2_prepare(q, TH)), want)
    one = np.zeros(72, dtype=np.uint64); one[:6] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)
    assert np.array_equal(want[3], one) and np.array_equal(want[7], one)


def test_final_exponentiation(ctx):
    f = dg.rand_field(150, 12, 14)
    f[4] = 0                       # None case
    f[5:50] = o.miller_loop(dg.g1_affine_points(45, 15), dg.g2_affine_points(45, 16), TH)
    want, wok = o.final_exponentiation(f, TH)
    got, gok = ctx.final_exponentiation(f)
    eq(got, want)
    assert np.array_equal(gok, wok) and gok[4] == 0 and gok.sum() == len(f) - 1


def test_pairing_batch(ctx):
    n = 300
    p = dg.g1_affine_points(n, 17, infinity_at=(0,))
    q = dg.g2_affine_points(n, 18, infinity_at=(1,))
    eq(ctx.pairing(p, q), o.pairing(p, q, TH))


@pytest.mark.parametrize("n", [0, 1, 2, 7, 129, 1000])
def test_multi_miller_loop(ctx, n):
    inf_p = (2,) if n > 4 else ()
    inf_q = (4,) if n > 4 else ()
    p = dg.g1_affine_points(max(n, 1), 19, infinity_at=inf_p)[:n]
    q = dg.g2_affine_points(max(n, 1), 20, infinity_at=inf_q)[:n]
    want = o.multi_miller_product(p, q, TH)
    if n <= 7:   # the literal shared-accumulator loop of the reference
        eq(o.multi_miller_loop(p, q), want)
    eq(ctx.multi_miller_loop(p, q), want)
    if n:
        eq(ctx.multi_miller_loop_prepared(p, o.g2_prepare(q, TH)), want)


def test_fq12_product(ctx):
    f = dg.rand_field(77, 12, 21)
    want = f[:1].copy()
    for i in range(1, len(f)):
        want = o.fq12_op("mul", want, f[i:i + 1])[0]
    eq(ctx.fq12_product(f), want)


@pytest.mark.parametrize("g2", [False, True])
def test_point_ops(ctx, g2):
    n = 128
    gen = dg.g2_points if g2 else dg.g1_points
    op_o = o.g2_op if g2 else o.g1_op
    op_g = ctx.g2_op if g2 else ctx.g1_op
    to_aff = o.g2_into_affine if g2 else o.g1_into_affine
    a = gen(n, 22, infinity_at=(0, 5)); b = gen(n, 23, infinity_at=(1, 5))
    b[10] = a[10]                                   # equal -> double branch
    b[11] = op_o("negate", a[11:12])[0]             # P + (-P): H = 0 fall-through, z = 0 with garbage x, y
    b[12] = op_o("double", a[12:13])[0]
    for op in ("double", "negate"):
        eq(op_g(op, a), op_o(op, a))
    for op in ("add", "sub"):
        eq(op_g(op, a, b), op_o(op, a, b))
    ba = to_aff(b)
    eq(op_g("add_mixed", a, ba), op_o("add_mixed", a, ba))
    # same-representative mixed addition (z == one) hits the doubling branch
    an = (o.g2_from_affine if g2 else o.g1_from_affine)(ba)
    eq(op_g("add_mixed", an, ba), op_o("add_mixed", an, ba))
    into_g = ctx.g2_into_affine if g2 else ctx.g1_into_affine
    eq(into_g(a), to_aff(a)); eq(into_g(an), to_aff(an))


@pytest.mark.parametrize("g2", [False, True])
def test_wnaf_mul_heuristic_windows(ctx, g2):
    n = 160
    bases = (dg.g2_points if g2 else dg.g1_points)(n, 24, infinity_at=(20,))
    k = dg.rand_scalars(n, 25)
    want = (o.g2_op if g2 else o.g1_op)("wnaf", bases, k=k, threads=TH)
    got = (ctx.g2_wnaf_mul if g2 else ctx.g1_wnaf_mul)(bases, k)
    eq(got, want)


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("w", [2, 3, 4, 5, 6, 7, 8, 10, 13])
def test_wnaf_mul_explicit_window(ctx, g2, w):
    n = 48
    bases = (dg.g2_points if g2 else dg.g1_points)(n, 26 + w)
    k = dg.rand_scalars(n, 27 + w)
    want = (o.g2_op if g2 else o.g1_op)("wnaf", bases, k=k, window=w, threads=TH)
    got = (ctx.g2_wnaf_mul if g2 else ctx.g1_wnaf_mul)(bases, k, window=w)
    eq(got, want)


@pytest.mark.parametrize("g2", [False, True])
def test_mul_assign(ctx, g2):
    n = 64
    bases = (dg.g2_points if g2 else dg.g1_points)(n, 40)
    k = dg.rand_scalars(n, 41)
    eq((ctx.g2_mul if g2 else ctx.g1_mul)(bases, k), (o.g2_op if g2 else o.g1_op)("mul", bases, k=k, threads=TH))


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("n", [1, 3, 100, 5000])
def test_batch_normalization(ctx, g2, n):
    """tests/curve.rs:347-388: batch_normalization == per-point into_affine, with sprinkled
    infinity and already-normalised entries; compared bit-exactly with the reference's sequential
    Montgomery trick."""
    gen = dg.g2_points if g2 else dg.g1_points
    v = gen(n, 42, infinity_at=tuple(i for i in (0, 17, 63) if i < n))
    if n > 30:
        aff = (o.g2_into_affine if g2 else o.g1_into_affine)(v[20:30])
        v[20:30] = (o.g2_from_affine if g2 else o.g1_from_affine)(aff)      # z == one
    want = (o.g2_batch_normalization if g2 else o.g1_batch_normalization)(v)
    got = (ctx.g2_batch_normalization if g2 else ctx.g1_batch_normalization)(v)
    eq(got, want)


def test_imad_peak_runs(ctx):
    for variant in (0, 1, 2):
        macs, ms = ctx.imad_peak(variant, 200)
        assert macs > 1e11 and ms > 0


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("w", [2, 4, 7, 8, 11])
def test_wnaf_fixed_base(ctx, g2, w):
    """Wnaf::base(g, n).scalar(s) (wnaf.rs:93-107, 169-178): one shared table; per scalar the result is the
    reference's wnaf_exp(wnaf_table(g, w), wnaf_form(s, w)) -- the oracle's explicit-window path on the same base."""
    n = 70
    base = (dg.g2_points if g2 else dg.g1_points)(1, 60 + w)
    k = dg.rand_scalars(n, 61 + w)
    want = (o.g2_op if g2 else o.g1_op)("wnaf", np.repeat(base, n, 0), k=k, window=w, threads=TH)
    got = (ctx.g2_wnaf_fixed_base if g2 else ctx.g1_wnaf_fixed_base)(base, w, k)
    eq(got, want)


@pytest.mark.parametrize("g2", [False, True])
def test_wnaf_table_and_builder(ctx, g2):
    """wnaf_table (wnaf.rs:4-15): table[i] = (2i+1) g as the chain of projective additions of 2g; and the
    typestate builder picks the window from the number of scalars (ec.rs:907-921 / 1598-1612)."""
    from pairing_b200 import engine
    curve = engine.G2 if g2 else engine.G1
    op_o = o.g2_op if g2 else o.g1_op
    base = (dg.g2_points if g2 else dg.g1_points)(1, 77)
    table = (ctx.g2_wnaf_table if g2 else ctx.g1_wnaf_table)(base, 6)
    want = [base]
    dbl = op_o("double", base)
    for _ in range(31):
        want.append(op_o("add", want[-1], dbl))
    eq(table, np.concatenate(want))
    for num_scalars in (1, 50, 2000):                       # windows 4, 9, 13 (G1) / 4, 9, 12 (G2)
        w = curve.recommended_wnaf_for_num_scalars(num_scalars)
        k = dg.rand_scalars(12, 78 + num_scalars)
        got = engine.Wnaf(curve, ctx).base(base, num_scalars).scalar(k)
        eq(got, op_o("wnaf", np.repeat(base, 12, 0), k=k, window=w, threads=TH))
    inf = base.copy(); inf[:] = 0
    inf[0, (12 if g2 else 6):(18 if g2 else 12)] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)
    k = dg.rand_scalars(5, 90)
    eq((ctx.g2_wnaf_fixed_base if g2 else ctx.g1_wnaf_fixed_base)(inf, 5, k), op_o("wnaf", np.repeat(inf, 5, 0), k=k, window=5, threads=TH))


def _oracle_fq12_pow(a, k):
    """a_i^(k_i) for 256-bit k from the oracle's u64 pow: prod_j (a^(2^(64 j)))^(limb_j)"""
    out = None
    base = a.copy()
    for j in range(4):
        term = np.concatenate([o.fq12_pow_u64(base[i:i + 1], int(k[i, j])) for i in range(len(a))])
        out = term if out is None else o.fq12_op("mul", out, term)[0]
        if j < 3:
            base = np.concatenate([o.fq12_pow_u64(o.fq12_pow_u64(base[i:i + 1], 1 << 63), 2) for i in range(len(a))])
    return out


def test_fq12_pow_and_bilinearity(ctx):
    """Field::pow (lib.rs:306-324) on Fq12 and the reference's bilinearity check (tests/engine.rs:93-126):
    e([a]P, [b]Q) == e(P, Q)^(ab) computed as (e(P,Q)^a)^b, == e([ab]P, Q)."""
    n = 40
    f = dg.rand_field(n, 12, 50)
    k = dg.rand_scalars(n, 51)                     # includes 0, 1, r-1 and the window-threshold scalars
    eq(ctx.fq12_pow(f, k), _oracle_fq12_pow(f, k))
    p = dg.g1_points(n, 52); q = dg.g2_points(n, 53)
    a = dg.rand_scalars(n, 54, edge_cases=False); b = dg.rand_scalars(n, 55, edge_cases=False)
    pa, qa = ctx.g1_into_affine(p), ctx.g2_into_affine(q)
    ap = ctx.g1_into_affine(ctx.g1_wnaf_mul(p, a)); bq = ctx.g2_into_affine(ctx.g2_wnaf_mul(q, b))
    lhs = ctx.pairing(ap, bq)
    rhs = ctx.fq12_pow(ctx.fq12_pow(ctx.pairing(pa, qa), a), b)
    eq(lhs, rhs)
    one = np.zeros((n, 72), dtype=np.uint64); one[:, :6] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)
    assert not np.array_equal(lhs, one)
    r = np.repeat(np.array([m.limbs64(m.R_ORDER, 4)], dtype=np.uint64), n, 0)     # GT has order r
    eq(ctx.fq12_pow(lhs, r), one)
    # the kernel squares with the cyclotomic routine when every operand of a warp passes its Frobenius test: GT operands
    # alone (that path) and GT operands mixed with arbitrary field elements inside one warp (the generic path) vs the oracle
    gt = ctx.pairing(pa, qa)
    eq(ctx.fq12_pow(gt, k), _oracle_fq12_pow(gt, k))
    mixed = gt.copy(); mixed[::5] = f[::5]
    eq(ctx.fq12_pow(mixed, k), _oracle_fq12_pow(mixed, k))


@pytest.mark.parametrize("op", ["add", "sub", "mul", "sqr", "neg", "dbl", "inv", "from_repr", "into_repr"])
def test_fr_ops(ctx, op):
    """The scalar field Fr (fr.rs:324-572) against the big-int model, edge values included, plus the reference's KATs."""
    L = lambda v: m.limbs64(v, 4)
    edge = [0, 1, 2, m.R_ORDER - 1, m.R_ORDER - 2, m.FR_MONT_R, (m.R_ORDER - 1) // 2, (m.R_ORDER + 1) // 2, 1 << 254, (1 << 255) - 19 - (1 << 254)]
    edge = [v % m.R_ORDER for v in edge]
    rnd = dg.rand_scalars(600, 70, edge_cases=False)
    a = np.concatenate([np.array([L(v) for v in edge], dtype=np.uint64), rnd])
    b = np.concatenate([np.array([L(v) for v in edge[::-1]], dtype=np.uint64), dg.rand_scalars(600, 71, edge_cases=False)])
    if op == "from_repr":
        extra = np.array([L(m.R_ORDER), L(m.R_ORDER + 1), L((1 << 256) - 1)], dtype=np.uint64)
        a = np.concatenate([a, extra]); b = np.concatenate([b, extra])
    binary = op in ("add", "sub", "mul")
    got, gok = ctx.fr_op(op, a, b if binary else None)
    for i in range(len(a)):
        want, wok = m.fr_op_mont(op, m.from_limbs64(a[i]), m.from_limbs64(b[i]) if binary else None)
        assert m.from_limbs64(got[i]) == want and bool(gok[i]) == wok, (op, i)
    if op == "mul":     # fr.rs:1240-1260
        x = np.array([[0x6b7e9b8faeefc81a, 0xe30a8463f348ba42, 0xeff3cb67a8279c9c, 0x3d303651bd7c774d]], dtype=np.uint64)
        y = np.array([[0x13ae28e3bc35ebeb, 0xa10f4488075cae2c, 0x8160e95a853c3b5d, 0x5ae3f03b561a841d]], dtype=np.uint64)
        assert ctx.fr_op("mul", x, y)[0].tolist() == [[0x23717213ce710f71, 0xdbee1fe53a16e1af, 0xf565d3e1c2a48000, 0x4426507ee75df9d7]]


def test_shared_q_miller_and_pairing(ctx):
    """e(P_i, Q) against ONE prepared Q (coefficients staged in shared memory by a TMA bulk copy): the same values as the
    per-pair kernels and as the oracle's literal miller_loop over the replicated G2Prepared; infinity on either side -> one."""
    n = 333
    p = dg.g1_affine_points(n, 95, infinity_at=(6,))
    q = dg.g2_affine_points(2, 96, infinity_at=(1,))
    qp = o.g2_prepare(q, TH)
    for j in (0, 1):                                    # j == 1: Q at infinity
        rep_q = np.repeat(q[j:j + 1], n, 0)
        want_m = o.miller_loop_prepared(p, np.repeat(qp[j:j + 1], n, 0), TH)
        eq(ctx.miller_loop_shared_q(p, qp[j:j + 1]), want_m)
        eq(ctx.pairing_shared_q(p, qp[j:j + 1]), o.pairing(p, rep_q, TH))
    eq(ctx.miller_loop_shared_q(p, ctx.g2_prepare(q[:1])), ctx.miller_loop(p, np.repeat(q[:1], n, 0)))
    assert ctx.pairing_shared_q(p[:0], qp[:1]).shape == (0, 72)


def test_sqrt_fq_and_fq2(ctx):
    """SqrtField::sqrt for Fq (fq.rs:1147-1170) and Fq2 (fq2.rs:167-221): the specific root of the reference's algorithm,
    None for non-residues; includes the reference's two Fq2 known answers (fq2.rs:795-864)."""
    n = 200
    a = dg.rand_fq(n, 81)
    a[0] = 0
    sq = o.fq_op("sqr", a)[0]
    both = np.concatenate([a, sq])
    got, ok = ctx.field_op(1, "sqrt", both)
    for i in range(len(both)):
        v = m.from_mont(m.from_limbs64(both[i]))
        want = m.fq_sqrt(v)
        assert bool(ok[i]) == (want is not None)
        if want is not None:
            assert m.from_mont(m.from_limbs64(got[i])) == want
    assert ok[n:].all() and 0 < ok[:n].sum() < n
    a2 = dg.rand_field(n, 2, 82)
    a2[0] = 0
    sq2 = o.fq2_op("sqr", a2)[0]
    L = m.from_limbs64
    kat = [(L([0x476b4c309720e227, 0x34c2d04faffdab6, 0xa57e6fc1bab51fd9, 0xdb4a116b5bf74aa1, 0x1e58b2159dfe10e2, 0x7ca7da1f13606ac]),
            L([0xfa8de88b7516d2c3, 0x371a75ed14f41629, 0x4cec2dca577a3eb6, 0x212611bca4e99121, 0x8ee5394d77afb3d, 0xec92336650e49d5])),
           (L([0xb9f78429d1517a6b, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624, 0x64774b84f38512bf, 0x4b1ba7b6434bacd7, 0x1a0111ea397fe69a]), 0)]
    krows = np.array([m.limbs64(m.to_mont(c0)) + m.limbs64(m.to_mont(c1)) for c0, c1 in kat], dtype=np.uint64)
    both2 = np.concatenate([a2, sq2, krows])
    got2, ok2 = ctx.field_op(2, "sqrt", both2)
    for i in range(len(both2)):
        v = (m.from_mont(m.from_limbs64(both2[i, :6])), m.from_mont(m.from_limbs64(both2[i, 6:])))
        want = m.fq2_sqrt(v)
        assert bool(ok2[i]) == (want is not None), i
        if want is not None:
            assert (m.from_mont(m.from_limbs64(got2[i, :6])), m.from_mont(m.from_limbs64(got2[i, 6:]))) == want, i
    assert ok2[n:].all()

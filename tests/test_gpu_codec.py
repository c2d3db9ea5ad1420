"""GPU point decoding / encoding against the reference's own vector files (k*G for
k = 0..999, compressed and uncompressed, bls12_381/tests/*.dat) and against the big-int model on the
reference's invalid-vector cases (bls12_381/tests/mod.rs:99-611)."""
import os

import numpy as np
import pytest

import bls_model as m
import oracle_lib as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TH = o.default_threads()


def _aff_rows(pts, g2):
    """model affine points -> ABI rows (Montgomery limbs + infinity word)"""
    rows = np.zeros((len(pts), 25 if g2 else 13), dtype=np.uint64)
    for i, (x, y, inf) in enumerate(pts):
        vals = ([x[0], x[1], y[0], y[1]] if g2 else [x, y])
        for j, v in enumerate(vals):
            rows[i, 6 * j:6 * j + 6] = np.array(m.limbs64(m.to_mont(v)), dtype=np.uint64)
        rows[i, -1] = 1 if inf else 0
    return rows


@pytest.mark.parametrize("g2", [False, True])
def test_generator_multiples_vectors(ctx, g2):
    """decode(entry k) == k*G for all 1000 entries (checked decoding: curve + subgroup), re-encoding reproduces the
    files byte for byte, and compressed and uncompressed decodings agree."""
    name = "g2" if g2 else "g1"
    bu = open(os.path.join(GOLD, name + "_uncompressed_multiples.bin"), "rb").read()
    bc = open(os.path.join(GOLD, name + "_compressed_multiples.bin"), "rb").read()
    au, su = ctx.decode(g2, bu, compressed=False, checked=True)
    ac, sc = ctx.decode(g2, bc, compressed=True, checked=True)
    assert not su.any() and not sc.any()
    assert np.array_equal(au, ac) and au.shape[0] == 1000
    # k*G from the oracle: running projective sum, then into_affine
    g1a, g2a = o.generators()
    one = (o.g2_from_affine if g2 else o.g1_from_affine)(g2a if g2 else g1a)
    w = 36 if g2 else 18
    acc = np.zeros((1, w), dtype=np.uint64)
    acc[0, w // 3:w // 3 + 6] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)
    pts = np.zeros((1000, w), dtype=np.uint64)
    for k in range(1000):
        pts[k] = acc[0]
        acc = (o.g2_op if g2 else o.g1_op)("add", acc, one)
    assert np.array_equal(au, (o.g2_into_affine if g2 else o.g1_into_affine)(pts))
    assert ctx.encode(g2, au, compressed=False) == bu
    assert ctx.encode(g2, au, compressed=True) == bc


@pytest.mark.parametrize("g2", [False, True])
def test_invalid_encodings_match_model(ctx, g2):
    """every malformed encoding of bls12_381/tests/mod.rs:99-611 (+ a few more) gets the model's status, i.e. the
    reference's GroupDecodingError, checked and unchecked."""
    gen = m.G2_GEN_AFFINE if g2 else m.G1_GEN_AFFINE
    qb = m.Q.to_bytes(48, "big")
    for compressed in (False, True):
        size = (96 if g2 else 48) * (1 if compressed else 2)
        z = bytearray(m.encode_point(m._affine_zero(g2), g2, compressed))
        g = bytearray(m.encode_point(gen, g2, compressed))
        cases = [bytes(z), bytes(g)]
        for src in (z, g):
            for bit in (0x80, 0x40, 0x20):
                c = bytearray(src); c[0] ^= bit; cases.append(bytes(c))
        for i in range(size):
            c = bytearray(z); c[i] |= 1; cases.append(bytes(c))
        for slot in range(size // 48):
            c = bytearray(g); c[48 * slot:48 * slot + 48] = qb
            if compressed:
                c[0] |= 0x80
            cases.append(bytes(c))
            c2 = bytearray(c); c2[48 * slot + 47] ^= 1; cases.append(bytes(c2))      # q - 1 (or q + 1 with flags): valid integer
        # small x values: on the curve or not, in the subgroup or not (tests/mod.rs:190-219, 418-470)
        for v in range(0, 12):
            x = (v, 0) if g2 else v
            p = m.get_point_from_x(x, bool(v & 1), g2)
            if p is not None:
                cases.append(m.encode_point(p, g2, compressed))
            elif compressed:
                c = bytearray((x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big")) if g2 else x.to_bytes(48, "big")); c[0] |= 0x80
                cases.append(bytes(c))
            else:
                y = m.G2_GEN_AFFINE[1] if g2 else m.G1_GEN_AFFINE[1]
                cases.append(m.encode_point((x, y, False), g2, False))
        blob = b"".join(cases)
        for checked in (True, False):
            aff, status = ctx.decode(g2, blob, compressed, checked)
            want = [m.decode_point(c, g2, compressed, checked) for c in cases]
            assert status.tolist() == [w[0] for w in want], (compressed, checked)
            good = [i for i, w in enumerate(want) if w[0] == 0]
            assert np.array_equal(aff[good], _aff_rows([want[i][1] for i in good], g2))
            assert set(status.tolist()) >= ({0, 1, 2, 3, 4} if checked else {0, 1, 2, 3}) - ({3} if (not compressed and not checked) else set())


@pytest.mark.parametrize("g2", [False, True])
def test_encode_decode_round_trip_random(ctx, g2):
    import datagen as dg
    n = 300
    aff = (dg.g2_affine_points if g2 else dg.g1_affine_points)(n, 91, infinity_at=(7,))
    for compressed in (False, True):
        enc = ctx.encode(g2, aff, compressed)
        back, status = ctx.decode(g2, enc, compressed, checked=True)
        assert not status.any() and np.array_equal(back, aff)
    # the greatest-y flag: negating y flips bit 5 of the compressed form and nothing else
    neg = aff.copy()
    negated = (o.g2_op if g2 else o.g1_op)("negate", (o.g2_from_affine if g2 else o.g1_from_affine)(aff))
    neg = (o.g2_into_affine if g2 else o.g1_into_affine)(negated)
    e1 = np.frombuffer(ctx.encode(g2, aff, True), dtype=np.uint8).reshape(n, -1)
    e2 = np.frombuffer(ctx.encode(g2, neg, True), dtype=np.uint8).reshape(n, -1)
    live = aff[:, -1] == 0
    assert np.array_equal(e1[live][:, 1:], e2[live][:, 1:]) and np.all((e1[live][:, 0] ^ e2[live][:, 0]) == 0x20)


@pytest.mark.parametrize("g2", [False, True])
def test_point_from_x_and_scale_by_cofactor(ctx, g2):
    """G::rand without the generator (ec.rs:199-214): get_point_from_x for random and small x (some have no root), then
    scale_by_cofactor -- the Jacobian triple of the reference's mul_bits -- and the result lies in the r-order subgroup."""
    import datagen as dg
    n = 24
    xs = [(v, 0) if g2 else v for v in range(8)]
    rnd = dg.rand_fq(2 * n, 123)
    for i in range(n):
        a, b = m.from_limbs64(rnd[2 * i]) % m.Q, m.from_limbs64(rnd[2 * i + 1]) % m.Q
        xs.append((a, b) if g2 else a)
    greatest = [i & 1 for i in range(len(xs))]
    rows = np.zeros((len(xs), 12 if g2 else 6), dtype=np.uint64)
    for i, x in enumerate(xs):
        vals = [x[0], x[1]] if g2 else [x]
        for j, v in enumerate(vals):
            rows[i, 6 * j:6 * j + 6] = np.array(m.limbs64(m.to_mont(v)), dtype=np.uint64)
    aff, ok = ctx.point_from_x(g2, rows, greatest)
    want = [m.get_point_from_x(x, bool(g), g2) for x, g in zip(xs, greatest)]
    assert ok.tolist() == [int(w is not None) for w in want] and 0 < ok.sum() < len(xs)
    good = [i for i, w in enumerate(want) if w is not None]
    assert np.array_equal(aff[good], _aff_rows([want[i] for i in good], g2))
    scaled = ctx.scale_by_cofactor(g2, aff[good])
    for row, i in zip(scaled, good[:6]):                       # the model is pure Python: a few elements
        assert row.tobytes() == m.to_bytes(m.scale_by_cofactor(want[i], g2))
    # every scaled point is in the subgroup: checked decoding of its encoding succeeds
    saff = (ctx.g2_into_affine if g2 else ctx.g1_into_affine)(scaled)
    back, status = ctx.decode(g2, ctx.encode(g2, saff, True), True, checked=True)
    assert not status.any() and np.array_equal(back, saff)
    # ... while the unscaled points are (almost surely) not
    _, st0 = ctx.decode(g2, ctx.encode(g2, aff[good], True), True, checked=True)
    assert (st0 == m.DEC_NOT_IN_SUBGROUP).sum() >= len(good) - 1


@pytest.mark.parametrize("g2", [False, True])
def test_affine_mul(ctx, g2):
    """CurveAffine::mul (ec.rs:174-177): the Jacobian triple of mul_bits with mixed additions (big-int model on a few
    elements), the same point as the projective wNAF / mul_assign after normalisation on all, infinity and scalar 0/1."""
    import datagen as dg
    n = 60
    aff = (dg.g2_affine_points if g2 else dg.g1_affine_points)(n, 97, infinity_at=(4,))
    k = dg.rand_scalars(n, 98)
    got = ctx.affine_mul(g2, aff, k)
    F = m._F2 if g2 else m._F1
    for i in (0, 1, 2, 5, 20):                                 # scalars 0, 1, r-1, 2^129-ish, random
        if g2:
            x = (m.from_mont(m.from_limbs64(aff[i, 0:6])), m.from_mont(m.from_limbs64(aff[i, 6:12])))
            y = (m.from_mont(m.from_limbs64(aff[i, 12:18])), m.from_mont(m.from_limbs64(aff[i, 18:24])))
        else:
            x, y = m.from_mont(m.from_limbs64(aff[i, 0:6])), m.from_mont(m.from_limbs64(aff[i, 6:12]))
        want = m.affine_mul((x, y, bool(aff[i, -1])), m.from_limbs64(k[i]), g2)
        assert got[i].tobytes() == m.to_bytes(want), i
    proj = (o.g2_from_affine if g2 else o.g1_from_affine)(aff)
    ref = (o.g2_op if g2 else o.g1_op)("mul", proj, k=k, threads=TH)
    to_aff = o.g2_into_affine if g2 else o.g1_into_affine
    assert np.array_equal(to_aff(got), to_aff(ref))


@pytest.mark.parametrize("g2", [False, True])
def test_subgroup_test_on_cofactor_points(ctx, g2):
    """The kernels decide subgroup membership with an endomorphism (codec.cuh: (beta x, y) = [-u^2]P on G1, psi(Q) = [u]Q on
    G2) where the reference multiplies by r (ec.rs:142-144).  Same answer required on the points built to tell the two apart:
    points of small prime order in the cofactor group, the cofactor part [r]P of random curve points, random curve points,
    and their cofactor-cleared images (in the subgroup) -- statuses against the big-integer model's multiplication by r."""
    import random
    rng = random.Random(77 + g2)
    F = m._F2 if g2 else m._F1
    u = -m.BLS_X
    h = (u**8 - 4 * u**7 + 5 * u**6 - 4 * u**4 + 6 * u**3 - 4 * u**2 - 4 * u + 13) // 9 if g2 else (u - 1) ** 2 // 3
    small = [13, 23, 2713, 11953, 262069] if g2 else [3, 11, 10177, 859267]
    assert all(h % l == 0 for l in small)

    def curve_point():
        while True:
            x = (rng.randrange(m.Q), rng.randrange(m.Q)) if g2 else rng.randrange(m.Q)
            p = m.get_point_from_x(x, bool(rng.getrandbits(1)), g2)
            if p is not None:
                return p

    def aff(j):
        return m.pt_to_affine(F, j)

    pts = []
    for l in small:                                   # order-l points (when the l-part of the random point is non-trivial)
        for _ in range(2):
            j = m.pt_mul(F, m.pt_from_affine(F, curve_point()), h * m.R_ORDER // l)
            if not m.pt_is_zero(F, j):
                pts.append(aff(j))
    for _ in range(3 if g2 else 6):
        p = curve_point()
        pts.append(p)                                                      # full order
        pts.append(aff(m.pt_mul(F, m.pt_from_affine(F, p), m.R_ORDER)))    # its cofactor part
        pts.append(aff(m.pt_mul(F, m.pt_from_affine(F, p), h)))            # its r-part: in the subgroup
    enc = [m.encode_point(p, g2, False) for p in pts]
    want = [m.decode_point(e, g2, False, True)[0] for e in enc]
    assert want.count(0) >= (3 if g2 else 6) and want.count(m.DEC_NOT_IN_SUBGROUP) >= len(pts) // 2
    _, status = ctx.decode(g2, b"".join(enc), False, True)
    assert status.tolist() == want

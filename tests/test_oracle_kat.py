"""Pins the oracle (oracle/bls_model.py and oracle/bls_oracle.c) on the reference's OWN known-answer
vectors (tests/golden/reference_kats.json, extracted by tests/golden/make_golden.py from the #[test]
functions of /root/reference/src/bls12_381/{fq,fq2,ec}.rs and tests/mod.rs) and on the k*G vector files.
CPU only."""
import json
import os

import numpy as np
import pytest

import bls_model as m
import oracle_lib as o

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))


def R(name):
    return [int(x, 16) for x in KATS[name]["reprs"]]


def limbs(x, n=6):
    return np.array([m.limbs64(x, n)], dtype=np.uint64)


def fq_from_repr(x):
    """Fq::from_repr(FqRepr(x)) through the C oracle -> Montgomery limbs (1,6)."""
    out, ok = o.fq_op("from_repr", limbs(x))
    assert ok[0] == 1
    return out


def test_montgomery_constants():
    """fq.rs:6-43, 501-508, 69-76: MODULUS, R, R2, INV, NEGATIVE_ONE, B_COEFF as quoted by the reference."""
    assert R("const_MODULUS") == [m.Q]
    assert R("const_R") == [m.MONT_R]
    assert R("const_R2") == [m.MONT_R2]
    assert R("const_INV") == [m.INV64]
    assert R("const_NEGATIVE_ONE") == [m.to_mont(m.Q - 1)]
    assert R("const_B_COEFF") == [m.to_mont(4)]
    # the same through the C oracle: from_repr(4) == B_COEFF (fq.rs:1174-1176), -one (fq.rs:1890-1895)
    assert fq_from_repr(4).tobytes() == limbs(R("const_B_COEFF")[0]).tobytes()
    one = limbs(m.MONT_R)
    assert o.fq_op("neg", one)[0].tobytes() == limbs(R("const_NEGATIVE_ONE")[0]).tobytes()


def test_generators():
    """fq.rs:85-136"""
    g1, g2 = o.generators()
    want1 = [R("const_G1_GENERATOR_X")[0], R("const_G1_GENERATOR_Y")[0]]
    assert [m.from_limbs64(g1[0, :6]), m.from_limbs64(g1[0, 6:12])] == want1
    want2 = [R("const_G2_GENERATOR_" + s)[0] for s in ("X_C0", "X_C1", "Y_C0", "Y_C1")]
    assert [m.from_limbs64(g2[0, 6 * i:6 * i + 6]) for i in range(4)] == want2
    assert want1 == [m.to_mont(m.G1_X), m.to_mont(m.G1_Y)]
    assert want2 == [m.to_mont(v) for v in (m.G2_X[0], m.G2_X[1], m.G2_Y[0], m.G2_Y[1])]


def test_frobenius_tables_match_reference_entries():
    """fq.rs:1179-1887 re-derives every table entry with pow(); here the derived tables are compared
    with the c0/c1 limbs the reference source quotes (fq.rs:139-498)."""
    assert R("const_FROBENIUS_COEFF_FQ2_C1") == [m.to_mont(v) for v in m.FROB_FQ2_C1]
    for name, table in (("FQ6_C1", m.FROB_FQ6_C1), ("FQ6_C2", m.FROB_FQ6_C2), ("FQ12_C1", m.FROB_FQ12_C1)):
        want = R("const_FROBENIUS_COEFF_" + name)
        assert len(want) == 2 * len(table)
        assert want == [m.to_mont(c) for e in table for c in e]
    # algebraic definition check (what the reference test does): coefficient = nonresidue^((q^i-1)/k)
    for i in range(1, 6):
        assert m.fq2_pow(m.FROB_FQ6_C1[i], 3) == m.fq2_pow(m.NONRES, m.Q ** i - 1)
        assert m.FROB_FQ6_C2[i] == m.fq2_mul(m.FROB_FQ6_C1[i], m.FROB_FQ6_C1[i])
    for i in range(1, 12):
        assert m.fq2_mul(m.FROB_FQ12_C1[i], m.FROB_FQ12_C1[i]) == m.FROB_FQ6_C1[i % 6]


def test_fq_mul_kat():
    """fq.rs:2558-2584 (Montgomery-form operands and result)"""
    a, b, c = R("test_fq_mul_assign")
    assert o.fq_op("mul", limbs(a), limbs(b))[0].tobytes() == limbs(c).tobytes()
    assert m.to_mont(m.fq_mul(m.from_mont(a), m.from_mont(b))) == c


def test_fq_square_kat():
    """fq.rs:2630-2651"""
    a, c = R("test_fq_squaring")
    assert o.fq_op("sqr", limbs(a))[0].tobytes() == fq_from_repr(c).tobytes()


def test_fq_from_into_repr_kat():
    """fq.rs:2778-2822: q is not in the field; from_repr(a) * from_repr(b) == from_repr(c)"""
    reprs = R("test_fq_from_into_repr")
    a, b, c = reprs[-3:]
    assert o.fq_op("from_repr", limbs(m.Q))[1][0] == 0
    prod = o.fq_op("mul", fq_from_repr(a), fq_from_repr(b))[0]
    assert o.fq_op("into_repr", prod)[0].tobytes() == limbs(c).tobytes()
    assert (a * b) % m.Q == c


def test_fq_add_sub_kats():
    """fq.rs:2325-2425, 2452-2537: the quoted operands/results obey the group law in the model and in C,
    including the wrap-around cases q-1 + 1 = 0 and 0 - 1 = q-1."""
    one, qm1 = limbs(1), limbs(m.Q - 1)
    assert o.fq_op("add", qm1, one)[0].tobytes() == limbs(0).tobytes()
    assert o.fq_op("sub", limbs(0), one)[0].tobytes() == qm1.tobytes()
    vals = [v for v in R("test_fq_add_assign") + R("test_fq_sub_assign") if v < m.Q]
    a = np.array([m.limbs64(v) for v in vals], dtype=np.uint64)
    b = np.roll(a, 1, axis=0)
    s = o.fq_op("add", a, b)[0]
    d = o.fq_op("sub", a, b)[0]
    for i, (x, y) in enumerate(zip(vals, vals[-1:] + vals[:-1])):
        assert m.from_limbs64(s[i]) == (x + y) % m.Q
        assert m.from_limbs64(d[i]) == (x - y) % m.Q
    # first add KAT: tmp + 1 (fq.rs:2352-2364)
    r = R("test_fq_add_assign")
    assert o.fq_op("add", limbs(r[0]), one)[0].tobytes() == limbs(r[2]).tobytes()


def _fq2(c0, c1):
    return np.concatenate([fq_from_repr(c0), fq_from_repr(c1)], axis=1)


def test_fq2_kats():
    """fq2.rs:273-457: squaring, multiplication and inverse known answers"""
    r = R("test_fq2_mul")
    assert o.fq2_op("mul", _fq2(r[0], r[1]), _fq2(r[2], r[3]))[0].tobytes() == _fq2(r[4], r[5]).tobytes()
    r = R("test_fq2_squaring")
    assert o.fq2_op("sqr", _fq2(r[0], r[1]))[0].tobytes() == _fq2(r[2], r[3]).tobytes()
    r = R("test_fq2_inverse")
    out, ok = o.fq2_op("inv", _fq2(r[0], r[1]))
    assert ok[0] == 1 and out.tobytes() == _fq2(r[2], r[3]).tobytes()
    assert o.fq2_op("inv", np.zeros((1, 12), dtype=np.uint64))[1][0] == 0
    r = R("test_fq2_addition")
    assert o.fq2_op("add", _fq2(r[0], r[1]), _fq2(r[2], r[3]))[0].tobytes() == _fq2(r[4], r[5]).tobytes()
    r = R("test_fq2_subtraction")
    assert o.fq2_op("sub", _fq2(r[0], r[1]), _fq2(r[2], r[3]))[0].tobytes() == _fq2(r[4], r[5]).tobytes()
    r = R("test_fq2_negation")
    assert o.fq2_op("neg", _fq2(r[0], r[1]))[0].tobytes() == _fq2(r[2], r[3]).tobytes()
    r = R("test_fq2_doubling")
    assert o.fq2_op("dbl", _fq2(r[0], r[1]))[0].tobytes() == _fq2(r[2], r[3]).tobytes()
    r = R("test_fq2_frobenius_map")   # fq2.rs:682-792: frobenius^0 = id, ^1 = conj, ^1 again = id, ^2 = id
    a = _fq2(r[0], r[1])
    f1 = o.fq2_op("frob1", a)[0]
    assert f1.tobytes() == _fq2(r[4], r[5]).tobytes()
    assert o.fq2_op("frob1", f1)[0].tobytes() == a.tobytes()


def _g1a(x, y):
    return np.concatenate([fq_from_repr(x), fq_from_repr(y), np.zeros((1, 1), dtype=np.uint64)], axis=1)


def _g2a(r):
    return np.concatenate([_fq2(r[0], r[1]), _fq2(r[2], r[3]), np.zeros((1, 1), dtype=np.uint64)], axis=1)


def test_g1_add_double_kats():
    """ec.rs:1060-1175: G1 addition / doubling results in affine form"""
    r = R("test_g1_addition_correctness")
    p, q, want = _g1a(r[0], r[1]), _g1a(r[2], r[3]), _g1a(r[4], r[5])
    s = o.g1_op("add", o.g1_from_affine(p), o.g1_from_affine(q))
    assert o.g1_into_affine(s).tobytes() == want.tobytes()
    r = R("test_g1_doubling_correctness")
    d = o.g1_op("double", o.g1_from_affine(_g1a(r[0], r[1])))
    assert o.g1_into_affine(d).tobytes() == _g1a(r[2], r[3]).tobytes()


def test_g1_same_y_kat():
    """ec.rs:1178-1262: different x, same y -- add_assign and add_assign_mixed agree with the quoted c"""
    r = R("test_g1_same_y")
    a, b, c = _g1a(r[0], r[1]), _g1a(r[2], r[3]), _g1a(r[4], r[5])
    t1 = o.g1_op("add", o.g1_from_affine(a), o.g1_from_affine(b))
    assert o.g1_into_affine(t1).tobytes() == c.tobytes()
    t2 = o.g1_op("add_mixed", o.g1_from_affine(a), b)
    assert o.g1_into_affine(t2).tobytes() == c.tobytes()


def test_g2_add_double_kats():
    """ec.rs:1802-2017"""
    r = R("test_g2_addition_correctness")
    p, q, want = _g2a(r[0:4]), _g2a(r[4:8]), _g2a(r[8:12])
    s = o.g2_op("add", o.g2_from_affine(p), o.g2_from_affine(q))
    assert o.g2_into_affine(s).tobytes() == want.tobytes()
    r = R("test_g2_doubling_correctness")
    d = o.g2_op("double", o.g2_from_affine(_g2a(r[0:4])))
    assert o.g2_into_affine(d).tobytes() == _g2a(r[4:8]).tobytes()


def test_relic_pairing_kat():
    """tests/mod.rs:5-53: e(G1::one(), G2::one()) == the RELIC value, in the model, in C, and in the
    committed fixture the GPU test compares against."""
    want = [int(x) for x in KATS["test_pairing_result_against_relic"]["decimal"]]
    assert len(want) == 12
    assert m.flatten(m.pairing(m.G1_GEN_AFFINE, m.G2_GEN_AFFINE)) == want
    g1, g2 = o.generators()
    e = o.pairing(g1, g2)
    assert [m.from_mont(m.from_limbs64(e[0, 6 * i:6 * i + 6])) for i in range(12)] == want
    assert e.tobytes() == open(os.path.join(GOLD, "relic_pairing_g1g2.bin"), "rb").read()


# ---- k*G serialisation vectors (tests/mod.rs:55-97; format: bls12_381/README.md:59-70, ec.rs:690-822)
def _encode_fq(x):
    return x.to_bytes(48, "big")


def _enc_uncompressed(pt, g2):
    x, y, inf = pt
    size = 192 if g2 else 96
    if inf:
        b = bytearray(size); b[0] |= 1 << 6
        return bytes(b)
    if g2:
        return _encode_fq(x[1]) + _encode_fq(x[0]) + _encode_fq(y[1]) + _encode_fq(y[0])   # c1 then c0 (ec.rs:1371-1374)
    return _encode_fq(x) + _encode_fq(y)


def _enc_compressed(pt, g2):
    x, y, inf = pt
    size = 96 if g2 else 48
    if inf:
        b = bytearray(size); b[0] |= (1 << 7) | (1 << 6)
        return bytes(b)
    if g2:
        b = bytearray(_encode_fq(x[1]) + _encode_fq(x[0]))
        negy = m.fq2_neg(y)
        greatest = (y[1], y[0]) > (negy[1], negy[0])          # Fq2 ordering: c1 first (fq2.rs:21-31)
    else:
        b = bytearray(_encode_fq(x))
        greatest = y > (m.Q - y) % m.Q
    b[0] |= 1 << 7
    if greatest:
        b[0] |= 1 << 5
    return bytes(b)


@pytest.mark.parametrize("g2", [False, True])
def test_generator_multiples_vectors(g2):
    """e = 0; 1000 times { encode(e.into_affine()); e += one } must reproduce the .dat files byte for byte,
    both through the C oracle (add_assign + into_affine on Montgomery limbs) and the big-int model."""
    F = m._F2 if g2 else m._F1
    gen = m.G2_GEN_AFFINE if g2 else m.G1_GEN_AFFINE
    name = "g2" if g2 else "g1"
    want_u = open(os.path.join(GOLD, name + "_uncompressed_multiples.bin"), "rb").read()
    want_c = open(os.path.join(GOLD, name + "_compressed_multiples.bin"), "rb").read()
    # C oracle: running sum with projective add_assign, then per-point into_affine
    g1a, g2a = o.generators()
    one = (o.g2_from_affine if g2 else o.g1_from_affine)(g2a if g2 else g1a)
    w = 36 if g2 else 18
    acc = np.zeros((1, w), dtype=np.uint64)
    acc[0, w // 3:w // 3 + 6] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)   # zero() = (0, 1, 0)
    pts = np.zeros((1000, w), dtype=np.uint64)
    op = o.g2_op if g2 else o.g1_op
    for k in range(1000):
        pts[k] = acc[0]
        acc = op("add", acc, one)
    aff = (o.g2_into_affine if g2 else o.g1_into_affine)(pts)
    got_u, got_c = bytearray(), bytearray()
    e_model = m.pt_zero(F)
    gen_j = m.pt_from_affine(F, gen)
    for k in range(1000):
        row = aff[k]
        if g2:
            x = (m.from_mont(m.from_limbs64(row[0:6])), m.from_mont(m.from_limbs64(row[6:12])))
            y = (m.from_mont(m.from_limbs64(row[12:18])), m.from_mont(m.from_limbs64(row[18:24])))
            pt = (x, y, bool(row[24]))
        else:
            pt = (m.from_mont(m.from_limbs64(row[0:6])), m.from_mont(m.from_limbs64(row[6:12])), bool(row[12]))
        if k < 40:   # the big-int model agrees with the C oracle point by point (and on the Jacobian triple)
            assert m.pt_to_affine(F, e_model) == pt
            assert m.to_bytes(e_model) == pts[k].tobytes()
            e_model = m.pt_add(F, e_model, gen_j)
        got_u += _enc_uncompressed(pt, g2)
        got_c += _enc_compressed(pt, g2)
    assert bytes(got_u) == want_u
    assert bytes(got_c) == want_c


# ---- point encodings: the model's decode/encode against the reference's own vectors
@pytest.mark.parametrize("g2", [False, True])
def test_model_decodes_and_encodes_generator_multiples(g2):
    """tests/mod.rs:55-97 read the other way: decoding entry k of the .dat files gives k*G (checked decoding
    passes: on the curve and in the subgroup for a sample), and re-encoding gives the bytes back."""
    F = m._F2 if g2 else m._F1
    name = "g2" if g2 else "g1"
    us = 192 if g2 else 96
    cs = us // 2
    want_u = open(os.path.join(GOLD, name + "_uncompressed_multiples.bin"), "rb").read()
    want_c = open(os.path.join(GOLD, name + "_compressed_multiples.bin"), "rb").read()
    e = m.pt_zero(F)
    gen = m.pt_from_affine(F, m.G2_GEN_AFFINE if g2 else m.G1_GEN_AFFINE)
    for k in range(120):
        aff = m.pt_to_affine(F, e)
        bu, bc = want_u[us * k:us * k + us], want_c[cs * k:cs * k + cs]
        checked = k < 6                                   # the subgroup check is a 255-bit scalar multiplication
        assert m.decode_point(bu, g2, False, checked) == (m.DEC_OK, aff)
        assert m.decode_point(bc, g2, True, checked) == (m.DEC_OK, aff)
        assert m.encode_point(aff, g2, False) == bu and m.encode_point(aff, g2, True) == bc
        e = m.pt_add(F, e, gen)


@pytest.mark.parametrize("g2", [False, True])
def test_model_rejects_invalid_encodings(g2):
    """bls12_381/tests/mod.rs:99-611 restated: every malformed encoding gets the reference's error."""
    us = 192 if g2 else 96
    gen = m.G2_GEN_AFFINE if g2 else m.G1_GEN_AFFINE
    for compressed in (False, True):
        size = us // 2 if compressed else us
        z = bytearray(m.encode_point(m._affine_zero(g2), g2, compressed))
        o_ = bytearray(m.encode_point(gen, g2, compressed))
        flip = lambda b, i, v: bytes(b[:i]) + bytes([b[i] ^ v]) + bytes(b[i + 1:])
        assert m.decode_point(flip(z, 0, 0x80), g2, compressed)[0] == m.DEC_UNEXPECTED_COMPRESSION_MODE
        assert m.decode_point(flip(o_, 0, 0x80), g2, compressed)[0] == m.DEC_UNEXPECTED_COMPRESSION_MODE
        assert m.decode_point(flip(z, 0, 0x20), g2, compressed)[0] == m.DEC_UNEXPECTED_INFORMATION
        for i in range(size):
            bad = bytearray(z); bad[i] |= 1
            assert m.decode_point(bytes(bad), g2, compressed)[0] == m.DEC_UNEXPECTED_INFORMATION
        qb = m.Q.to_bytes(48, "big")
        flags = bytes([o_[0] & 0xe0])
        ncoord = size // 48
        for slot in range(ncoord):                          # Fq::char() written over one coordinate
            bad = bytearray(o_); bad[48 * slot:48 * slot + 48] = qb
            if slot == 0:
                bad[0] |= flags[0] & 0x80                   # keep the compression flag as the reference's write_be does not
            st = m.decode_point(bytes(bad), g2, compressed)[0]
            idx = slot if not g2 else [1, 0, 3, 2][slot]    # G2 bytes are c1 then c0
            assert st == m.DEC_COORDINATE + idx, (compressed, slot, st)
        # not on the curve: uncompressed x := 0 with the generator's y (tests/mod.rs:172-188); compressed: the first
        # x = 1, 2, ... whose x^3 + b has no square root (tests/mod.rs:418-440)
        bad = bytearray(o_)
        if compressed:
            x = (1, 0) if g2 else 1
            while m.get_point_from_x(x, False, g2) is not None:
                x = m.fq2_add(x, m.FQ2_ONE) if g2 else x + 1
            bad[:] = (x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big")) if g2 else x.to_bytes(48, "big")
            bad[0] |= 0x80
        else:
            bad[0:48 * (2 if g2 else 1)] = bytes(48 * (2 if g2 else 1))
        assert m.decode_point(bytes(bad), g2, compressed)[0] == m.DEC_NOT_ON_CURVE
    # a point on the curve but outside the r-order subgroup (tests/mod.rs:190-219): first x = 1, 2, ... with a root
    x = (1, 0) if g2 else 1
    while True:
        p = m.get_point_from_x(x, False, g2)
        if p is not None:
            break
        x = m.fq2_add(x, m.FQ2_ONE) if g2 else x + 1
    assert m.is_on_curve(p, g2) and not m.is_in_correct_subgroup_assuming_on_curve(p, g2)
    for compressed in (False, True):
        enc = m.encode_point(p, g2, compressed)
        assert m.decode_point(enc, g2, compressed, checked=False) == (m.DEC_OK, p)
        assert m.decode_point(enc, g2, compressed)[0] == m.DEC_NOT_IN_SUBGROUP


def test_fr_constants_and_kats():
    """fr.rs:20-36 (R, R2, INV) and the Montgomery-form KATs of fr.rs:1240-1260 (mul), 1306-1323 (square), 1450-1486 (repr)"""
    L = m.from_limbs64
    assert m.FR_MONT_R == L([0x1fffffffe, 0x5884b7fa00034802, 0x998c4fefecbc4ff5, 0x1824b159acc5056f])
    assert m.FR_MONT_R2 == L([0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x5d314967254398f, 0x748d9d99f59ff11])
    assert m.FR_INV64 == 0xfffffffeffffffff
    a = L([0x6b7e9b8faeefc81a, 0xe30a8463f348ba42, 0xeff3cb67a8279c9c, 0x3d303651bd7c774d])
    b = L([0x13ae28e3bc35ebeb, 0xa10f4488075cae2c, 0x8160e95a853c3b5d, 0x5ae3f03b561a841d])
    assert m.fr_op_mont("mul", a, b)[0] == L([0x23717213ce710f71, 0xdbee1fe53a16e1af, 0xf565d3e1c2a48000, 0x4426507ee75df9d7])
    s = L([0xffffffffffffffff, 0xffffffffffffffff, 0xffffffffffffffff, 0x73eda753299d7d47])
    want = m.fr_op_mont("from_repr", L([0xc0d698e7bde077b8, 0xb79a310579e76ec2, 0xac1da8d0a9af4e5f, 0x13f629c49bf23e97]))[0]
    assert m.fr_op_mont("sqr", s)[0] == want
    ra = L([0x25ebe3a3ad3c0c6a, 0x6990e39d092e817c, 0x941f900d42f5658e, 0x44f8a103b38a71e0])
    rb = L([0x264e9454885e2475, 0x46f7746bb0308370, 0x4683ef5347411f9, 0x58838d7f208d4492])
    rc = L([0x48a09ab93cfc740d, 0x3a6600fbfc7a671, 0x838567017501d767, 0x7161d6da77745512])
    prod = m.fr_op_mont("mul", m.fr_op_mont("from_repr", ra)[0], m.fr_op_mont("from_repr", rb)[0])[0]
    assert m.fr_op_mont("into_repr", prod)[0] == rc
    assert m.fr_op_mont("from_repr", m.R_ORDER) == (0, False) and m.fr_op_mont("from_repr", m.R_ORDER + 1) == (0, False)
    assert m.fr_op_mont("inv", 0) == (0, False)


def test_fq2_sqrt_kats():
    """fq2.rs:795-864: the two Fq2 square-root known answers (the root the algorithm returns, not just a root)"""
    L = m.from_limbs64
    a = (L([0x476b4c309720e227, 0x34c2d04faffdab6, 0xa57e6fc1bab51fd9, 0xdb4a116b5bf74aa1, 0x1e58b2159dfe10e2, 0x7ca7da1f13606ac]),
         L([0xfa8de88b7516d2c3, 0x371a75ed14f41629, 0x4cec2dca577a3eb6, 0x212611bca4e99121, 0x8ee5394d77afb3d, 0xec92336650e49d5]))
    r = (L([0x40b299b2704258c5, 0x6ef7de92e8c68b63, 0x6d2ddbe552203e82, 0x8d7f1f723d02c1d3, 0x881b3e01b611c070, 0x10f6963bbad2ebc5]),
         L([0xc099534fc209e752, 0x7670594665676447, 0x28a20faed211efe7, 0x6b852aeaf2afcb1b, 0xa4c93b08105d71a9, 0x8d7cfff94216330]))
    assert m.fq2_sqrt(a) == r
    b = (L([0xb9f78429d1517a6b, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624, 0x64774b84f38512bf, 0x4b1ba7b6434bacd7, 0x1a0111ea397fe69a]), 0)
    rb = (0, L([0xb9fefffffd4357a3, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624, 0x64774b84f38512bf, 0x4b1ba7b6434bacd7, 0x1a0111ea397fe69a]))
    assert m.fq2_sqrt(b) == rb
    assert m.fq_sqrt(0) == 0 and m.fq2_sqrt((0, 0)) == (0, 0)
    for v in (2, 3, 4, 5, 7, 12345678901234567890):          # fq.rs:2747-2775: sqrt(a^2) = +-a, sqrt(a)^2 = a
        s = m.fq_sqrt(v * v % m.Q)
        assert s in (v % m.Q, (-v) % m.Q)
        t = m.fq_sqrt(v)
        assert t is None or t * t % m.Q == v

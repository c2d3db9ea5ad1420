"""The micro-programs of the warp-cooperative engine (tools/wide_gen.py -> pairing_b200/csrc/wide_prog_gen.cuh), simulated
on integers with the interpreter's round semantics, against the big-integer model of the reference."""
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools")]
import bls_model as m
import wide_gen as wg


def _rand_fq12(rnd):
    return tuple(tuple((rnd.randrange(m.Q), rnd.randrange(m.Q)) for _ in range(3)) for _ in range(2))


def _inputs12(name, a):
    return {"%s.c%d.c%d" % (name, i, j): a[i][j] for i in range(2) for j in range(3)}


def _as12(outs):
    return (tuple(outs[0:3]), tuple(outs[3:6]))


def test_constants_match_the_reference_tables():
    """the Frobenius coefficients the simulator derives as powers of xi are the reference's tables (fq.rs:1179-1887)"""
    for idx, (name, p) in enumerate(wg.CONSTS):
        want = {"ONE": lambda p: m.FQ2_ONE, "FROB_FQ6_C1": lambda p: m.FROB_FQ6_C1[p],
                "FROB_FQ6_C2": lambda p: m.FROB_FQ6_C2[p], "FROB_FQ12_C1": lambda p: m.FROB_FQ12_C1[p]}[name](p)
        assert wg.const_value(idx) == tuple(want)


def test_fq12_mul_program():
    rnd = random.Random(5)
    for _ in range(3):
        a, b = _rand_fq12(rnd), _rand_fq12(rnd)
        ins = _inputs12("a", a); ins.update(_inputs12("b", b))
        outs, prog = wg.run_program("FQ12_MUL", ins)
        assert _as12(outs) == m.fq12_mul(a, b)
    assert prog.stats["mul_ops"] == 18


def test_final_exponentiation_program():
    """mod.rs:104-160 as one scheduled program: the same GT element as the model's final_exponentiation"""
    rnd = random.Random(7)
    f = _rand_fq12(rnd)
    outs, prog = wg.run_program("FINAL_EXP", _inputs12("f", f))
    assert _as12(outs) == m.final_exponentiation(f)
    assert prog.stats["inv"] == 1


def test_pairing_program_known_answer():
    """the whole pairing program on the generators: e(g1, g2) of the RELIC vector (bls12_381/tests/mod.rs:5-53)"""
    p, q = m.G1_GEN_AFFINE, m.G2_GEN_AFFINE
    outs, _ = wg.run_program("PAIRING", {"px": (p[0], 0), "py": (p[1], 0), "qx": q[0], "qy": q[1]})
    want = m.pairing(p, q)
    assert _as12(outs) == want
    kat = open(os.path.join(ROOT, "tests", "golden", "relic_pairing_g1g2.bin"), "rb").read()
    assert m.to_bytes(_as12(outs)) == kat


def test_miller_program_random_pair():
    rnd = random.Random(11)
    p = m.pt_to_affine(m._F1, m.pt_mul(m._F1, m.pt_from_affine(m._F1, m.G1_GEN_AFFINE), rnd.randrange(1, m.R_ORDER)))
    q = m.pt_to_affine(m._F2, m.pt_mul(m._F2, m.pt_from_affine(m._F2, m.G2_GEN_AFFINE), rnd.randrange(1, m.R_ORDER)))
    outs, _ = wg.run_program("MILLER", {"px": (p[0], 0), "py": (p[1], 0), "qx": q[0], "qy": q[1]})
    assert _as12(outs) == m.miller_loop([(p, m.g2_prepare(q))])


def test_to_affine_program():
    """the two into_affine conversions in front of the pairing on projective inputs (ec.rs:586-619)"""
    rnd = random.Random(13)
    pj = m.pt_mul(m._F1, m.pt_from_affine(m._F1, m.G1_GEN_AFFINE), rnd.randrange(1, m.R_ORDER))
    qj = m.pt_mul(m._F2, m.pt_from_affine(m._F2, m.G2_GEN_AFFINE), rnd.randrange(1, m.R_ORDER))
    pj = m.pt_double(m._F1, pj); qj = m.pt_double(m._F2, qj)              # Z != 1
    ins = {"pX": (pj[0], 0), "pY": (pj[1], 0), "pZ": (pj[2], 0), "qX": qj[0], "qY": qj[1], "qZ": qj[2]}
    outs, prog = wg.run_program("TO_AFFINE", ins)
    pa, qa = m.pt_to_affine(m._F1, pj), m.pt_to_affine(m._F2, qj)
    assert outs == [(pa[0], 0), (pa[1], 0), qa[0], qa[1]]
    assert prog.stats["inv"] == 2


def test_committed_header_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "wide_gen.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr

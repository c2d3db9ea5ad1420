#!/usr/bin/env python3
"""Word-level emulation of the carry-chain schedules in pairing_b200/csrc/fp.cuh (fp_mul, fp_mul2) on the
relaxed operand range [0, 2q].  Checks the arithmetic identity (mod q), the output bound (<= 2q without any
conditional subtraction) and that every carry the PTX drops is zero.  CPU only."""
import random, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bls_model as m

M32 = (1 << 32) - 1
Q = [(m.Q >> (32 * i)) & M32 for i in range(12)]
NINV = m.INV32


def limbs(x): return [(x >> (32 * i)) & M32 for i in range(12)]
def val(l): return sum(v << (32 * i) for i, v in enumerate(l))


def cmad_row(acc, x, off, b, top_holder=None):
    """acc += (x[off], x[off+2], ...) * b as ONE 12-word chain; returns carry-out (added to top if given)."""
    carry = 0
    for k in range(6):
        p = x[off + 2 * k] * b
        lo = acc[2 * k] + (p & M32) + carry
        acc[2 * k] = lo & M32; carry = lo >> 32
        hi = acc[2 * k + 1] + (p >> 32) + carry
        acc[2 * k + 1] = hi & M32; carry = hi >> 32
    return carry


def rshift_row(w, idx, add0, acc, x, off, b):
    """w[idx] += add0 (carry c); acc = (acc >> 64) + x_sel * b + c ; final carry must be 0."""
    s = w[idx] + add0
    w[idx] = s & M32; carry = s >> 32
    old = acc[:]
    for k in range(6):
        p = x[off + 2 * k] * b
        a_lo = old[2 * k + 2] if 2 * k + 2 < 12 else 0
        a_hi = old[2 * k + 3] if 2 * k + 3 < 12 else 0
        lo = a_lo + (p & M32) + carry
        acc[2 * k] = lo & M32; carry = lo >> 32
        hi = a_hi + (p >> 32) + carry
        acc[2 * k + 1] = hi & M32; carry = hi >> 32
    assert carry == 0, "rshift row dropped a carry"


def redc_row(e, o):
    mm = (e[0] * NINV) & M32
    c = cmad_row(o, Q, 1, mm)
    assert c == 0, "q-odd chain dropped a carry"
    c = cmad_row(e, Q, 0, mm)
    o[11] += c
    assert o[11] <= M32, "top word overflow"
    assert e[0] == 0


def merge(e, o):
    """(e >> 32) + o with e word-0 aligned; no conditional subtraction: the result must already be <= 2q."""
    t = (val(e) >> 32) + val(o)
    assert t <= 2 * m.Q, "merge result above 2q: %.3f q" % (t / m.Q)
    return t


def fp_mul(a, b):
    A, B = limbs(a), limbs(b)
    e = [0] * 12; o = [0] * 12
    for k in range(6):
        p = A[2 * k] * B[0]; e[2 * k] = p & M32; e[2 * k + 1] = p >> 32
        p = A[2 * k + 1] * B[0]; o[2 * k] = p & M32; o[2 * k + 1] = p >> 32
    redc_row(e, o)
    for i in range(1, 12, 2):
        rshift_row(o, 0, e[1], e, A, 1, B[i])
        c = cmad_row(o, A, 0, B[i]); e[11] += c; assert e[11] <= M32
        redc_row(o, e)
        if i + 1 < 12:
            rshift_row(e, 0, o[1], o, A, 1, B[i + 1])
            c = cmad_row(e, A, 0, B[i + 1]); o[11] += c; assert o[11] <= M32
            redc_row(e, o)
    return merge(o, e)


def fp_mul2(x, bx, y, by):
    """(x*bx + y*by) / 2^384 mod q, dual-product CIOS (x, y, bx, by <= 2q)."""
    X, BX, Y, BY = limbs(x), limbs(bx), limbs(y), limbs(by)
    e = [0] * 12; o = [0] * 12
    ev, od = e, o            # ev: word-0 aligned accumulator of this row, od: word-1 aligned
    for i in range(12):
        if i == 0:
            for k in range(6):
                p = X[2 * k] * BX[0]; ev[2 * k] = p & M32; ev[2 * k + 1] = p >> 32
                p = X[2 * k + 1] * BX[0]; od[2 * k] = p & M32; od[2 * k + 1] = p >> 32
        else:
            # roles swap every row: previous word-1 accumulator becomes word-0 aligned
            ev, od = od, ev
            rshift_row(ev, 0, od[1], od, X, 1, BX[i])
            c = cmad_row(ev, X, 0, BX[i]); od[11] += c; assert od[11] <= M32
        c = cmad_row(od, Y, 1, BY[i]); assert c == 0, "second product odd chain dropped a carry"
        c = cmad_row(ev, Y, 0, BY[i]); od[11] += c; assert od[11] <= M32
        redc_row(ev, od)
    t = (val(ev) >> 32) + val(od)
    assert t <= 2 * m.Q, "fp_mul2 result above 2q: %.3f q" % (t / m.Q)
    return t


if __name__ == "__main__":
    random.seed(7)
    RI = pow(1 << 384, -1, m.Q)
    Q2 = 2 * m.Q
    edge = [0, 1, m.Q - 1, m.Q, m.Q + 1, Q2 - 1, Q2, (1 << 381) - 1, 1 << 381, m.MONT_R, (m.Q - 1) // 2,
            Q2 - (1 << 32), Q2 - (Q2 % (1 << 352)), int("ffffffff" * 11, 16)]
    assert all(v <= Q2 for v in edge)
    cases = [(a, b) for a in edge for b in edge] + [(random.randrange(Q2 + 1), random.randrange(Q2 + 1)) for _ in range(4000)]
    worst = 0
    for a, b in cases:
        r = fp_mul(a, b)
        assert r % m.Q == a * b * RI % m.Q
        worst = max(worst, r)
    print("fp_mul: %d cases ok, largest result %.3f q" % (len(cases), worst / m.Q))
    cases4 = [(a, b, c, d) for a in edge for b in edge for c in (0, m.Q, Q2 - 1, Q2) for d in (1, Q2)]
    cases4 += [tuple(random.randrange(Q2 + 1) for _ in range(4)) for _ in range(4000)]
    worst = 0
    for a, b, c, d in cases4:
        r = fp_mul2(a, b, c, d)
        assert r % m.Q == (a * b + c * d) * RI % m.Q
        worst = max(worst, r)
    print("fp_mul2: %d cases ok, largest result %.3f q" % (len(cases4), worst / m.Q))

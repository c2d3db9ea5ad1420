#!/usr/bin/env python3
"""Generator, scheduler and simulator of the micro-programs of the warp-cooperative ("wide") tower engine
(pairing_b200/csrc/wide.cuh runs them; pairing_b200/csrc/wide_prog_gen.cuh holds the generated tables).

Why: the throughput kernels give ONE lane pair a whole pairing (9.9 ms of latency for one pairing or for a thousand,
6.5 ms for the single final exponentiation at the tail of a multi-pairing product).  Here ONE WARP works on one
element: its 16 lane pairs execute up to 16 independent Fq2 products per round, operands and results travelling
through a small shared-memory file of Fq2 "slots".  The reference's formulas (bls12_381/fq6.rs, fq12.rs, mod.rs) are
written below once over symbolic Fq2 values; every Fq2 product becomes a node of a data-flow graph whose operands are
small linear combinations (+-1, 2, 4, 8 multiples, optionally times the non-residue xi = 1 + u, or conjugated) of
earlier nodes; the graph is list-scheduled into rounds (critical path first), slots are assigned by liveness, and the
rounds are emitted as a table of 16-bit micro-operations.  The values computed are the reference's field values, so
the canonical outputs are bit-identical.

A round is one of
  MUL  every active lane pair:  dst <- (sum of <= 4 terms) * (sum of <= 4 terms)      (one lazily reduced dual product per lane)
  MUL4 the same product with FOUR lanes per micro-op (rounds of at most 8 products): one plain product per lane
  LIN  every active lane pair:  dst <- sum of <= 8 terms
  INV  the whole warp:          dst <- src^-1 in Fq2                                  (fq2.rs:138-155)
and a term is  [-] 2^k [xi] [conj] slot.

The simulator at the bottom executes the emitted table on integers mod q with the interpreter's read-before-write
round semantics; tests/test_wide_program.py compares its outputs with the big-integer model of the reference.
`python tools/wide_gen.py` rewrites the header, `--check` verifies that the committed header is current."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "pairing_b200", "csrc", "wide_prog_gen.cuh")

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS_X = 0xd201000000010000            # |x|; x is negative (mod.rs:15-16)
W = 16                                # lane pairs of a warp
T_MUL, T_LIN = 4, 8                   # terms per multiplication operand / per linear micro-op
K_MUL, K_LIN, K_INV, K_MUL4 = 0, 1, 2, 3
NONE = 0xffff
COST = {"mul": 1.0, "lin": 0.25, "inv": 240.0, "in": 0.0, "const": 0.0}

# constant table shared with wide.cuh: index -> (name, power)
CONSTS = [("ONE", 0)] + [("FROB_FQ6_C1", p) for p in (1, 2, 3)] + [("FROB_FQ6_C2", p) for p in (1, 2, 3)] + \
         [("FROB_FQ12_C1", p) for p in (1, 2, 3)]


# ------------------------------------------------------------------------------------------------ Fq2 on integers
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_xi(a): return ((a[0] - a[1]) % Q, (a[0] + a[1]) % Q)
def f2_conj(a): return (a[0], (-a[1]) % Q)
def f2_scale(a, k): return ((a[0] * k) % Q, (a[1] * k) % Q)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % Q, Q - 2, Q)
    return ((a[0] * n) % Q, (-a[1] * n) % Q)


def const_value(idx):
    """plain (non-Montgomery) value of constant idx: Frobenius coefficients as powers of xi (fq.rs:1179-1887)"""
    name, p = CONSTS[idx]
    if name == "ONE":
        return (1, 0)
    e = {"FROB_FQ6_C1": (Q ** p - 1) // 3, "FROB_FQ6_C2": (2 * Q ** p - 2) // 3, "FROB_FQ12_C1": (Q ** p - 1) // 6}[name]
    return f2_pow((1, 1), e)


# ------------------------------------------------------------------------------------------------ data-flow graph
class Graph:
    def __init__(self):
        self.kind, self.a, self.b, self.meta = [], [], [], []
        self.lin_cache = {}
        self.const_nodes = {}

    def node(self, kind, a=(), b=(), meta=None):
        self.kind.append(kind); self.a.append(list(a)); self.b.append(list(b)); self.meta.append(meta)
        return len(self.kind) - 1

    def input(self, name):
        return Val(self, {(self.node("in", meta=name), 0, 0): 1})

    def const(self, idx):
        if idx not in self.const_nodes:
            self.const_nodes[idx] = self.node("const", meta=idx)
        return Val(self, {(self.const_nodes[idx], 0, 0): 1})

    def zero(self):
        return Val(self, {})

    # a combination as emitted terms: (node, neg, xi, conj, log2 multiplier)
    @staticmethod
    def terms(v):
        out = []
        for (n, xi, cj), k in sorted(v.t.items()):
            neg, k = k < 0, abs(k)
            assert k < 16, "coefficient too large for the 2^k term encoding: %d" % k
            for bit in range(4):
                if (k >> bit) & 1:
                    out.append((n, neg, xi, cj, bit))
        return out

    def materialize(self, v):
        """a node holding the value of v (no new node when v already is a plain node)"""
        items = tuple(sorted(v.t.items()))
        if len(items) == 1 and items[0] == ((items[0][0][0], 0, 0), 1):
            return v
        if items in self.lin_cache:
            return self.lin_cache[items]
        ts = self.terms(v)
        if len(ts) <= T_LIN:
            r = Val(self, {(self.node("lin", a=ts[:T_LIN // 2], b=ts[T_LIN // 2:]), 0, 0): 1})
        else:
            # split by keys so that each part is a valid combination, then add the parts
            keys = sorted(v.t)
            half = len(keys) // 2
            assert half > 0, "a single coefficient cannot need more than %d terms" % T_LIN
            p0 = self.materialize(Val(self, {k: v.t[k] for k in keys[:half]}))
            p1 = self.materialize(Val(self, {k: v.t[k] for k in keys[half:]}))
            r = self.materialize(p0 + p1)
        self.lin_cache[items] = r
        return r

    def fit(self, v, limit):
        return v if len(self.terms(v)) <= limit else self.materialize(v)

    def mul(self, a, b):
        if not a.t or not b.t:
            return self.zero()
        one = self.const_nodes.get(0)
        for x, y in ((a, b), (b, a)):       # products with a multiple of the constant one
            if one is not None and len(x.t) == 1 and (one, 0, 0) in x.t:
                return y.scale(x.t[(one, 0, 0)])
        a, b = self.fit(a, T_MUL), self.fit(b, T_MUL)
        return Val(self, {(self.node("mul", a=self.terms(a), b=self.terms(b)), 0, 0): 1})

    def inv(self, a):
        a = self.materialize(a)
        (n, _, _), = a.t.keys()
        return Val(self, {(self.node("inv", a=[(n, False, 0, 0, 0)]), 0, 0): 1})


class Val:
    """a symbolic Fq2 value: integer combination of (node, xi, conj) keys"""
    __slots__ = ("g", "t")

    def __init__(self, g, t):
        self.g, self.t = g, {k: c for k, c in t.items() if c}

    def __add__(self, o):
        t = dict(self.t)
        for k, c in o.t.items():
            t[k] = t.get(k, 0) + c
        return Val(self.g, t)

    def __sub__(self, o): return self + (-o)
    def __neg__(self): return Val(self.g, {k: -c for k, c in self.t.items()})
    def scale(self, k): return Val(self.g, {key: c * k for key, c in self.t.items()})
    def dbl(self): return self.scale(2)

    def xi(self):       # times the non-residue 1 + u (fq2.rs:41-45)
        v = self if all(k[1] == 0 and k[2] == 0 for k in self.t) else self.g.materialize(self)
        return Val(self.g, {(n, 1, 0): c for (n, _, _), c in v.t.items()})

    def conj(self):     # Frobenius of Fq2 (fq2.rs:157-159, odd powers)
        v = self if all(k[1] == 0 and k[2] == 0 for k in self.t) else self.g.materialize(self)
        return Val(self.g, {(n, 0, 1): c for (n, _, _), c in v.t.items()})


# ------------------------------------------------------------------------------------------------ the tower, symbolically
# (same formulas as pair_tower.cuh, which cites the reference routine by routine)
def f6_add(a, b): return tuple(x + y for x, y in zip(a, b))
def f6_sub(a, b): return tuple(x - y for x, y in zip(a, b))
def f6_neg(a): return tuple(-x for x in a)
def f6_nr(a): return (a[2].xi(), a[0], a[1])                    # fq6.rs:32-38


def f6_mul(g, a, b):                                            # fq6.rs:199-248
    aa, bb, cc = g.mul(a[0], b[0]), g.mul(a[1], b[1]), g.mul(a[2], b[2])
    t1 = (g.mul(b[1] + b[2], a[1] + a[2]) - bb - cc).xi() + aa
    t3 = g.mul(b[0] + b[2], a[0] + a[2]) - aa + bb - cc
    t2 = g.mul(b[0] + b[1], a[0] + a[1]) - aa - bb + cc.xi()
    return (t1, t2, t3)


def f6_sqr(g, a):                                               # fq6.rs:166-197
    s0 = g.mul(a[0], a[0])
    s1 = g.mul(a[0], a[1]).dbl()
    t = a[0] - a[1] + a[2]
    s2 = g.mul(t, t)
    s3 = g.mul(a[1], a[2]).dbl()
    s4 = g.mul(a[2], a[2])
    return (s3.xi() + s0, s4.xi() + s1, s1 + s2 + s3 - s0 - s4)


def f6_mul_by_1(g, a, c1):                                      # fq6.rs:40-66
    bb = g.mul(a[1], c1)
    return ((g.mul(c1, a[1] + a[2]) - bb).xi(), g.mul(c1, a[0] + a[1]) - bb, bb)


def f6_mul_by_01(g, a, c0, c1):                                 # fq6.rs:68-109
    aa, bb = g.mul(a[0], c0), g.mul(a[1], c1)
    t1 = (g.mul(c1, a[1] + a[2]) - bb).xi() + aa
    t3 = g.mul(c0, a[0] + a[2]) - aa + bb
    t2 = g.mul(c0 + c1, a[0] + a[1]) - aa - bb
    return (t1, t2, t3)


def f6_inv(g, a):                                               # fq6.rs:250-301
    c0 = -g.mul(a[2].xi(), a[1]) + g.mul(a[0], a[0])
    c1 = g.mul(a[2], a[2]).xi() - g.mul(a[0], a[1])
    c2 = g.mul(a[1], a[1]) - g.mul(a[0], a[2])
    t = (g.mul(a[2], c1) + g.mul(a[1], c2)).xi() + g.mul(a[0], c0)
    ti = g.inv(t)
    return (g.mul(ti, c0), g.mul(ti, c1), g.mul(ti, c2))


def f2_frob(a, power): return a.conj() if power & 1 else a


def f6_frob(g, a, power):                                       # fq6.rs:157-164
    c1 = g.const(CONSTS.index(("FROB_FQ6_C1", power)))
    c2 = g.const(CONSTS.index(("FROB_FQ6_C2", power)))
    return (f2_frob(a[0], power), g.mul(f2_frob(a[1], power), c1), g.mul(f2_frob(a[2], power), c2))


def f12_conj(a): return (a[0], f6_neg(a[1]))                   # fq12.rs:30-32


def f12_mul(g, a, b):                                           # fq12.rs:116-130
    aa, bb = f6_mul(g, a[0], b[0]), f6_mul(g, a[1], b[1])
    s = f6_mul(g, f6_add(a[1], a[0]), f6_add(b[0], b[1]))
    return (f6_add(f6_nr(bb), aa), f6_sub(f6_sub(s, aa), bb))


def f12_sqr(g, a):                                              # fq12.rs:99-114
    ab = f6_mul(g, a[0], a[1])
    c0 = f6_mul(g, f6_add(f6_nr(a[1]), a[0]), f6_add(a[0], a[1]))
    return (f6_sub(f6_sub(c0, ab), f6_nr(ab)), f6_add(ab, ab))


def f4_sqr(g, a, b):
    tmp = g.mul(a, b)
    s = g.mul(a + b, b.xi() + a)
    return s - tmp - tmp.xi(), tmp.dbl()


def f12_cyclotomic_sqr(g, f):                                   # Granger-Scott; same value as fq12.rs:99-114 in the cyclotomic subgroup
    (c00, c01, c02), (c10, c11, c12) = f
    t0, t1 = f4_sqr(g, c00, c11)
    t2, t3 = f4_sqr(g, c10, c02)
    t4, t5 = f4_sqr(g, c01, c12)
    x5 = t5.xi()
    z0 = (t0 - c00).dbl() + t0
    z1 = (t1 + c11).dbl() + t1
    z2 = (x5 + c10).dbl() + x5
    z3 = (t4 - c02).dbl() + t4
    z4 = (t2 - c01).dbl() + t2
    z5 = (t3 + c12).dbl() + t3
    return ((z0, z4, z3), (z2, z1, z5))


def f12_mul_by_014(g, f, c0, c1, c4):                           # fq12.rs:34-48
    aa = f6_mul_by_01(g, f[0], c0, c1)
    bb = f6_mul_by_1(g, f[1], c4)
    s = f6_mul_by_01(g, f6_add(f[1], f[0]), c0, c1 + c4)
    return (f6_add(f6_nr(bb), aa), f6_sub(f6_sub(s, aa), bb))


def f12_inv(g, a):                                              # fq12.rs:132-148
    t = f6_inv(g, f6_sub(f6_sqr(g, a[0]), f6_nr(f6_sqr(g, a[1]))))
    return (f6_mul(g, t, a[0]), f6_neg(f6_mul(g, t, a[1])))


def f12_frob(g, a, power):                                      # fq12.rs:90-97
    c0, c1 = f6_frob(g, a[0], power), f6_frob(g, a[1], power)
    k = g.const(CONSTS.index(("FROB_FQ12_C1", power)))
    return (c0, tuple(g.mul(x, k) for x in c1))


def f12_materialize(g, a):
    return tuple(tuple(g.materialize(x) for x in h) for h in a)


def exp_by_x(g, a, x):                                          # mod.rs:116-121 (pow, then conjugate: x is negative)
    a = f12_materialize(g, a)
    res = a
    for n in range(x.bit_length() - 2, -1, -1):
        res = f12_materialize(g, f12_cyclotomic_sqr(g, res))
        if (x >> n) & 1:
            res = f12_materialize(g, f12_mul(g, res, a))
    return f12_conj(res)


def final_exponentiation(g, r):                                 # mod.rs:104-160
    f1 = f12_conj(r)
    f2 = f12_inv(g, r)
    r = f12_mul(g, f1, f2)
    f2 = r
    r = f12_mul(g, f12_frob(g, r, 2), f2)
    y0 = f12_sqr(g, r)
    y1 = exp_by_x(g, y0, BLS_X)
    y2 = exp_by_x(g, y1, BLS_X >> 1)
    y3 = f12_conj(r)
    y1 = f12_mul(g, y1, y3)
    y1 = f12_conj(y1)
    y1 = f12_mul(g, y1, y2)
    y2 = exp_by_x(g, y1, BLS_X)
    y3 = exp_by_x(g, y2, BLS_X)
    y1 = f12_conj(y1)
    y3 = f12_mul(g, y3, y1)
    y1 = f12_conj(y1)
    y1 = f12_frob(g, y1, 3)
    y2 = f12_frob(g, y2, 2)
    y1 = f12_mul(g, y1, y2)
    y2 = exp_by_x(g, y3, BLS_X)
    y2 = f12_mul(g, y2, y0)
    y2 = f12_mul(g, y2, r)
    y1 = f12_mul(g, y1, y2)
    y2 = f12_frob(g, y3, 1)
    return f12_mul(g, y1, y2)


def doubling_step(g, r):                                        # mod.rs:176-245
    rx, ry, rz = r
    tmp0, tmp1 = g.mul(rx, rx), g.mul(ry, ry)
    tmp2 = g.mul(tmp1, tmp1)
    t = tmp1 + rx
    tmp3 = (g.mul(t, t) - tmp0 - tmp2).dbl()
    tmp4 = tmp0.scale(3)
    tmp6 = rx + tmp4
    tmp5 = g.mul(tmp4, tmp4)
    zsq = g.mul(rz, rz)
    nx = tmp5 - tmp3 - tmp3
    t = rz + ry
    nz = g.mul(t, t) - tmp1 - zsq
    ny = g.mul(tmp3 - nx, tmp4) - tmp2.scale(8)
    c1 = -g.mul(tmp4, zsq).dbl()
    c2 = g.mul(tmp6, tmp6) - tmp0 - tmp5 - tmp1.scale(4)
    c0 = g.mul(nz, zsq).dbl()
    return (nx, ny, nz), (c0, c1, c2)


def addition_step(g, r, q):                                     # mod.rs:247-333
    rx, ry, rz = r
    qx, qy = q
    zsq, ysq = g.mul(rz, rz), g.mul(qy, qy)
    t0 = g.mul(zsq, qx)
    t = qy + rz
    t1 = g.mul(g.mul(t, t) - ysq - zsq, zsq)
    t2 = t0 - rx
    t3 = g.mul(t2, t2)
    t4 = t3.scale(4)
    t5 = g.mul(t4, t2)
    t6 = t1 - ry - ry
    t9 = g.mul(t6, qx)
    t7 = g.mul(t4, rx)
    nx = g.mul(t6, t6) - t5 - t7 - t7
    t = rz + t2
    nz = g.mul(t, t) - zsq - t3
    t10 = qy + nz
    t8 = g.mul(t7 - nx, t6)
    ny = t8 - g.mul(ry, t5).dbl()
    t10 = g.mul(t10, t10) - ysq - g.mul(nz, nz)
    return (nx, ny, nz), (nz.dbl(), (-t6).dbl(), t9.dbl() - t10)


def ell(g, f, c, px, py):                                       # mod.rs:57-69; px, py are Fq2 values with a zero u-part
    return f12_mul_by_014(g, f, c[2], g.mul(c[1], px), g.mul(c[0], py))


def miller_loop_single(g, px, py, qx, qy):                      # mod.rs:40-102 for one pair, G2 steps on the fly
    one = g.const(0)
    zero = g.zero()
    f = ((one, zero, zero), (zero, zero, zero))
    r = (qx, qy, one)
    bits = BLS_X >> 1
    for b in range(bits.bit_length() - 2, -1, -1):
        # values that feed many later products are materialised once (this also keeps the integer coefficients small)
        r, c = doubling_step(g, r)
        r = tuple(g.materialize(x) for x in r)
        f = ell(g, f, c, px, py)
        if (bits >> b) & 1:
            r, c = addition_step(g, r, (qx, qy))
            r = tuple(g.materialize(x) for x in r)
            f = ell(g, f, c, px, py)
        f = f12_materialize(g, f12_sqr(g, f))
    r, c = doubling_step(g, r)
    f = ell(g, f, c, px, py)
    return f12_conj(f)


# ------------------------------------------------------------------------------------------------ programs
def f12_inputs(g, name):
    return tuple(tuple(g.input("%s.c%d.c%d" % (name, i, j)) for j in range(3)) for i in range(2))


def flat12(a): return [a[0][0], a[0][1], a[0][2], a[1][0], a[1][1], a[1][2]]


def build_final_exp():
    g = Graph()
    g.const(0)
    f = f12_inputs(g, "f")
    return g, flat12(final_exponentiation(g, f))


def build_pairing():
    g = Graph()
    g.const(0)
    px, py, qx, qy = g.input("px"), g.input("py"), g.input("qx"), g.input("qy")
    f = miller_loop_single(g, px, py, qx, qy)
    return g, flat12(final_exponentiation(g, f12_materialize(g, f)))


def build_to_affine():
    """The two into_affine conversions of Engine::pairing on PROJECTIVE inputs (the crate's bench_pairing_full shape,
    benches/bls12_381/mod.rs:91-107; ec.rs:586-619: x / z^2, y / z^3 -- the z == one shortcut yields the same canonical values).
    Runs in front of the PAIRING program.  G1 coordinates enter (and leave) as Fq2 values with a zero u-part.
    Outputs: px, py, qx, qy."""
    g = Graph()
    names = ["pX", "pY", "pZ", "qX", "qY", "qZ"]
    pX, pY, pZ, qX, qY, qZ = [g.input(nm) for nm in names]
    def to_affine(x, y, z):
        zi = g.inv(z)
        zi2 = g.mul(zi, zi)
        return g.mul(x, zi2), g.mul(y, g.mul(zi2, zi))
    px, py = to_affine(pX, pY, pZ)
    qx, qy = to_affine(qX, qY, qZ)
    return g, [px, py, qx, qy]


def build_miller():
    g = Graph()
    g.const(0)
    px, py, qx, qy = g.input("px"), g.input("py"), g.input("qx"), g.input("qy")
    return g, flat12(miller_loop_single(g, px, py, qx, qy))


def build_fq12_mul():
    g = Graph()
    a, b = f12_inputs(g, "a"), f12_inputs(g, "b")
    return g, flat12(f12_mul(g, a, b))


PROGRAMS = {"FINAL_EXP": build_final_exp, "PAIRING": build_pairing, "TO_AFFINE": build_to_affine, "MILLER": build_miller,
            "FQ12_MUL": build_fq12_mul}


# ------------------------------------------------------------------------------------------------ scheduling
class Program:
    pass


def schedule(g, outputs):
    """list-schedule the nodes the outputs depend on; returns a Program with rounds over slots"""
    outs = []
    for v in outputs:
        v = g.materialize(v) if v.t else None
        outs.append(next(iter(v.t))[0] if v is not None else None)
    n = len(g.kind)
    deps = [sorted({t[0] for t in g.a[i]} | {t[0] for t in g.b[i]}) for i in range(n)]
    need = [False] * n
    stack = [o for o in outs if o is not None]
    while stack:
        i = stack.pop()
        if need[i]:
            continue
        need[i] = True
        stack.extend(deps[i])
    users = [[] for _ in range(n)]
    for i in range(n):
        if need[i]:
            for d in deps[i]:
                users[d].append(i)
    prio = [0.0] * n                       # longest path to an output, own cost included
    for i in range(n - 1, -1, -1):
        if need[i]:
            prio[i] = COST[g.kind[i]] + max([prio[u] for u in users[i]], default=0.0)
    pending = [len(deps[i]) for i in range(n)]
    ready = {"mul": [], "lin": [], "inv": []}
    done_round = [None] * n
    for i in range(n):
        if need[i] and g.kind[i] in ("in", "const"):
            done_round[i] = -1
            for u in users[i]:
                pending[u] -= 1
    for i in range(n):
        if need[i] and g.kind[i] not in ("in", "const") and pending[i] == 0:
            ready[g.kind[i]].append(i)
    rounds = []
    remaining = sum(1 for i in range(n) if need[i] and g.kind[i] not in ("in", "const"))
    while remaining:
        best = {k: max((prio[i] for i in v), default=-1.0) for k, v in ready.items()}
        if best["inv"] >= 0 and best["inv"] >= max(best["mul"], best["lin"]):
            kind = "inv"
        elif best["lin"] >= 0 and (best["lin"] >= best["mul"] or best["mul"] < 0):
            kind = "lin"
        elif best["mul"] >= 0:
            kind = "mul"
        else:
            kind = "lin" if ready["lin"] else "inv"
        cand = sorted(ready[kind], key=lambda i: -prio[i])
        take = cand[:1] if kind == "inv" else cand[:W]
        ready[kind] = [i for i in ready[kind] if i not in set(take)]
        r = len(rounds)
        rounds.append((kind, take))
        remaining -= len(take)
        for i in take:
            done_round[i] = r
        for i in take:
            for u in users[i]:
                pending[u] -= 1
                if pending[u] == 0:
                    ready[g.kind[u]].append(u)
    # slot assignment by liveness: a value lives from the round that writes it to the last round that reads it
    last_use = [-1] * n
    for i in range(n):
        if need[i]:
            for d in deps[i]:
                last_use[d] = max(last_use[d], done_round[i])
    for o in outs:
        if o is not None:
            last_use[o] = len(rounds)
    slot = [None] * n
    free, nslots = [], 0
    inputs = [i for i in range(n) if g.kind[i] == "in"]           # every declared input owns a slot, used or not
    consts = [i for i in range(n) if g.kind[i] == "const" and need[i]]
    for i in inputs + consts:
        slot[i] = nslots; nslots += 1
    expiring = {}
    for i in inputs + consts:
        expiring.setdefault(last_use[i], []).append(i)
    prog = Program()
    prog.rounds = []
    import heapq
    for r, (kind, take) in enumerate(rounds):
        # reads of round r happen before its writes: slots whose last reader is round r can be reused by its writers
        for i in expiring.pop(r, []) + (expiring.pop(-1, []) if r == 0 else []):
            heapq.heappush(free, slot[i])
        ops = []
        for i in take:
            if free:
                slot[i] = heapq.heappop(free)
            else:
                slot[i] = nslots; nslots += 1
            expiring.setdefault(last_use[i], []).append(i)
        for i in take:
            ta = [(slot[t[0]],) + tuple(t[1:]) for t in g.a[i]]
            tb = [(slot[t[0]],) + tuple(t[1:]) for t in g.b[i]]
            ops.append((slot[i], ta, tb))
        prog.rounds.append((kind, ops))
    prog.nslots = nslots
    prog.inputs = [(g.meta[i], slot[i]) for i in inputs]
    prog.consts = [(g.meta[i], slot[i]) for i in consts]
    prog.outputs = [slot[o] if o is not None else None for o in outs]
    prog.stats = {k: sum(1 for kk, _ in rounds if kk == k) for k in ("mul", "lin", "inv")}
    prog.stats["mul_ops"] = sum(len(t) for kk, t in rounds if kk == "mul")
    prog.stats["lin_ops"] = sum(len(t) for kk, t in rounds if kk == "lin")
    return prog


# ------------------------------------------------------------------------------------------------ encoding
def enc_term(t):
    s, neg, xi, cj, k = t
    assert s < 1023 and k < 4
    return s | (int(neg) << 10) | (int(xi) << 11) | (k << 12) | (int(cj) << 14)


def encode(prog):
    """u16 words: per round [kind | nops << 8, max A-terms | max B-terms << 8], then per op [dst, 4 A-terms, 4 B-terms]"""
    code = []
    for kind, ops in prog.rounds:
        k = {"mul": K_MUL, "lin": K_LIN, "inv": K_INV}[kind]
        if kind == "mul" and len(ops) <= W // 2:
            k = K_MUL4          # at most 8 products: FOUR lanes per product, one plain Montgomery product each (wide.cuh)
        nta = max([len(a) for _, a, b in ops], default=0)       # terms to walk per operand side (maxima over the round's micro-ops)
        ntb = max([len(b) for _, a, b in ops], default=0)
        code += [k | (len(ops) << 8), nta | (ntb << 8)]
        for dst, a, b in ops:
            assert len(a) <= 4 and len(b) <= 4
            code += [dst] + [enc_term(t) for t in a] + [NONE] * (4 - len(a)) + [enc_term(t) for t in b] + [NONE] * (4 - len(b))
    return code


# ------------------------------------------------------------------------------------------------ simulator
def simulate(code, nrounds, slots):
    """run the encoded table on a dict slot -> (c0, c1) of integers mod q, with the interpreter's semantics"""
    def comb(words):
        acc = (0, 0)
        for w in words:
            if w == NONE:
                continue
            v = slots[w & 1023]
            if (w >> 14) & 1:
                v = f2_conj(v)
            if (w >> 11) & 1:
                v = f2_xi(v)
            v = f2_scale(v, 1 << ((w >> 12) & 3))
            if (w >> 10) & 1:
                v = f2_scale(v, -1)
            acc = f2_add(acc, v)
        return acc
    pos = 0
    for _ in range(nrounds):
        kind, nops = code[pos] & 0xff, code[pos] >> 8
        pos += 2
        writes = []
        for _ in range(nops):
            op = code[pos:pos + 9]
            pos += 9
            a, b = comb(op[1:5]), comb(op[5:9])
            if kind in (K_MUL, K_MUL4):
                r = f2_mul(a, b)
            elif kind == K_LIN:
                r = f2_add(a, b)
            else:
                r = f2_inv(a) if a != (0, 0) else (0, 0)
            writes.append((op[0], r))
        for d, r in writes:
            slots[d] = r
    assert pos == len(code)
    return slots


def run_program(name, inputs):
    """inputs: dict input-name -> (c0, c1); returns the list of output Fq2 values (plain integers mod q)"""
    g, outs = PROGRAMS[name]()
    prog = schedule(g, outs)
    slots = {}
    for nm, s in prog.inputs:
        slots[s] = inputs[nm]
    for idx, s in prog.consts:
        slots[s] = const_value(idx)
    slots = simulate(encode(prog), len(prog.rounds), slots)
    return [slots[s] if s is not None else (0, 0) for s in prog.outputs], prog


# ------------------------------------------------------------------------------------------------ header
def render(names=("FINAL_EXP", "PAIRING", "TO_AFFINE", "MILLER")):
    out = ["// wide_prog_gen.cuh -- GENERATED by tools/wide_gen.py (do not edit; `python tools/wide_gen.py` rewrites it).",
           "// Micro-programs of the warp-cooperative tower engine (wide.cuh): u16 words, per round",
           "//   [kind | nops << 8, max A-terms | max B-terms << 8] then per micro-op [dst, 4 A-terms, 4 B-terms];",
           "//   term = slot | neg << 10 | xi << 11 | log2(multiplier) << 12 | conj << 14, 0xffff = none.",
           "#pragma once", "#include <stdint.h>", ""]
    for name in names:
        g, outs = PROGRAMS[name]()
        prog = schedule(g, outs)
        code = encode(prog)
        st = prog.stats
        out.append("// %s: %d MUL rounds (%d products), %d LIN rounds (%d sums), %d INV rounds, %d slots"
                   % (name, st["mul"], st["mul_ops"], st["lin"], st["lin_ops"], st["inv"], prog.nslots))
        out.append("#define WIDE_%s_NROUNDS %d" % (name, len(prog.rounds)))
        out.append("#define WIDE_%s_NSLOTS %d" % (name, prog.nslots))
        out.append("// inputs: " + ", ".join("%s -> slot %d" % (nm, s) for nm, s in prog.inputs))
        out.append("static __device__ const uint16_t WIDE_%s_CONST[][2] = { %s };   // {constant index, slot}"
                   % (name, ", ".join("{%d, %d}" % (idx, s) for idx, s in prog.consts) or "{0, 0}"))
        out.append("#define WIDE_%s_NCONST %d" % (name, len(prog.consts)))
        out.append("static __device__ const uint16_t WIDE_%s_OUT[%d] = { %s };"
                   % (name, len(prog.outputs), ", ".join(str(s if s is not None else NONE) for s in prog.outputs)))
        out.append("static __device__ const uint16_t WIDE_%s_CODE[%d] = {" % (name, len(code)))
        for i in range(0, len(code), 24):
            out.append("  " + ",".join("%d" % w for w in code[i:i + 24]) + ",")
        out.append("};")
        out.append("")
    return "\n".join(out)


def main():
    text = render()
    if "--check" in sys.argv:
        assert open(OUT).read() == text, "wide_prog_gen.cuh is stale: run python tools/wide_gen.py"
        print("wide_prog_gen.cuh is current")
        return
    open(OUT, "w").write(text)
    for line in text.splitlines():
        if line.startswith("// ") and "rounds" in line:
            print(line)
    print("wrote %s (%d KB)" % (OUT, len(text) // 1024))


if __name__ == "__main__":
    main()

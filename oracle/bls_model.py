"""Big-integer model of the BLS12-381 pairing / wNAF path of the `pairing` crate (v0.14.2).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker.

This file is the *value-level* oracle: every field element is a Python int in [0, q), every
formula is the reference's formula (same operation order, same special cases), so that not only
canonical values (Fq12, affine points) but also representative-dependent outputs (Jacobian
(X, Y, Z) triples, raw Miller values, G2Prepared coefficients) are the reference's.  It is slow
(pure Python) and is used for small cases, for pinning the C restatement in ``bls_oracle.c`` and
for generating the constants both the C oracle and the CUDA kernels embed.

Pinned by (see tests/test_oracle_kat.py): the RELIC pairing vector
(src/bls12_381/tests/mod.rs:5-53), the Montgomery constants R, R2, INV (src/bls12_381/fq.rs:22-43),
the generator coordinates (fq.rs:85-136), the k*G vectors (src/bls12_381/tests/*.dat).

Reference citations are ``file:line`` into /root/reference/src.
"""

# ----------------------------------------------------------------------------------------------
# Parameters (bls12_381/fq.rs:5-13, fr.rs:6-12, mod.rs:23-25)
# ----------------------------------------------------------------------------------------------
Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xD201000000010000
BLS_X_IS_NEGATIVE = True

MONT_R = (1 << 384) % Q            # fq.rs:22-30
MONT_R2 = (MONT_R * MONT_R) % Q    # fq.rs:33-40
MONT_RINV = pow(MONT_R, -1, Q)
INV64 = (-pow(Q, -1, 1 << 64)) % (1 << 64)   # fq.rs:43
INV32 = INV64 & 0xFFFFFFFF

# Generators (fq.rs:85-136), canonical (non-Montgomery) integers.
G1_X = 3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507
G1_Y = 1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569
G2_X = (352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
        3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758)
G2_Y = (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
        927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582)


def to_mont(x):
    return (x * MONT_R) % Q


def from_mont(x):
    return (x * MONT_RINV) % Q


def limbs64(x, n=6):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs64(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


# ----------------------------------------------------------------------------------------------
# Fq  (fq.rs:796-1016).  Values, so Montgomery form is invisible here.
# ----------------------------------------------------------------------------------------------
def fq_add(a, b): return (a + b) % Q
def fq_sub(a, b): return (a - b) % Q
def fq_neg(a): return (-a) % Q
def fq_dbl(a): return (2 * a) % Q
def fq_mul(a, b): return (a * b) % Q
def fq_sqr(a): return (a * a) % Q


def fq_inv(a):
    """fq.rs:849-902; None for zero."""
    if a == 0:
        return None
    return pow(a, -1, Q)


# ----------------------------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1)   (fq2.rs:39-160)
# ----------------------------------------------------------------------------------------------
FQ2_ZERO = (0, 0)
FQ2_ONE = (1, 0)


def fq2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def fq2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def fq2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
def fq2_dbl(a): return ((2 * a[0]) % Q, (2 * a[1]) % Q)
def fq2_is_zero(a): return a[0] == 0 and a[1] == 0


def fq2_mul(a, b):
    """fq2.rs:123-136"""
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def fq2_sqr(a):
    """fq2.rs:87-101"""
    return (((a[0] + a[1]) * (a[0] - a[1])) % Q, (2 * a[0] * a[1]) % Q)


def fq2_mul_by_nonresidue(a):
    """x (1+u); fq2.rs:41-45"""
    return ((a[0] - a[1]) % Q, (a[0] + a[1]) % Q)


def fq2_mul_fq(a, s): return ((a[0] * s) % Q, (a[1] * s) % Q)


def fq2_inv(a):
    """fq2.rs:138-155"""
    t = fq_inv((a[0] * a[0] + a[1] * a[1]) % Q)
    if t is None:
        return None
    return ((a[0] * t) % Q, (-(a[1] * t)) % Q)


def fq2_pow(a, e):
    r = FQ2_ONE
    for bit in bin(e)[2:]:
        r = fq2_sqr(r)
        if bit == '1':
            r = fq2_mul(r, a)
    return r


# Frobenius coefficient tables (fq.rs:139-498), derived, not copied:
#   FQ2_C1[i]  = (-1)^((q^i-1)/2)            (fq.rs:139-158)
#   FQ6_C1[i]  = (1+u)^((q^i-1)/3)           (fq.rs:160-233)
#   FQ6_C2[i]  = (1+u)^((2q^i-2)/3)          (fq.rs:235-308)
#   FQ12_C1[i] = (1+u)^((q^i-1)/6)           (fq.rs:311-498)
NONRES = (1, 1)
FROB_FQ2_C1 = [pow(Q - 1, (Q ** i - 1) // 2, Q) for i in range(2)]
FROB_FQ6_C1 = [fq2_pow(NONRES, (Q ** i - 1) // 3) if i else FQ2_ONE for i in range(6)]
FROB_FQ6_C2 = [fq2_pow(NONRES, (2 * Q ** i - 2) // 3) if i else FQ2_ONE for i in range(6)]
FROB_FQ12_C1 = [fq2_pow(NONRES, (Q ** i - 1) // 6) if i else FQ2_ONE for i in range(12)]


def fq2_frobenius(a, power):
    """fq2.rs:157-159"""
    return (a[0], (a[1] * FROB_FQ2_C1[power % 2]) % Q)


# ----------------------------------------------------------------------------------------------
# Fq6 = Fq2[v]/(v^3-(1+u))   (fq6.rs:30-302)
# ----------------------------------------------------------------------------------------------
FQ6_ZERO = (FQ2_ZERO, FQ2_ZERO, FQ2_ZERO)
FQ6_ONE = (FQ2_ONE, FQ2_ZERO, FQ2_ZERO)


def fq6_add(a, b): return tuple(fq2_add(x, y) for x, y in zip(a, b))
def fq6_sub(a, b): return tuple(fq2_sub(x, y) for x, y in zip(a, b))
def fq6_neg(a): return tuple(fq2_neg(x) for x in a)
def fq6_is_zero(a): return all(fq2_is_zero(x) for x in a)


def fq6_mul_by_nonresidue(a):
    """fq6.rs:32-38: (c0,c1,c2) -> (c2*(1+u), c0, c1)"""
    return (fq2_mul_by_nonresidue(a[2]), a[0], a[1])


def fq6_mul(a, b):
    """fq6.rs:199-248"""
    aa, bb, cc = fq2_mul(a[0], b[0]), fq2_mul(a[1], b[1]), fq2_mul(a[2], b[2])
    t1 = fq2_mul(fq2_add(b[1], b[2]), fq2_add(a[1], a[2]))
    t1 = fq2_add(fq2_mul_by_nonresidue(fq2_sub(fq2_sub(t1, bb), cc)), aa)
    t3 = fq2_mul(fq2_add(b[0], b[2]), fq2_add(a[0], a[2]))
    t3 = fq2_sub(fq2_add(fq2_sub(t3, aa), bb), cc)
    t2 = fq2_mul(fq2_add(b[0], b[1]), fq2_add(a[0], a[1]))
    t2 = fq2_add(fq2_sub(fq2_sub(t2, aa), bb), fq2_mul_by_nonresidue(cc))
    return (t1, t2, t3)


def fq6_sqr(a):
    """fq6.rs:166-197"""
    s0 = fq2_sqr(a[0])
    ab = fq2_mul(a[0], a[1])
    s1 = fq2_dbl(ab)
    s2 = fq2_sqr(fq2_add(fq2_sub(a[0], a[1]), a[2]))
    bc = fq2_mul(a[1], a[2])
    s3 = fq2_dbl(bc)
    s4 = fq2_sqr(a[2])
    c0 = fq2_add(fq2_mul_by_nonresidue(s3), s0)
    c1 = fq2_add(fq2_mul_by_nonresidue(s4), s1)
    c2 = fq2_sub(fq2_sub(fq2_add(fq2_add(s1, s2), s3), s0), s4)
    return (c0, c1, c2)


def fq6_mul_by_1(a, c1):
    """fq6.rs:40-66"""
    bb = fq2_mul(a[1], c1)
    t1 = fq2_mul_by_nonresidue(fq2_sub(fq2_mul(c1, fq2_add(a[1], a[2])), bb))
    t2 = fq2_sub(fq2_mul(c1, fq2_add(a[0], a[1])), bb)
    return (t1, t2, bb)


def fq6_mul_by_01(a, c0, c1):
    """fq6.rs:68-109"""
    aa = fq2_mul(a[0], c0)
    bb = fq2_mul(a[1], c1)
    t1 = fq2_add(fq2_mul_by_nonresidue(fq2_sub(fq2_mul(c1, fq2_add(a[1], a[2])), bb)), aa)
    t3 = fq2_add(fq2_sub(fq2_mul(c0, fq2_add(a[0], a[2])), aa), bb)
    t2 = fq2_sub(fq2_sub(fq2_mul(fq2_add(c0, c1), fq2_add(a[0], a[1])), aa), bb)
    return (t1, t2, t3)


def fq6_inv(a):
    """fq6.rs:250-301"""
    c0 = fq2_add(fq2_neg(fq2_mul(fq2_mul_by_nonresidue(a[2]), a[1])), fq2_sqr(a[0]))
    c1 = fq2_sub(fq2_mul_by_nonresidue(fq2_sqr(a[2])), fq2_mul(a[0], a[1]))
    c2 = fq2_sub(fq2_sqr(a[1]), fq2_mul(a[0], a[2]))
    tmp1 = fq2_mul_by_nonresidue(fq2_add(fq2_mul(a[2], c1), fq2_mul(a[1], c2)))
    tmp1 = fq2_add(tmp1, fq2_mul(a[0], c0))
    t = fq2_inv(tmp1)
    if t is None:
        return None
    return (fq2_mul(t, c0), fq2_mul(t, c1), fq2_mul(t, c2))


def fq6_frobenius(a, power):
    """fq6.rs:157-164"""
    c0 = fq2_frobenius(a[0], power)
    c1 = fq2_mul(fq2_frobenius(a[1], power), FROB_FQ6_C1[power % 6])
    c2 = fq2_mul(fq2_frobenius(a[2], power), FROB_FQ6_C2[power % 6])
    return (c0, c1, c2)


# ----------------------------------------------------------------------------------------------
# Fq12 = Fq6[w]/(w^2-v)   (fq12.rs:29-149)
# ----------------------------------------------------------------------------------------------
FQ12_ZERO = (FQ6_ZERO, FQ6_ZERO)
FQ12_ONE = (FQ6_ONE, FQ6_ZERO)


def fq12_is_zero(a): return fq6_is_zero(a[0]) and fq6_is_zero(a[1])
def fq12_conjugate(a): return (a[0], fq6_neg(a[1]))


def fq12_mul(a, b):
    """fq12.rs:116-130"""
    aa = fq6_mul(a[0], b[0])
    bb = fq6_mul(a[1], b[1])
    c1 = fq6_sub(fq6_sub(fq6_mul(fq6_add(a[1], a[0]), fq6_add(b[0], b[1])), aa), bb)
    c0 = fq6_add(fq6_mul_by_nonresidue(bb), aa)
    return (c0, c1)


def fq12_sqr(a):
    """fq12.rs:99-114"""
    ab = fq6_mul(a[0], a[1])
    c0c1 = fq6_add(a[0], a[1])
    c0 = fq6_add(fq6_mul_by_nonresidue(a[1]), a[0])
    c0 = fq6_sub(fq6_mul(c0, c0c1), ab)
    c1 = fq6_add(ab, ab)
    c0 = fq6_sub(c0, fq6_mul_by_nonresidue(ab))
    return (c0, c1)


def fq12_mul_by_014(a, c0, c1, c4):
    """fq12.rs:34-48"""
    aa = fq6_mul_by_01(a[0], c0, c1)
    bb = fq6_mul_by_1(a[1], c4)
    o = fq2_add(c1, c4)
    r1 = fq6_mul_by_01(fq6_add(a[1], a[0]), c0, o)
    r1 = fq6_sub(fq6_sub(r1, aa), bb)
    r0 = fq6_add(fq6_mul_by_nonresidue(bb), aa)
    return (r0, r1)


def fq12_inv(a):
    """fq12.rs:132-148"""
    t = fq6_sub(fq6_sqr(a[0]), fq6_mul_by_nonresidue(fq6_sqr(a[1])))
    t = fq6_inv(t)
    if t is None:
        return None
    return (fq6_mul(t, a[0]), fq6_neg(fq6_mul(t, a[1])))


def fq12_frobenius(a, power):
    """fq12.rs:90-97"""
    c0 = fq6_frobenius(a[0], power)
    c1 = fq6_frobenius(a[1], power)
    k = FROB_FQ12_C1[power % 12]
    c1 = tuple(fq2_mul(x, k) for x in c1)
    return (c0, c1)


def fq12_pow(a, e):
    """lib.rs:306-324 (Field::pow)."""
    r = FQ12_ONE
    found = False
    for bit in bin(e)[2:].zfill(64):
        if found:
            r = fq12_sqr(r)
        else:
            found = bit == '1'
        if bit == '1':
            r = fq12_mul(r, a)
    return r


# ----------------------------------------------------------------------------------------------
# Curve groups, Jacobian (ec.rs:216-619).  F is a field-op namespace so G1 and G2 share the code,
# exactly as the reference's `curve_impl!` macro does.
# ----------------------------------------------------------------------------------------------
class _F1:
    zero, one = 0, 1
    add, sub, neg, dbl, mul, sqr, inv = (staticmethod(f) for f in
                                         (fq_add, fq_sub, fq_neg, fq_dbl, fq_mul, fq_sqr, fq_inv))
    is_zero = staticmethod(lambda a: a == 0)


class _F2:
    zero, one = FQ2_ZERO, FQ2_ONE
    add, sub, neg, dbl, mul, sqr, inv = (staticmethod(f) for f in
                                         (fq2_add, fq2_sub, fq2_neg, fq2_dbl, fq2_mul, fq2_sqr, fq2_inv))
    is_zero = staticmethod(fq2_is_zero)


def pt_zero(F):
    """ec.rs:224-230"""
    return (F.zero, F.one, F.zero)


def pt_is_zero(F, p):
    return F.is_zero(p[2])


def pt_double(F, p):
    """dbl-2009-l, ec.rs:296-354"""
    if pt_is_zero(F, p):
        return p
    x, y, z = p
    a = F.sqr(x)
    b = F.sqr(y)
    c = F.sqr(b)
    d = F.dbl(F.sub(F.sub(F.sqr(F.add(x, b)), a), c))
    e = F.add(F.dbl(a), a)
    f = F.sqr(e)
    z3 = F.dbl(F.mul(z, y))
    x3 = F.sub(F.sub(f, d), d)
    c8 = F.dbl(F.dbl(F.dbl(c)))
    y3 = F.sub(F.mul(F.sub(d, x3), e), c8)
    return (x3, y3, z3)


def pt_add(F, p, o):
    """add-2007-bl, ec.rs:356-444 (incl. copy / no-op / double special cases)."""
    if pt_is_zero(F, p):
        return o
    if pt_is_zero(F, o):
        return p
    x1, y1, z1 = p
    x2, y2, z2 = o
    z1z1 = F.sqr(z1)
    z2z2 = F.sqr(z2)
    u1 = F.mul(x1, z2z2)
    u2 = F.mul(x2, z1z1)
    s1 = F.mul(F.mul(y1, z2), z2z2)
    s2 = F.mul(F.mul(y2, z1), z1z1)
    if u1 == u2 and s1 == s2:
        return pt_double(F, p)
    h = F.sub(u2, u1)
    i = F.sqr(F.dbl(h))
    j = F.mul(h, i)
    r = F.dbl(F.sub(s2, s1))
    v = F.mul(u1, i)
    x3 = F.sub(F.sub(F.sub(F.sqr(r), j), v), v)
    y3 = F.sub(F.mul(F.sub(v, x3), r), F.dbl(F.mul(s1, j)))
    z3 = F.mul(F.sub(F.sub(F.sqr(F.add(z1, z2)), z1z1), z2z2), h)
    return (x3, y3, z3)


def pt_add_mixed(F, p, o):
    """madd-2007-bl, ec.rs:446-526.  o = (x, y, infinity)."""
    if o[2]:
        return p
    if pt_is_zero(F, p):
        return (o[0], o[1], F.one)
    x1, y1, z1 = p
    z1z1 = F.sqr(z1)
    u2 = F.mul(o[0], z1z1)
    s2 = F.mul(F.mul(o[1], z1), z1z1)
    if x1 == u2 and y1 == s2:
        return pt_double(F, p)
    h = F.sub(u2, x1)
    hh = F.sqr(h)
    i = F.dbl(F.dbl(hh))
    j = F.mul(h, i)
    r = F.dbl(F.sub(s2, y1))
    v = F.mul(x1, i)
    x3 = F.sub(F.sub(F.sub(F.sqr(r), j), v), v)
    j2 = F.dbl(F.mul(j, y1))
    y3 = F.sub(F.mul(F.sub(v, x3), r), j2)
    z3 = F.sub(F.sub(F.sqr(F.add(z1, h)), z1z1), hh)
    return (x3, y3, z3)


def pt_negate(F, p):
    """ec.rs:528-532"""
    if pt_is_zero(F, p):
        return p
    return (p[0], F.neg(p[1]), p[2])


def pt_mul(F, p, k):
    """double-and-add, ec.rs:534-553 (MSB-first over 256 bits)."""
    res = pt_zero(F)
    found = False
    for bit in bin(k)[2:].zfill(256):
        if found:
            res = pt_double(F, res)
        else:
            found = bit == '1'
        if bit == '1':
            res = pt_add(F, res, p)
    return res


def pt_to_affine(F, p):
    """ec.rs:586-619.  Returns (x, y, infinity)."""
    if pt_is_zero(F, p):
        return (F.zero, F.one, True)
    if p[2] == F.one:
        return (p[0], p[1], False)
    zinv = F.inv(p[2])
    zi2 = F.sqr(zinv)
    return (F.mul(p[0], zi2), F.mul(p[1], F.mul(zi2, zinv)), False)


def pt_from_affine(F, a):
    """ec.rs:570-582"""
    if a[2]:
        return pt_zero(F)
    return (a[0], a[1], F.one)


def pt_is_normalized(F, p):
    return pt_is_zero(F, p) or p[2] == F.one


def pt_batch_normalization(F, v):
    """ec.rs:246-294.  Returns the new list."""
    v = list(v)
    prod = []
    tmp = F.one
    idx = [i for i, g in enumerate(v) if not pt_is_normalized(F, g)]
    for i in idx:
        tmp = F.mul(tmp, v[i][2])
        prod.append(tmp)
    tmp = F.inv(tmp)
    ss = list(reversed(prod))[1:] + [F.one]
    for i, s in zip(reversed(idx), ss):
        g = v[i]
        newtmp = F.mul(tmp, g[2])
        v[i] = (g[0], g[1], F.mul(tmp, s))
        tmp = newtmp
    for i in idx:
        g = v[i]
        z = F.sqr(g[2])
        x = F.mul(g[0], z)
        z = F.mul(z, g[2])
        y = F.mul(g[1], z)
        v[i] = (x, y, F.one)
    return v


# ----------------------------------------------------------------------------------------------
# wNAF (wnaf.rs:4-71) and window heuristics (ec.rs:895-921, 1586-1612)
# ----------------------------------------------------------------------------------------------
def wnaf_form(c, window):
    """wnaf.rs:18-43"""
    out = []
    while c != 0:
        if c & 1:
            u = c % (1 << (window + 1))
            if u > (1 << window):
                u -= 1 << (window + 1)
            c -= u
        else:
            u = 0
        out.append(u)
        c >>= 1
    return out


def wnaf_table(F, base, window):
    """wnaf.rs:4-15"""
    table = []
    dbl = pt_double(F, base)
    for _ in range(1 << (window - 1)):
        table.append(base)
        base = pt_add(F, base, dbl)
    return table


def wnaf_exp(F, table, wnaf):
    """wnaf.rs:49-71"""
    result = pt_zero(F)
    found = False
    for n in reversed(wnaf):
        if found:
            result = pt_double(F, result)
        if n != 0:
            found = True
            if n > 0:
                result = pt_add(F, result, table[n // 2])
            else:
                result = pt_add(F, result, pt_negate(F, table[(-n) // 2]))
    return result


def g1_recommended_wnaf_for_scalar(k):
    """ec.rs:895-905"""
    nb = k.bit_length()
    return 4 if nb >= 130 else (3 if nb >= 34 else 2)


def g2_recommended_wnaf_for_scalar(k):
    """ec.rs:1586-1596"""
    nb = k.bit_length()
    return 4 if nb >= 103 else (3 if nb >= 37 else 2)


G1_NUM_SCALARS_REC = [1, 3, 7, 20, 43, 120, 273, 563, 1630, 3128, 7933, 62569]      # ec.rs:907-921
G2_NUM_SCALARS_REC = [1, 3, 8, 20, 47, 126, 260, 826, 1501, 4555, 84071]            # ec.rs:1598-1612


def recommended_wnaf_for_num_scalars(rec, n):
    ret = 4
    for r in rec:
        if n > r:
            ret += 1
        else:
            break
    return ret


def wnaf_mul(F, base, k, window=None, g2=False):
    """Wnaf::new().scalar(k).base(g)  (wnaf.rs:111-128, 158-164)."""
    if window is None:
        window = (g2_recommended_wnaf_for_scalar if g2 else g1_recommended_wnaf_for_scalar)(k)
    return wnaf_exp(F, wnaf_table(F, base, window), wnaf_form(k, window))


# ----------------------------------------------------------------------------------------------
# Pairing engine (bls12_381/mod.rs:40-358)
# ----------------------------------------------------------------------------------------------
def _loop_bits():
    """Bits of BLS_X>>1 after the leading one, MSB first (mod.rs:72-78, 337-343)."""
    return [c == '1' for c in bin(BLS_X >> 1)[3:]]


def g2_doubling_step(r):
    """mod.rs:176-245.  Returns (new r, (c0, c1, c2))."""
    rx, ry, rz = r
    tmp0 = fq2_sqr(rx)
    tmp1 = fq2_sqr(ry)
    tmp2 = fq2_sqr(tmp1)
    tmp3 = fq2_dbl(fq2_sub(fq2_sub(fq2_sqr(fq2_add(tmp1, rx)), tmp0), tmp2))
    tmp4 = fq2_add(fq2_dbl(tmp0), tmp0)
    tmp6 = fq2_add(rx, tmp4)
    tmp5 = fq2_sqr(tmp4)
    zsq = fq2_sqr(rz)
    nx = fq2_sub(fq2_sub(tmp5, tmp3), tmp3)
    nz = fq2_sub(fq2_sub(fq2_sqr(fq2_add(rz, ry)), tmp1), zsq)
    ny = fq2_mul(fq2_sub(tmp3, nx), tmp4)
    tmp2 = fq2_dbl(fq2_dbl(fq2_dbl(tmp2)))
    ny = fq2_sub(ny, tmp2)
    tmp3 = fq2_neg(fq2_dbl(fq2_mul(tmp4, zsq)))
    tmp6 = fq2_sub(fq2_sub(fq2_sqr(tmp6), tmp0), tmp5)
    tmp1 = fq2_dbl(fq2_dbl(tmp1))
    tmp6 = fq2_sub(tmp6, tmp1)
    tmp0 = fq2_dbl(fq2_mul(nz, zsq))
    return (nx, ny, nz), (tmp0, tmp3, tmp6)


def g2_addition_step(r, q):
    """mod.rs:247-333.  q = (x, y) affine."""
    rx, ry, rz = r
    qx, qy = q
    zsq = fq2_sqr(rz)
    ysq = fq2_sqr(qy)
    t0 = fq2_mul(zsq, qx)
    t1 = fq2_mul(fq2_sub(fq2_sub(fq2_sqr(fq2_add(qy, rz)), ysq), zsq), zsq)
    t2 = fq2_sub(t0, rx)
    t3 = fq2_sqr(t2)
    t4 = fq2_dbl(fq2_dbl(t3))
    t5 = fq2_mul(t4, t2)
    t6 = fq2_sub(fq2_sub(t1, ry), ry)
    t9 = fq2_mul(t6, qx)
    t7 = fq2_mul(t4, rx)
    nx = fq2_sub(fq2_sub(fq2_sub(fq2_sqr(t6), t5), t7), t7)
    nz = fq2_sub(fq2_sub(fq2_sqr(fq2_add(rz, t2)), zsq), t3)
    t10 = fq2_add(qy, nz)
    t8 = fq2_mul(fq2_sub(t7, nx), t6)
    t0 = fq2_dbl(fq2_mul(ry, t5))
    ny = fq2_sub(t8, t0)
    t10 = fq2_sub(fq2_sqr(t10), ysq)
    ztsq = fq2_sqr(nz)
    t10 = fq2_sub(t10, ztsq)
    t9 = fq2_sub(fq2_dbl(t9), t10)
    t10 = fq2_dbl(nz)
    t6 = fq2_neg(t6)
    t1 = fq2_dbl(t6)
    return (nx, ny, nz), (t10, t1, t9)


def g2_prepare(q):
    """G2Prepared::from_affine, mod.rs:168-358.  q = (x, y, infinity).  Returns (coeffs, infinity)."""
    if q[2]:
        return [], True
    coeffs = []
    r = (q[0], q[1], FQ2_ONE)
    for bit in _loop_bits():
        r, c = g2_doubling_step(r)
        coeffs.append(c)
        if bit:
            r, c = g2_addition_step(r, (q[0], q[1]))
            coeffs.append(c)
    r, c = g2_doubling_step(r)
    coeffs.append(c)
    return coeffs, False


def _ell(f, coeffs, p):
    """mod.rs:57-69"""
    c0 = fq2_mul_fq(coeffs[0], p[1])
    c1 = fq2_mul_fq(coeffs[1], p[0])
    return fq12_mul_by_014(f, coeffs[2], c1, c0)


def miller_loop(pairs):
    """mod.rs:40-102.  pairs = [(p_affine, (coeffs, infinity))]."""
    live = [(p, iter(pr[0])) for p, pr in pairs if not p[2] and not pr[1]]
    f = FQ12_ONE
    for bit in _loop_bits():
        for p, it in live:
            f = _ell(f, next(it), p)
        if bit:
            for p, it in live:
                f = _ell(f, next(it), p)
        f = fq12_sqr(f)
    for p, it in live:
        f = _ell(f, next(it), p)
    if BLS_X_IS_NEGATIVE:
        f = fq12_conjugate(f)
    return f


def _exp_by_x(f, x):
    """mod.rs:116-121"""
    f = fq12_pow(f, x)
    if BLS_X_IS_NEGATIVE:
        f = fq12_conjugate(f)
    return f


def final_exponentiation(r):
    """mod.rs:104-160.  None iff r == 0."""
    f1 = fq12_conjugate(r)
    f2 = fq12_inv(r)
    if f2 is None:
        return None
    r = fq12_mul(f1, f2)
    f2 = r
    r = fq12_mul(fq12_frobenius(r, 2), f2)
    x = BLS_X
    y0 = fq12_sqr(r)
    y1 = _exp_by_x(y0, x)
    y2 = _exp_by_x(y1, x >> 1)
    y3 = fq12_conjugate(r)
    y1 = fq12_mul(y1, y3)
    y1 = fq12_conjugate(y1)
    y1 = fq12_mul(y1, y2)
    y2 = _exp_by_x(y1, x)
    y3 = _exp_by_x(y2, x)
    y1 = fq12_conjugate(y1)
    y3 = fq12_mul(y3, y1)
    y1 = fq12_conjugate(y1)
    y1 = fq12_frobenius(y1, 3)
    y2 = fq12_frobenius(y2, 2)
    y1 = fq12_mul(y1, y2)
    y2 = _exp_by_x(y3, x)
    y2 = fq12_mul(y2, y0)
    y2 = fq12_mul(y2, r)
    y1 = fq12_mul(y1, y2)
    y2 = fq12_frobenius(y3, 1)
    y1 = fq12_mul(y1, y2)
    return y1


def pairing(p, q):
    """Engine::pairing on affine inputs, lib.rs:101-109."""
    return final_exponentiation(miller_loop([(p, g2_prepare(q))]))


G1_GEN_AFFINE = (G1_X, G1_Y, False)
G2_GEN_AFFINE = (G2_X, G2_Y, False)


# ----------------------------------------------------------------------------------------------
# Byte-level (de)serialisation used by the fixtures and the C/CUDA ABI: Montgomery LE u64 limbs.
# ----------------------------------------------------------------------------------------------
def fq_to_bytes(a):
    return to_mont(a).to_bytes(48, 'little')


def fq_from_bytes(b):
    return from_mont(int.from_bytes(b, 'little'))


def flatten(x):
    """Nested tuples of ints -> flat list of Fq values in the reference's struct order."""
    if isinstance(x, int):
        return [x]
    out = []
    for y in x:
        out.extend(flatten(y))
    return out


def to_bytes(x):
    return b''.join(fq_to_bytes(v) for v in flatten(x))


def fq2_from_bytes(b): return (fq_from_bytes(b[:48]), fq_from_bytes(b[48:96]))
def fq6_from_bytes(b): return tuple(fq2_from_bytes(b[96 * i:96 * i + 96]) for i in range(3))
def fq12_from_bytes(b): return (fq6_from_bytes(b[:288]), fq6_from_bytes(b[288:576]))


# ----------------------------------------------------------------------------------------------
# Point encodings and their validation (the step BEFORE the path in batch
# verification -- bytes -> affine points).  ec.rs:97-144 (get_point_from_x, is_on_curve, subgroup check),
# ec.rs:645-868 (G1Uncompressed / G1Compressed), ec.rs:1292-1540 (G2), fq.rs:1147-1170 (Fq::sqrt),
# fq2.rs:167-221 (Fq2::sqrt), fq.rs:703-708 + fq2.rs:20-31 (orderings).
# Pinned by the reference's k*G vector files and its invalid-vector tests (bls12_381/tests/mod.rs:55-611).
# ----------------------------------------------------------------------------------------------
DEC_OK, DEC_UNEXPECTED_COMPRESSION_MODE, DEC_UNEXPECTED_INFORMATION, DEC_NOT_ON_CURVE, DEC_NOT_IN_SUBGROUP = 0, 1, 2, 3, 4
DEC_COORDINATE = 16          # + index: G1 0 = "x coordinate", 1 = "y coordinate";
#                              G2 0 = "x coordinate (c0)", 1 = "x coordinate (c1)", 2 = "y coordinate (c0)", 3 = "y coordinate (c1)"
SQRT_EXP = (Q - 3) // 4      # fq.rs:1152-1159
assert SQRT_EXP == from_limbs64([0xee7fbfffffffeaaa, 0x7aaffffac54ffff, 0xd9cc34a83dac3d89, 0xd91dd2e13ce144af, 0x92c6e9ed90d2eb35, 0x680447a8e5ff9a6])
G1_B = 4                     # y^2 = x^3 + 4
G2_B = (4, 4)                # y^2 = x^3 + 4(1 + u)


def fq_sqrt(a):
    """fq.rs:1147-1170 (Shanks for q = 3 mod 4): returns the specific root the reference returns, or None."""
    a1 = pow(a, SQRT_EXP, Q)
    a0 = a1 * a1 % Q * a % Q
    if a0 == Q - 1:
        return None
    return a1 * a % Q


def fq2_sqrt(a):
    """fq2.rs:167-221 (Algorithm 9 of eprint 2012/685)."""
    if fq2_is_zero(a):
        return FQ2_ZERO
    a1 = fq2_pow(a, SQRT_EXP)
    alpha = fq2_mul(fq2_sqr(a1), a)
    a0 = fq2_mul(fq2_frobenius(alpha, 1), alpha)
    neg1 = (Q - 1, 0)
    if a0 == neg1:
        return None
    a1 = fq2_mul(a1, a)
    if alpha == neg1:
        return fq2_mul(a1, (0, 1))
    alpha = fq2_pow(fq2_add(alpha, FQ2_ONE), (Q - 1) // 2)
    return fq2_mul(a1, alpha)


def _fq2_key(a):             # Fq2 ordering: c1 first, then c0 (fq2.rs:20-31)
    return (a[1], a[0])


def get_point_from_x(x, greatest, g2):
    """ec.rs:102-123"""
    if g2:
        y = fq2_sqrt(fq2_add(fq2_mul(fq2_sqr(x), x), G2_B))
        if y is None:
            return None
        negy = fq2_neg(y)
        return (x, y if (_fq2_key(y) < _fq2_key(negy)) ^ greatest else negy, False)
    y = fq_sqrt((x * x % Q * x + G1_B) % Q)
    if y is None:
        return None
    negy = fq_neg(y)
    return (x, y if (y < negy) ^ greatest else negy, False)


def is_on_curve(p, g2):
    """ec.rs:125-140"""
    x, y, inf = p
    if inf:
        return True
    if g2:
        return fq2_sqr(y) == fq2_add(fq2_mul(fq2_sqr(x), x), G2_B)
    return y * y % Q == (x * x % Q * x + G1_B) % Q


def is_in_correct_subgroup_assuming_on_curve(p, g2):
    """ec.rs:142-144: self.mul(Fr::char()).is_zero()"""
    F = _F2 if g2 else _F1
    return pt_is_zero(F, pt_mul(F, pt_from_affine(F, p), R_ORDER))


def _affine_zero(g2):
    return (FQ2_ZERO, FQ2_ONE, True) if g2 else (0, 1, True)


def decode_point(b, g2, compressed, checked=True):
    """EncodedPoint::into_affine / into_affine_unchecked.  Returns (status, affine or None)."""
    b = bytearray(b)
    ncoord = (2 if g2 else 1) * (1 if compressed else 2)
    assert len(b) == 48 * ncoord
    if bool(b[0] & 0x80) != bool(compressed):
        return DEC_UNEXPECTED_COMPRESSION_MODE, None
    if b[0] & 0x40:
        b[0] &= 0x3f
        if any(b):
            return DEC_UNEXPECTED_INFORMATION, None
        return DEC_OK, _affine_zero(g2)
    greatest = bool(b[0] & 0x20)
    if greatest and not compressed:
        return DEC_UNEXPECTED_INFORMATION, None
    b[0] &= 0x1f
    vals = [int.from_bytes(b[48 * i:48 * i + 48], "big") for i in range(ncoord)]
    if g2:
        # bytes hold c1 then c0; the reference converts c0 first, then c1 (ec.rs:1378-1393, 1480-1487)
        order = [(1, 0), (0, 1)] + ([(3, 2), (2, 3)] if not compressed else [])     # (byte slot, coordinate index)
        for slot, idx in order:
            if vals[slot] >= Q:
                return DEC_COORDINATE + idx, None
        x = (vals[1], vals[0])
        if compressed:
            p = get_point_from_x(x, greatest, True)
            if p is None:
                return DEC_NOT_ON_CURVE, None
        else:
            p = (x, (vals[3], vals[2]), False)
    else:
        for idx, v in enumerate(vals):
            if v >= Q:
                return DEC_COORDINATE + idx, None
        if compressed:
            p = get_point_from_x(vals[0], greatest, False)
            if p is None:
                return DEC_NOT_ON_CURVE, None
        else:
            p = (vals[0], vals[1], False)
    if checked:
        if not compressed and not is_on_curve(p, g2):
            return DEC_NOT_ON_CURVE, None
        if not is_in_correct_subgroup_assuming_on_curve(p, g2):
            return DEC_NOT_IN_SUBGROUP, None
    return DEC_OK, p


def encode_point(p, g2, compressed):
    """EncodedPoint::from_affine (ec.rs:739-757, 846-867, 1397-1416, 1519-1540)."""
    x, y, inf = p
    ncoord = (2 if g2 else 1) * (1 if compressed else 2)
    b = bytearray(48 * ncoord)
    if inf:
        b[0] |= 0x40
    else:
        e = lambda v: v.to_bytes(48, "big")
        if g2:
            b[:96] = e(x[1]) + e(x[0])
            if not compressed:
                b[96:] = e(y[1]) + e(y[0])
            greatest = _fq2_key(y) > _fq2_key(fq2_neg(y))
        else:
            b[:48] = e(x)
            if not compressed:
                b[48:] = e(y)
            greatest = y > fq_neg(y)
        if compressed and greatest:
            b[0] |= 0x20
    if compressed:
        b[0] |= 0x80
    return bytes(b)


# ----------------------------------------------------------------------------------------------
# The scalar field Fr (bls12_381/fr.rs:324-572), Montgomery form with R = 2^256 -- the step after the path
# (SURVEY.md 8f item 4).  Pinned by the reference's Montgomery-form KATs (fr.rs:1240-1260, 1306-1323, 1450-1486).
# ----------------------------------------------------------------------------------------------
FR_MONT_R = (1 << 256) % R_ORDER            # fr.rs:20-26
FR_MONT_R2 = pow(1 << 256, 2, R_ORDER)      # fr.rs:28-34
FR_INV64 = (-pow(R_ORDER, -1, 1 << 64)) % (1 << 64)   # fr.rs:36
FR_RINV = pow(FR_MONT_R, -1, R_ORDER)


def fr_to_mont(x): return x * FR_MONT_R % R_ORDER
def fr_from_mont(x): return x * FR_RINV % R_ORDER


def fr_op_mont(op, a, b=None):
    """The reference's Fr operation on MONTGOMERY-FORM integers a, b (< r).  Returns (value, ok)."""
    if op == "add": return (a + b) % R_ORDER, True
    if op == "sub": return (a - b) % R_ORDER, True
    if op == "mul": return a * b * FR_RINV % R_ORDER, True
    if op == "sqr": return a * a * FR_RINV % R_ORDER, True
    if op == "neg": return (-a) % R_ORDER, True
    if op == "dbl": return 2 * a % R_ORDER, True
    if op == "inv":
        if a == 0:
            return 0, False
        return fr_to_mont(pow(fr_from_mont(a), -1, R_ORDER)), True
    if op == "from_repr":                       # a is a canonical integer (possibly >= r): fr.rs:279-288
        return (fr_to_mont(a), True) if a < R_ORDER else (0, False)
    if op == "into_repr": return fr_from_mont(a), True
    raise ValueError(op)


# ----------------------------------------------------------------------------------------------
# G::rand minus the random number generator (ec.rs:199-214): get_point_from_x (above) then scale_by_cofactor
# (ec.rs:86-94 mul_bits, cofactors ec.rs:871-875 / 1564-1578); the caller retries on None / infinity.
# ----------------------------------------------------------------------------------------------
G1_COFACTOR = from_limbs64([0x8c00aaab0000aaab, 0x396c8c005555e156])
G2_COFACTOR = from_limbs64([0xcf1c38e31c7238e5, 0x1616ec6e786f0c70, 0x21537e293a6691ae, 0xa628f1cb4d9e82ef,
                            0xa68a205b2e5a7ddf, 0xcd91de4547085aba, 0x91d50792876a202, 0x5d543a95414e7f1])
assert G1_COFACTOR == 76329603384216526031706109802092473003


def scale_by_cofactor(p, g2):
    """affine p -> Jacobian [h]p by the reference's mul_bits: MSB-first over ALL bits of the limb array, mixed additions."""
    F = _F2 if g2 else _F1
    res = pt_zero(F)
    nbits = 512 if g2 else 128
    cof = G2_COFACTOR if g2 else G1_COFACTOR
    for i in range(nbits - 1, -1, -1):
        res = pt_double(F, res)
        if (cof >> i) & 1:
            res = pt_add_mixed(F, res, p)
    return res


def affine_mul(p, k, g2):
    """CurveAffine::mul (ec.rs:174-177): mul_bits over all 256 bits of the FrRepr, MSB first, mixed additions."""
    F = _F2 if g2 else _F1
    res = pt_zero(F)
    for i in range(255, -1, -1):
        res = pt_double(F, res)
        if (k >> i) & 1:
            res = pt_add_mixed(F, res, p)
    return res

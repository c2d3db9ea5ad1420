#!/usr/bin/env python3
"""Generator + word-level emulator of the dedicated Montgomery SQUARING of pairing_b200/csrc/fp_sqr_gen.cuh.

The reference squares with a distinct routine (bls12_381/fq.rs:963-1016: off-diagonal products once, doubled,
plus the diagonal, then mont_reduce).  Same structure here on 12 x 32-bit limbs:

  1. S = sum_{i<j} a_i a_j 2^(32(i+j))  -- 66 wide multiply-accumulates.  As in fp_mul the products are kept in two
     accumulators, E (products at even word positions) and O (odd positions, O[k] <-> word k+1), so that every
     `mad.lo.cc / madc.hi.cc` pair hits an even-aligned register pair and fuses into one IMAD.WIDE.U32.X.
  2. S = E + (O << 32); T = 2 S (24 funnel shifts) + sum_i a_i^2 2^(64 i) (12 fused multiply-accumulates).
  3. r = (T_lo + M q) / 2^384 + T_hi: twelve reduction rows over the LOW half only (144 MACs + 12 IMAD for the m's),
     again in the even/odd form, the 64-bit down-shift fused into the multiply-accumulate chain; then the high half
     is added.  r <= q + (2q)^2 / 2^384 < 1.41 q + 1 for a <= 2q: inside the relaxed range, no conditional subtraction.

78 + 144 = 222 wide MACs + 12 IMAD = 234 MAC32 against 300 for fp_mul(a, a).

ONE instruction list drives both the PTX emitter and the emulator below, which executes it on Python integers with the
PTX carry-flag semantics and asserts (a) the value identity, (b) the output bound and (c) that every carry a non-`.cc`
instruction drops is zero.  `python tools/gen_fp_sqr.py` rewrites the header; `--check` verifies the committed header
is what this script generates and runs the emulation (tests/test_host_logic.py calls it)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class m:   # the base-field modulus (bls12_381/fq.rs:6-13) -- a public constant, no oracle import in build tooling
    Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


M32 = (1 << 32) - 1
Q = [(m.Q >> (32 * i)) & M32 for i in range(12)]
NINV = (-pow(m.Q, -1, 1 << 32)) & M32
assert NINV == 0xfffcfffd
OUT = os.path.join(ROOT, "pairing_b200", "csrc", "fp_sqr_gen.cuh")


class Prog:
    """a straight-line program over named 32-bit registers; operands are register names, ints (immediates) or 'Z' (zero)"""

    def __init__(self):
        self.ins = []
        self.temps = []

    def reg(self, name):
        if name not in self.temps:
            self.temps.append(name)
        return name

    def emit(self, op, d, *src):
        self.ins.append((op, d, src))


def build():
    p = Prog()
    a = ["a%d" % i for i in range(12)]
    E = [p.reg("e%d" % k) for k in range(24)]
    O = [p.reg("o%d" % k) for k in range(24)]
    fresh_e, fresh_o = [True] * 24, [True] * 24

    def chain(acc, fresh, first_idx, i, js):
        """acc[first_idx ...] += a_i * a_j for j in js (consecutive register pairs), one carry chain"""
        k = first_idx
        top = first_idx + 2 * len(js) - 1
        top_was_fresh = fresh[top]
        for n, j in enumerate(js):
            for half in ("lo", "hi"):
                addend = "Z" if fresh[k] else acc[k]
                p.emit(("mad.%s.cc" if (n == 0 and half == "lo") else "madc.%s.cc") % half, acc[k], a[i], a[j], addend)
                fresh[k] = False
                k += 1
        if top_was_fresh:
            # the last high word is hi(product) + carry <= 0xfffffffe + 1: nothing can carry out (the emulator asserts it)
            p.emit("drop_carry", None)
        else:
            assert fresh[k], "carry word of a chain must be untouched"
            p.emit("addc", acc[k], "Z", "Z")
            fresh[k] = False

    for i in range(11):
        other = [j for j in range(i + 1, 12, 2)]     # i + j odd  -> O index (i+j-1, i+j), first pair at O[2i]
        same = [j for j in range(i + 2, 12, 2)]      # i + j even -> E words (i+j, i+j+1), first pair at E[2i+2]
        if other:
            chain(O, fresh_o, 2 * i, i, other)
        if same:
            chain(E, fresh_e, 2 * i + 2, i, same)
    # S = E + (O << 32): word w of S = E[w] + O[w-1]; E spans words 2..21 (+ carries), O index 0..21 -> words 1..22
    S = [p.reg("s%d" % k) for k in range(24)]
    p.emit("mov", S[0], "Z")                          # no product lands on word 0
    first = True
    for w in range(1, 24):
        ew = "Z" if fresh_e[w] else E[w]
        ow = "Z" if (w - 1 > 23 or fresh_o[w - 1]) else O[w - 1]
        if w == 23:
            p.emit("addc", S[w], ew, ow)
        else:
            p.emit("add.cc" if first else "addc.cc", S[w], ew, ow)
        first = False
    # T = 2 S: funnel shifts (independent instructions, no carry chain)
    T = [p.reg("t%d" % k) for k in range(24)]
    for w in range(23, 0, -1):
        p.emit("shf.l", T[w], S[w - 1], S[w])          # (S[w] << 1) | (S[w-1] >> 31)
    p.emit("shl1", T[0], S[0])
    p.emit("assert_top_bit_clear", None, S[23])
    # T += sum a_i^2 2^(64 i): one chain of 12 fused multiply-accumulates
    for i in range(12):
        p.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", T[2 * i], a[i], a[i], T[2 * i])
        p.emit("madc.hi.cc" if i < 11 else "madc.hi", T[2 * i + 1], a[i], a[i], T[2 * i + 1])
    # Montgomery reduction of the low half in the even/odd form.  L = low-aligned accumulator (12 words), H = the one
    # aligned one word higher.  Start: L = T[0..11], H = 0.
    L = [T[k] for k in range(12)]
    H = [p.reg("h%d" % k) for k in range(12)]
    mreg = p.reg("m")
    qodd = [Q[1], Q[3], Q[5], Q[7], Q[9], Q[11]]
    qeven = [Q[0], Q[2], Q[4], Q[6], Q[8], Q[10]]
    # row 0: m = L[0] * ninv; H = m * q_odd (fresh); L += m * q_even, carry -> H[11]
    p.emit("mul.lo", mreg, L[0], NINV)
    for k in range(6):
        p.emit("mul.lo", H[2 * k], mreg, qodd[k])
        p.emit("mul.hi", H[2 * k + 1], mreg, qodd[k])
    for k in range(6):
        p.emit(("mad.lo.cc" if k == 0 else "madc.lo.cc"), L[2 * k], mreg, qeven[k], L[2 * k])
        p.emit("madc.hi.cc", L[2 * k + 1], mreg, qeven[k], L[2 * k + 1])
    p.emit("addc", H[11], H[11], "Z")
    p.emit("assert_zero", None, L[0])
    for row in range(1, 12):
        # roles swap: H is now the low-aligned accumulator; (L >> 64) becomes the high one; L[1] folds into H[0]
        p.emit("add.cc", H[0], H[0], L[1])
        p.emit("mul.lo", mreg, H[0], NINV)
        newH = [p.reg("r%d_%d" % (row, k)) for k in range(12)]
        for k in range(6):
            lo_add = L[2 * k + 2] if 2 * k + 2 < 12 else "Z"
            hi_add = L[2 * k + 3] if 2 * k + 3 < 12 else "Z"
            p.emit("madc.lo.cc", newH[2 * k], mreg, qodd[k], lo_add)
            p.emit("madc.hi.cc" if k < 5 else "madc.hi", newH[2 * k + 1], mreg, qodd[k], hi_add)
        for k in range(6):
            p.emit(("mad.lo.cc" if k == 0 else "madc.lo.cc"), H[2 * k], mreg, qeven[k], H[2 * k])
            p.emit("madc.hi.cc", H[2 * k + 1], mreg, qeven[k], H[2 * k + 1])
        p.emit("addc", newH[11], newH[11], "Z")
        p.emit("assert_zero", None, H[0])
        L, H = H, newH
    # low result = (L >> 32) + H, plus the high half T[12..23]
    R = ["r%d" % k for k in range(12)]
    tmp = [p.reg("u%d" % k) for k in range(12)]
    for k in range(12):
        l = L[k + 1] if k + 1 < 12 else "Z"
        op = "add.cc" if k == 0 else ("addc.cc" if k < 11 else "addc")
        p.emit(op, tmp[k], l, H[k])
    for k in range(12):
        op = "add.cc" if k == 0 else ("addc.cc" if k < 11 else "addc")
        p.emit(op, R[k], tmp[k], T[12 + k])
    return p


# ------------------------------------------------------------------------------------------------ emulator
def run(p, aval):
    regs = {"Z": 0}
    for i in range(12):
        regs["a%d" % i] = (aval >> (32 * i)) & M32
    cf = 0

    def g(x):
        return x if isinstance(x, int) else regs[x]
    for op, d, src in p.ins:
        if op == "drop_carry":
            assert cf == 0, "a chain ended on untouched words but carried out"
            continue
        if op == "assert_zero":
            assert g(src[0]) == 0
            continue
        if op == "assert_top_bit_clear":
            assert g(src[0]) >> 31 == 0
            continue
        if op == "mov":
            regs[d] = g(src[0]); continue
        if op == "shl1":
            regs[d] = (g(src[0]) << 1) & M32; continue
        if op == "shf.l":
            regs[d] = ((g(src[1]) << 1) | (g(src[0]) >> 31)) & M32; continue
        if op == "mul.lo":
            regs[d] = (g(src[0]) * g(src[1])) & M32; continue
        if op == "mul.hi":
            regs[d] = (g(src[0]) * g(src[1])) >> 32; continue
        base, *flags = op.split(".")
        if base in ("mad", "madc"):
            half = flags[0]
            prod = g(src[0]) * g(src[1])
            part = (prod & M32) if half == "lo" else (prod >> 32)
            t = part + g(src[2]) + (cf if base == "madc" else 0)
        elif base in ("add", "addc"):
            t = g(src[0]) + g(src[1]) + (cf if base == "addc" else 0)
        else:
            raise ValueError(op)
        regs[d] = t & M32
        if "cc" in flags:
            cf = t >> 32
        else:
            assert t >> 32 == 0, "dropped carry in %s %s" % (op, d)
    return sum(regs["r%d" % k] << (32 * k) for k in range(12))


def check(p, trials=3000, seed=1):
    rnd = random.Random(seed)
    q = m.Q
    edge = [0, 1, q - 1, q, q + 1, 2 * q - 1, 2 * q, (1 << 381), (1 << 382) - 1 if (1 << 382) - 1 <= 2 * q else 2 * q,
            int("ffffffff" * 12, 16) % (2 * q + 1), sum(M32 << (32 * i) for i in range(0, 12, 2)) % (2 * q + 1)]
    # operands with all-ones limbs inside the range exercise the longest carries
    for i in range(12):
        edge.append(((1 << (32 * (i + 1))) - 1) % (2 * q + 1))
        edge.append((2 * q) - ((1 << (32 * i)) - 1))
    rinv = pow(1 << 384, -1, q)
    worst = 0
    for t in range(trials + len(edge)):
        a = edge[t] if t < len(edge) else rnd.randrange(0, 2 * q + 1)
        r = run(p, a)
        assert r % q == a * a * rinv % q, "wrong value for a = %x" % a
        assert r <= 2 * q, "result above 2q"
        worst = max(worst, r)
    return worst / q


# ------------------------------------------------------------------------------------------------ PTX emitter
def ptx(p):
    def o(x):
        if x == "Z":
            return "0"
        if isinstance(x, int):
            return "0x%08x" % x
        if x.startswith("a") and x[1:].isdigit():
            return "%%%d" % (12 + int(x[1:]))
        if x.startswith("r") and x[1:].isdigit():
            return "%%%d" % int(x[1:])
        return x
    lines = []
    temps = [t for t in p.temps]
    for i in range(0, len(temps), 12):
        lines.append(".reg .u32 " + ", ".join(temps[i:i + 12]) + ";")
    for op, d, src in p.ins:
        if op in ("drop_carry", "assert_zero", "assert_top_bit_clear"):
            continue
        if op == "mov":
            lines.append("mov.u32 %s, %s;" % (o(d), o(src[0])))
        elif op == "shl1":
            lines.append("shl.b32 %s, %s, 1;" % (o(d), o(src[0])))
        elif op == "shf.l":
            lines.append("shf.l.wrap.b32 %s, %s, %s, 1;" % (o(d), o(src[0]), o(src[1])))
        else:
            lines.append("%s.u32 %s, %s;" % (op, o(d), ", ".join(o(s) for s in src)))
    return lines


HEADER = '''// fp_sqr_gen.cuh -- GENERATED by tools/gen_fp_sqr.py (do not edit; `python tools/gen_fp_sqr.py` rewrites it,
// `--check` verifies it and replays the instruction list on integers: value, bound <= 2q, no dropped carry).
// Dedicated Montgomery squaring (the reference's fq.rs:963-1016 is a distinct routine too): 66 off-diagonal products,
// doubled, + 12 diagonal products + 144-MAC reduction of the low half + 12 IMAD = 234 MAC32 (fp_mul(a, a): 300).
// Operand and result in the relaxed range [0, 2q].
#pragma once
namespace bls {
__device__ __forceinline__ Fp fp_sqr_inline(const Fp& a) {
  Fp r;
  asm("{\\n\\t"
'''


def render(p):
    body = "".join('      "%s\\n\\t"\n' % l for l in ptx(p))
    outs = ", ".join('"=r"(r.v[%d])' % k for k in range(12))
    ins = ", ".join('"r"(a.v[%d])' % k for k in range(12))
    tail = '      "}"\n      : %s\n      : %s);\n  return r;\n}\n}  // namespace bls\n' % (outs, ins)
    return HEADER + body + tail


def main():
    p = build()
    nmac = sum(1 for op, d, s in p.ins if op.startswith("mad") and ".lo" in op) + sum(1 for op, d, s in p.ins if op == "mul.lo" and d != "m" and isinstance(s[1], int))
    text = render(p)
    if "--check" in sys.argv:
        assert open(OUT).read() == text, "fp_sqr_gen.cuh is stale: run python tools/gen_fp_sqr.py"
        w = check(p, trials=400)
        print("fp_sqr schedule ok: %d wide MACs, worst result %.3f q" % (nmac, w))
        return
    w = check(p)
    open(OUT, "w").write(text)
    print("wrote %s: %d instructions, %d wide MACs, worst result %.3f q" % (OUT, len(ptx(p)), nmac, w))


if __name__ == "__main__":
    main()

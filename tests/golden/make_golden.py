#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the hot path into small committed fixtures.

Run HERE (the container that has /root/reference); the GPU box never reads /root/reference.
Only test DATA is extracted (numbers quoted inside the reference's #[test] functions and its
binary .dat vector files), never source code.

  reference_kats.json    every FqRepr([..6 limbs..]) literal, in order of appearance, of the listed
                         #[test] functions and constants (file:line recorded per entry)
  g{1,2}_{un,}compressed_multiples.bin   src/bls12_381/tests/*.dat: encodings of k*G, k = 0..999
"""
import json, os, re, shutil, sys

REF = "/root/reference/src/bls12_381"
OUT = os.path.dirname(os.path.abspath(__file__))

REPR = re.compile(r"FqRepr\(\[\s*((?:0x[0-9a-fA-F]+\s*,?\s*){6})\]\)")


def reprs(text):
    out = []
    for mm in REPR.finditer(text):
        limbs = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", mm.group(1))]
        out.append("%x" % sum(l << (64 * i) for i, l in enumerate(limbs)))
    return out


def fn_body(src, name):
    i = src.index("fn %s(" % name)
    line = src.count("\n", 0, i) + 1
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        c = src[k]
        if c == "{": depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0: break
        k += 1
    return src[j:k + 1], line


def const_body(src, name):
    i = src.index("const %s:" % name)
    line = src.count("\n", 0, i) + 1
    e = src.index(" = ", i)
    j = src.index(";\n", e)          # the terminating ';' of the item (skips the one inside `[Fq; 2]`)
    return src[e:j], line


kats = {}
for fname, fns in {
    "fq.rs": ["test_fq_mul_assign", "test_fq_squaring", "test_fq_from_into_repr", "test_fq_add_assign",
              "test_fq_sub_assign", "test_fq_inverse", "test_fq_double", "test_fq_negate"],
    "fq2.rs": ["test_fq2_squaring", "test_fq2_mul", "test_fq2_inverse", "test_fq2_addition", "test_fq2_subtraction",
               "test_fq2_negation", "test_fq2_doubling", "test_fq2_frobenius_map"],
    "ec.rs": ["test_g1_addition_correctness", "test_g1_doubling_correctness", "test_g1_same_y",
              "test_g2_addition_correctness", "test_g2_doubling_correctness"],
}.items():
    src = open(os.path.join(REF, fname)).read()
    for fn in fns:
        body, line = fn_body(src, fn)
        kats[fn] = {"source": "src/bls12_381/%s:%d" % (fname, line), "reprs": reprs(body)}

src = open(os.path.join(REF, "fq.rs")).read()
for c in ["MODULUS", "R", "R2", "NEGATIVE_ONE", "B_COEFF", "G1_GENERATOR_X", "G1_GENERATOR_Y", "G2_GENERATOR_X_C0",
          "G2_GENERATOR_X_C1", "G2_GENERATOR_Y_C0", "G2_GENERATOR_Y_C1"]:
    body, line = const_body(src, c)
    kats["const_" + c] = {"source": "src/bls12_381/fq.rs:%d" % line, "reprs": reprs(body)}
mm = re.search(r"const INV: u64 = (0x[0-9a-f]+);", src)
kats["const_INV"] = {"source": "src/bls12_381/fq.rs:43", "reprs": ["%x" % int(mm.group(1), 16)]}
# Frobenius tables: every Fq2 entry in order (fq.rs:139-498)
for c in ["FROBENIUS_COEFF_FQ2_C1", "FROBENIUS_COEFF_FQ6_C1", "FROBENIUS_COEFF_FQ6_C2", "FROBENIUS_COEFF_FQ12_C1"]:
    body, line = const_body(src, c)
    kats["const_" + c] = {"source": "src/bls12_381/fq.rs:%d" % line, "reprs": reprs(body)}

# RELIC pairing vector: the 12 decimal coefficients (tests/mod.rs:23-52)
tsrc = open(os.path.join(REF, "tests", "mod.rs")).read()
body, line = fn_body(tsrc, "test_pairing_result_against_relic")
kats["test_pairing_result_against_relic"] = {
    "source": "src/bls12_381/tests/mod.rs:%d" % line,
    "decimal": re.findall(r'from_str\("(\d+)"\)', body)}

json.dump(kats, open(os.path.join(OUT, "reference_kats.json"), "w"), indent=1, sort_keys=True)
for a, b in [("g1_uncompressed_valid_test_vectors.dat", "g1_uncompressed_multiples.bin"),
             ("g1_compressed_valid_test_vectors.dat", "g1_compressed_multiples.bin"),
             ("g2_uncompressed_valid_test_vectors.dat", "g2_uncompressed_multiples.bin"),
             ("g2_compressed_valid_test_vectors.dat", "g2_compressed_multiples.bin")]:
    shutil.copyfile(os.path.join(REF, "tests", a), os.path.join(OUT, b))
print({k: len(v.get("reprs", v.get("decimal"))) for k, v in kats.items()})

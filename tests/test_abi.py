"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/pairing_b200.h declares, struct sizes match the header's layout, and the product
fails loudly (no fallback) when no device is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pairing_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bls_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import pairing_b200._native as nat
    lib = nat.load()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libpairing_b200.so does not export %s" % n
    assert sorted(nat.SYMBOLS) == names, "pairing_b200._native.SYMBOLS is out of sync with the header"
    out = subprocess.check_output(["nm", "-D", "--defined-only", nat.LIB_PATH], text=True)
    exported = set(re.findall(r" T (bls_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_struct_sizes_match_header():
    """A throw-away C program prints sizeof() of every ABI struct; the numpy row widths must agree."""
    import pairing_b200._native as nat
    import tempfile
    src = '#include <stdio.h>\n#include "pairing_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(bls_fq),sizeof(bls_fq2),sizeof(bls_fq6),sizeof(bls_fq12),sizeof(bls_g1_affine),sizeof(bls_g1),' \
          'sizeof(bls_g2_affine),sizeof(bls_g2),sizeof(bls_fr_repr),sizeof(bls_g2_prepared));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        sizes = [int(x) for x in subprocess.check_output([exe], text=True).split()]
    want = [8 * w for w in (nat.W_FQ, nat.W_FQ2, nat.W_FQ6, nat.W_FQ12, nat.W_G1A, nat.W_G1, nat.W_G2A, nat.W_G2, nat.W_FR, nat.W_G2P)]
    assert sizes == want == [48, 96, 288, 576, 104, 144, 200, 288, 32, 19592]


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the context cannot be created and the API raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pairing_b200
    with pytest.raises(pairing_b200.BlsError):
        pairing_b200.Context(0)
    lib = pairing_b200.load()
    err = ctypes.c_int(0)
    assert not lib.bls_ctx_create(0, ctypes.byref(err))
    assert err.value == -2 and b"no CPU fallback" in lib.bls_strerror(err.value)
    # NULL context -> invalid argument, not a crash
    assert lib.bls_pairing_batch(None, None, None, None, 4) == -1


def test_product_package_never_imports_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "pairing_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "pairing_b200", fn)).read()
            assert "oracle_lib" not in src and "bls_model" not in src and "import oracle" not in src, fn
    for fn in os.listdir(os.path.join(ROOT, "pairing_b200", "csrc")):
        if os.path.isdir(os.path.join(ROOT, "pairing_b200", "csrc", fn)):
            continue                                   # _obj/: build products
        src = open(os.path.join(ROOT, "pairing_b200", "csrc", fn)).read()
        assert "oracle/" not in src and "bls_oracle" not in src, fn


def _sass_histogram(function_substr):
    """opcode histogram of one kernel of the built library (cuobjdump -sass)"""
    import collections
    import shutil
    import pairing_b200._native as nat
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.check_output([exe, "-sass", nat.LIB_PATH], text=True)
    hist, on = collections.Counter(), False
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            on = function_substr in m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if on and m:
            hist[m.group(1)] += 1
    return hist


def test_roofline_denominator_kernel_is_imad_wide_only():
    """the integer-multiply peak is measured by a loop that is SASS-checked to contain only
    IMAD.WIDE.  (A loop-invariant mad.wide is hoisted by ptxas and measures 64-bit adds instead.)"""
    h = _sass_histogram("k_imad_wide_peak")
    assert h["IMAD.WIDE.U32"] >= 256, h
    other_math = sum(v for k, v in h.items() if k.startswith(("IADD3", "LOP3", "SHF", "IMAD.HI", "IMAD.MOV")) or k == "IMAD")
    assert other_math <= 40, h          # prologue/epilogue only


def test_montgomery_product_is_carry_linked_imad_wide():
    """fp_mul (fq.rs:910-960 + 1037-1122) is 288 IMAD.WIDE.U32(.X) + 12 IMAD per product: two products in the loop body"""
    h = _sass_histogram("k_fpmul_peak")
    wide = h["IMAD.WIDE.U32"] + h["IMAD.WIDE.U32.X"] + h["IMAD.HI.U32"]
    assert 2 * 288 <= wide <= 2 * 288 + 16, h


def test_dedicated_squaring_is_222_wide_macs():
    """fp_sqr (fq.rs:963-1016) is the generated 66 + 12 + 144 = 222 IMAD.WIDE.U32(.X) + 12 IMAD routine:
    two independent squarings in the loop body of the measurement kernel"""
    h = _sass_histogram("k_fpsqr_peak")
    wide = h["IMAD.WIDE.U32"] + h["IMAD.WIDE.U32.X"] + h["IMAD.HI.U32"]
    assert 2 * 222 <= wide <= 2 * 222 + 16, h


def _cuobjdump(*args):
    import shutil
    import pairing_b200._native as nat
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    return subprocess.check_output([exe] + list(args) + [nat.LIB_PATH], text=True)


def _sass_text(name_part):
    out = _cuobjdump("-sass")
    keep, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = name_part in line
        elif on:
            keep.append(line)
    return "\n".join(keep)


def test_shared_q_kernel_stages_coefficients_with_a_bulk_copy():
    """k_pair_miller_shared_q loads the 19 584 B of line coefficients by ONE TMA bulk copy completing on an mbarrier:
    UBLKCP + SYNCS in the SASS"""
    sass = _sass_text("k_pair_miller_shared_q")
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_resource_usage_of_the_throughput_kernels():
    """Performance of the headline kernels hangs on ptxas' allocation (abi_common.cuh): pin registers / stack / shared
    memory so that a perturbing change is caught here, before GPU time is spent (profiles/r2_resource_usage.txt is the
    committed snapshot).  VERDICT r1 weak 9."""
    out = _cuobjdump("--dump-resource-usage")
    usage = {}
    name = None
    for line in out.splitlines():
        if "Function " in line:
            name = line.split("Function ")[1].rstrip(":").strip()
        elif "REG:" in line and name:
            usage[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED):(\d+)", line)}
    def of(part):
        hits = [v for k, v in usage.items() if part in k]
        assert hits, part
        return hits[0]
    fused = of("k_pair_millerILb1")
    assert fused["REG"] == 255 and fused["STACK"] <= 5904 and fused["SHARED"] == 0, fused     # 2 blocks x 128 threads per SM; the stack holds the six compressed powers of exp_by_x
    mm = of("k_pair_multi_millerPK")
    assert mm["REG"] == 255 and mm["STACK"] <= 1968 and mm["SHARED"] <= 19456 + 1024, mm
    g1 = of("k_wnaf_mul_lazykIN3bls2FpELb0ELi3ELi8")
    assert g1["REG"] <= 128 and g1["STACK"] <= 5488, g1                                         # 4 blocks per SM
    wide = of("k_wide_pairing")
    assert wide["REG"] <= 128, wide

"""The C++ host mirror (include/pairing_b200.hpp) of the reference's trait surface: compiles against the C ABI on
CPU; on the GPU box the restated reference tests (tests/cpp/test_engine.cpp = src/tests/engine.rs + curve.rs in batch
form, compared bit-exactly with the oracle) run as a native program linked to libpairing_b200.so."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(out):
    import oracle_lib
    import pairing_b200._native as nat
    oracle_lib.lib()                                   # builds oracle/_build/libbls_oracle.so if needed
    libdir = os.path.dirname(nat.LIB_PATH)
    odir = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "test_engine.cpp"),
                           "-L", libdir, "-lpairing_b200", "-L", odir, "-lbls_oracle",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir, "-ldl"])


def test_cpp_host_mirror_compiles_and_links():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "test_engine")
        _build(exe)
        assert os.path.getsize(exe) > 0
        import torch
        if not torch.cuda.is_available():
            # without a device the program must fail loudly at context creation (no CPU fallback)
            r = subprocess.run([exe], capture_output=True, text=True)
            assert r.returncode == 2 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_reference_engine_and_curve_tests_in_cpp():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "test_engine")
        _build(exe)
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "relic_pairing_g1g2.bin")], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "cpp engine tests ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

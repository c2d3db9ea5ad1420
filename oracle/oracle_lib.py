"""ctypes wrapper around oracle/_build/libbls_oracle.so (the C restatement in bls_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package `pairing_b200`.

All arrays are numpy uint64 with the ABI layouts of include/pairing_b200.h:
  fq (n,6)  fq2 (n,12)  fq6 (n,36)  fq12 (n,72)  g1_affine (n,13)  g1 (n,18)
  g2_affine (n,25)  g2 (n,36)  fr_repr (n,4)  g2_prepared (n,2449)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbls_oracle.so")
_SRCS = [os.path.join(_HERE, f) for f in ("bls_oracle.c", "curve_impl.inc", "constants.h")]

W_FQ, W_FQ2, W_FQ6, W_FQ12 = 6, 12, 36, 72
W_G1A, W_G1, W_G2A, W_G2, W_FR, W_G2P = 13, 18, 25, 36, 4, 68 * 3 * 12 + 1

OP = dict(add=0, sub=1, mul=2, sqr=3, neg=4, dbl=5, inv=6, from_repr=7, into_repr=8, mul_nonres=9,
          frob1=10, frob2=11, frob3=12, conj=13, mul_by_014=14, mul_by_01=15, mul_by_1=16)
PT = dict(double=0, add=1, add_mixed=2, negate=3, mul=4, wnaf=5, sub=6)


def build(force=False):
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRCS)
    if stale:
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_sizeof.restype = ctypes.c_size_t
        assert [_lib.oracle_sizeof(i) for i in range(10)] == [8 * w for w in (
            W_FQ, W_FQ2, W_FQ6, W_FQ12, W_G1A, W_G1, W_G2A, W_G2, W_FR, W_G2P)]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _arr(a, w):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == w, (a.shape, w)
    return a


def default_threads():
    return len(os.sched_getaffinity(0))


def _field_op(fn, w, op, a, b=None):
    a = _arr(a, w)
    n = a.shape[0]
    if b is not None:
        b = _arr(b, w)
        assert b.shape[0] == n
    out = np.zeros_like(a)
    ok = np.zeros(n, dtype=np.uint8)
    rc = fn(OP[op], _p(a), _p(b), _p(out), _p(ok), ctypes.c_size_t(n))
    assert rc == 0, "unknown op %s" % op
    return out, ok


def fq_op(op, a, b=None): return _field_op(lib().oracle_fq_op, W_FQ, op, a, b)
def fq2_op(op, a, b=None): return _field_op(lib().oracle_fq2_op, W_FQ2, op, a, b)
def fq6_op(op, a, b=None): return _field_op(lib().oracle_fq6_op, W_FQ6, op, a, b)
def fq12_op(op, a, b=None): return _field_op(lib().oracle_fq12_op, W_FQ12, op, a, b)


def fq12_pow_u64(a, e):
    a = _arr(a, W_FQ12)
    out = np.zeros_like(a)
    lib().oracle_fq12_pow_u64(_p(a), ctypes.c_uint64(e), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def g2_prepare(q, threads=1):
    q = _arr(q, W_G2A)
    out = np.zeros((q.shape[0], W_G2P), dtype=np.uint64)
    lib().oracle_g2_prepare(_p(q), _p(out), ctypes.c_size_t(q.shape[0]), threads)
    return out


def miller_loop(p, q, threads=1):
    p, q = _arr(p, W_G1A), _arr(q, W_G2A)
    out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
    lib().oracle_miller_loop(_p(p), _p(q), _p(out), ctypes.c_size_t(p.shape[0]), threads)
    return out


def miller_loop_prepared(p, qp, threads=1):
    p, qp = _arr(p, W_G1A), _arr(qp, W_G2P)
    out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
    lib().oracle_miller_loop_prepared(_p(p), _p(qp), _p(out), ctypes.c_size_t(p.shape[0]), threads)
    return out


def multi_miller_loop(p, q):
    """Engine::miller_loop(&[(p0,q0),...]) with ONE shared accumulator."""
    p, q = _arr(p, W_G1A), _arr(q, W_G2A)
    out = np.zeros((1, W_FQ12), dtype=np.uint64)
    lib().oracle_multi_miller_loop(_p(p), _p(q), ctypes.c_size_t(p.shape[0]), _p(out))
    return out


def multi_miller_product(p, q, threads=1):
    p, q = _arr(p, W_G1A), _arr(q, W_G2A)
    out = np.zeros((1, W_FQ12), dtype=np.uint64)
    lib().oracle_multi_miller_product(_p(p), _p(q), ctypes.c_size_t(p.shape[0]), _p(out), threads)
    return out


def final_exponentiation(f, threads=1):
    f = _arr(f, W_FQ12)
    out = np.zeros_like(f)
    ok = np.zeros(f.shape[0], dtype=np.uint8)
    lib().oracle_final_exponentiation(_p(f), _p(out), _p(ok), ctypes.c_size_t(f.shape[0]), threads)
    return out, ok


def pairing(p, q, threads=1):
    p, q = _arr(p, W_G1A), _arr(q, W_G2A)
    out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
    lib().oracle_pairing(_p(p), _p(q), _p(out), ctypes.c_size_t(p.shape[0]), threads)
    return out


def _pt_op(fn, w, wa, op, a, b=None, k=None, window=0, threads=1):
    a = _arr(a, w)
    n = a.shape[0]
    if b is not None:
        b = _arr(b, wa if op == "add_mixed" else w)
    if k is not None:
        k = _arr(k, W_FR)
    out = np.zeros_like(a)
    fn(PT[op], _p(a), _p(b), _p(k), _p(out), ctypes.c_size_t(n), window, threads)
    return out


def g1_op(op, a, b=None, k=None, window=0, threads=1):
    return _pt_op(lib().oracle_g1_op, W_G1, W_G1A, op, a, b, k, window, threads)


def g2_op(op, a, b=None, k=None, window=0, threads=1):
    return _pt_op(lib().oracle_g2_op, W_G2, W_G2A, op, a, b, k, window, threads)


def g1_batch_normalization(v):
    v = _arr(v, W_G1).copy()
    lib().oracle_g1_batch_normalization(_p(v), ctypes.c_size_t(v.shape[0]))
    return v


def g2_batch_normalization(v):
    v = _arr(v, W_G2).copy()
    lib().oracle_g2_batch_normalization(_p(v), ctypes.c_size_t(v.shape[0]))
    return v


def _conv(fn, a, win, wout):
    a = _arr(a, win)
    out = np.zeros((a.shape[0], wout), dtype=np.uint64)
    fn(_p(a), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def g1_into_affine(v): return _conv(lib().oracle_g1_into_affine, v, W_G1, W_G1A)
def g2_into_affine(v): return _conv(lib().oracle_g2_into_affine, v, W_G2, W_G2A)
def g1_from_affine(v): return _conv(lib().oracle_g1_from_affine, v, W_G1A, W_G1)
def g2_from_affine(v): return _conv(lib().oracle_g2_from_affine, v, W_G2A, W_G2)


def wnaf_form(k, window):
    k = _arr(np.asarray(k, dtype=np.uint64).reshape(1, 4), W_FR)
    d = np.zeros(260, dtype=np.int64)
    n = lib().oracle_wnaf_form(_p(k), window, _p(d))
    return d[:n].tolist()


def generators():
    g1 = np.zeros((1, W_G1A), dtype=np.uint64)
    g2 = np.zeros((1, W_G2A), dtype=np.uint64)
    lib().oracle_generators(_p(g1), _p(g2))
    return g1, g2

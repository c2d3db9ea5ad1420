"""ctypes binding of libpairing_b200.so (the C ABI of include/pairing_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or no device is present every
entry point raises.  Nothing here imports or calls the oracle.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAIRING_B200_LIB lets tuning experiments point at an alternative build of the same library
LIB_PATH = os.environ.get("PAIRING_B200_LIB") or os.path.join(_HERE, "lib", "libpairing_b200.so")

# u64 words per ABI struct
W_FQ, W_FQ2, W_FQ6, W_FQ12 = 6, 12, 36, 72
W_G1A, W_G1, W_G2A, W_G2, W_FR, W_G2P = 13, 18, 25, 36, 4, 68 * 3 * 12 + 1

OPS = dict(add=0, sub=1, mul=2, sqr=3, neg=4, dbl=5, inv=6, from_repr=7, into_repr=8, mul_nonres=9,
           frob1=10, frob2=11, frob3=12, conj=13, mul_by_014=14, mul_by_01=15, mul_by_1=16, sqrt=17,
           mul_by_line_pair=18, cyclotomic_sqr=19)
PT_OPS = dict(double=0, add=1, add_mixed=2, negate=3, sub=6)

# every symbol include/pairing_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "bls_ctx_create", "bls_ctx_destroy", "bls_strerror", "bls_ctx_last_error", "bls_ctx_device",
    "bls_ctx_sm_count", "bls_ctx_launch_count", "bls_ctx_set_latency_path_limits", "bls_ctx_trim",
    "bls_g2_prepare_batch", "bls_miller_loop_batch", "bls_miller_loop_prepared_batch",
    "bls_multi_miller_loop", "bls_multi_miller_loop_prepared", "bls_final_exponentiation_batch",
    "bls_pairing_batch", "bls_fq12_product",
    "bls_g1_wnaf_mul_batch", "bls_g2_wnaf_mul_batch", "bls_g1_wnaf_mul_window_batch",
    "bls_g2_wnaf_mul_window_batch", "bls_g1_mul_batch", "bls_g2_mul_batch",
    "bls_g1_batch_normalization", "bls_g2_batch_normalization", "bls_g1_into_affine_batch",
    "bls_g2_into_affine_batch", "bls_g1_op_batch", "bls_g2_op_batch", "bls_field_op_batch",
    "bls_g2_prepare_dev", "bls_miller_loop_dev", "bls_miller_loop_prepared_dev",
    "bls_final_exponentiation_dev", "bls_pairing_dev", "bls_multi_miller_scratch_bytes",
    "bls_multi_miller_loop_dev", "bls_fq12_product_scratch_bytes", "bls_fq12_product_dev",
    "bls_g1_wnaf_mul_dev", "bls_g2_wnaf_mul_dev", "bls_batch_normalization_scratch_bytes",
    "bls_g1_batch_normalization_dev", "bls_g2_batch_normalization_dev", "bls_imad_peak",
    "bls_g1_wnaf_fixed_base_batch", "bls_g2_wnaf_fixed_base_batch", "bls_g1_wnaf_table", "bls_g2_wnaf_table",
    "bls_fq12_pow_batch", "bls_fq12_pow_dev", "bls_fr_op_batch",
    "bls_miller_loop_shared_q_batch", "bls_pairing_shared_q_batch", "bls_miller_loop_shared_q_dev",
    "bls_g1_affine_mul_batch", "bls_g2_affine_mul_batch",
    "bls_g1_point_from_x_batch", "bls_g2_point_from_x_batch", "bls_g1_scale_by_cofactor_batch", "bls_g2_scale_by_cofactor_batch",
    "bls_g1_decode_batch", "bls_g2_decode_batch", "bls_g1_encode_batch", "bls_g2_encode_batch",
    "bls_g1_wnaf_table_dev", "bls_g2_wnaf_table_dev", "bls_g1_wnaf_fixed_base_dev", "bls_g2_wnaf_fixed_base_dev",
    "bls_pairing_product", "bls_pairing_product_dev", "bls_fq12_product_tail_dev",
    "bls_pair_field_op_batch", "bls_pair_field_op_dev", "bls_pairing_projective_batch", "bls_pairing_projective_dev",
    "bls_mgpu_create", "bls_mgpu_destroy", "bls_mgpu_device_count", "bls_mgpu_ctx", "bls_mgpu_multi_miller_loop",
    "bls_mgpu_pairing_product", "bls_mgpu_pairing_batch", "bls_mgpu_g1_wnaf_mul_batch", "bls_mgpu_g2_wnaf_mul_batch",
    "bls_mgpu_last_phase_ms",
]


class BlsError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BlsError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C pairing_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.bls_ctx_create.restype = vp
    lib.bls_ctx_create.argtypes = [ci, ctypes.POINTER(ci)]
    lib.bls_ctx_destroy.argtypes = [vp]
    lib.bls_ctx_destroy.restype = None
    lib.bls_strerror.restype = ctypes.c_char_p
    lib.bls_strerror.argtypes = [ci]
    lib.bls_ctx_last_error.restype = ctypes.c_char_p
    lib.bls_ctx_last_error.argtypes = [vp]
    lib.bls_ctx_device.argtypes = [vp]
    lib.bls_ctx_sm_count.argtypes = [vp]
    lib.bls_ctx_launch_count.restype = ctypes.c_uint64
    lib.bls_ctx_launch_count.argtypes = [vp]
    lib.bls_ctx_set_latency_path_limits.argtypes = [vp, sz, sz]
    lib.bls_ctx_trim.argtypes = [vp, sz]
    for name in ("bls_multi_miller_scratch_bytes", "bls_fq12_product_scratch_bytes"):
        getattr(lib, name).restype = sz
        getattr(lib, name).argtypes = [vp, sz]
    lib.bls_batch_normalization_scratch_bytes.restype = sz
    lib.bls_batch_normalization_scratch_bytes.argtypes = [vp, ci, sz]
    sig = {
        "bls_g2_prepare_batch": [vp, vp, vp, sz],
        "bls_miller_loop_batch": [vp, vp, vp, vp, sz],
        "bls_miller_loop_prepared_batch": [vp, vp, vp, vp, sz],
        "bls_multi_miller_loop": [vp, vp, vp, sz, vp],
        "bls_multi_miller_loop_prepared": [vp, vp, vp, sz, vp],
        "bls_final_exponentiation_batch": [vp, vp, vp, vp, sz],
        "bls_pairing_batch": [vp, vp, vp, vp, sz],
        "bls_fq12_product": [vp, vp, sz, vp],
        "bls_g1_wnaf_mul_batch": [vp, vp, vp, vp, sz],
        "bls_g2_wnaf_mul_batch": [vp, vp, vp, vp, sz],
        "bls_g1_wnaf_mul_window_batch": [vp, vp, vp, vp, sz, ci],
        "bls_g2_wnaf_mul_window_batch": [vp, vp, vp, vp, sz, ci],
        "bls_g1_mul_batch": [vp, vp, vp, vp, sz],
        "bls_g2_mul_batch": [vp, vp, vp, vp, sz],
        "bls_g1_batch_normalization": [vp, vp, sz],
        "bls_g2_batch_normalization": [vp, vp, sz],
        "bls_g1_into_affine_batch": [vp, vp, vp, sz],
        "bls_g2_into_affine_batch": [vp, vp, vp, sz],
        "bls_g1_op_batch": [vp, ci, vp, vp, vp, sz],
        "bls_g2_op_batch": [vp, ci, vp, vp, vp, sz],
        "bls_field_op_batch": [vp, ci, ci, vp, vp, vp, vp, sz],
        "bls_g1_wnaf_fixed_base_batch": [vp, vp, ci, vp, vp, sz],
        "bls_g2_wnaf_fixed_base_batch": [vp, vp, ci, vp, vp, sz],
        "bls_g1_wnaf_table": [vp, vp, ci, vp],
        "bls_g2_wnaf_table": [vp, vp, ci, vp],
        "bls_g1_wnaf_table_dev": [vp, vp, ci, vp, vp],
        "bls_g2_wnaf_table_dev": [vp, vp, ci, vp, vp],
        "bls_g1_wnaf_fixed_base_dev": [vp, vp, ci, vp, vp, sz, vp],
        "bls_g2_wnaf_fixed_base_dev": [vp, vp, ci, vp, vp, sz, vp],
        "bls_g1_decode_batch": [vp, vp, ci, ci, vp, vp, sz],
        "bls_g2_decode_batch": [vp, vp, ci, ci, vp, vp, sz],
        "bls_g1_encode_batch": [vp, vp, ci, vp, sz],
        "bls_g2_encode_batch": [vp, vp, ci, vp, sz],
        "bls_fr_op_batch": [vp, ci, vp, vp, vp, vp, sz],
        "bls_miller_loop_shared_q_batch": [vp, vp, vp, vp, sz],
        "bls_pairing_shared_q_batch": [vp, vp, vp, vp, sz],
        "bls_miller_loop_shared_q_dev": [vp, vp, vp, vp, sz, ci, vp],
        "bls_g1_affine_mul_batch": [vp, vp, vp, vp, sz],
        "bls_g2_affine_mul_batch": [vp, vp, vp, vp, sz],
        "bls_g1_point_from_x_batch": [vp, vp, vp, vp, vp, sz],
        "bls_g2_point_from_x_batch": [vp, vp, vp, vp, vp, sz],
        "bls_g1_scale_by_cofactor_batch": [vp, vp, vp, sz],
        "bls_g2_scale_by_cofactor_batch": [vp, vp, vp, sz],
        "bls_fq12_pow_batch": [vp, vp, vp, vp, sz],
        "bls_fq12_pow_dev": [vp, vp, vp, vp, sz, vp],
        "bls_g2_prepare_dev": [vp, vp, vp, sz, vp],
        "bls_miller_loop_dev": [vp, vp, vp, vp, sz, vp],
        "bls_miller_loop_prepared_dev": [vp, vp, vp, vp, sz, vp],
        "bls_final_exponentiation_dev": [vp, vp, vp, vp, sz, vp],
        "bls_pairing_dev": [vp, vp, vp, vp, sz, vp],
        "bls_multi_miller_loop_dev": [vp, vp, vp, sz, vp, vp, vp],
        "bls_fq12_product_dev": [vp, vp, sz, vp, vp, vp],
        "bls_g1_wnaf_mul_dev": [vp, vp, vp, vp, sz, ci, vp],
        "bls_g2_wnaf_mul_dev": [vp, vp, vp, vp, sz, ci, vp],
        "bls_g1_batch_normalization_dev": [vp, vp, sz, vp, vp],
        "bls_g2_batch_normalization_dev": [vp, vp, sz, vp, vp],
        "bls_imad_peak": [vp, ci, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)],
        "bls_pairing_product": [vp, vp, vp, sz, vp, vp],
        "bls_pair_field_op_batch": [vp, ci, ci, vp, vp, vp, vp, sz],
        "bls_pairing_projective_batch": [vp, vp, vp, vp, sz],
        "bls_pairing_projective_dev": [vp, vp, vp, vp, sz, vp],
        "bls_pair_field_op_dev": [vp, ci, ci, vp, vp, vp, vp, sz, vp],
        "bls_pairing_product_dev": [vp, vp, vp, sz, vp, vp, vp, vp],
        "bls_fq12_product_tail_dev": [vp, vp, sz, vp, ci, vp, vp],
        "bls_mgpu_multi_miller_loop": [vp, vp, vp, sz, vp],
        "bls_mgpu_pairing_product": [vp, vp, vp, sz, vp, vp],
        "bls_mgpu_pairing_batch": [vp, vp, vp, vp, sz],
        "bls_mgpu_g1_wnaf_mul_batch": [vp, vp, vp, vp, sz],
        "bls_mgpu_g2_wnaf_mul_batch": [vp, vp, vp, vp, sz],
        "bls_mgpu_last_phase_ms": [vp, ctypes.POINTER(ctypes.c_double)],
    }
    lib.bls_mgpu_create.restype = vp
    lib.bls_mgpu_create.argtypes = [ctypes.POINTER(ci), ci, ctypes.POINTER(ci)]
    lib.bls_mgpu_destroy.restype = None
    lib.bls_mgpu_destroy.argtypes = [vp]
    lib.bls_mgpu_device_count.argtypes = [vp]
    lib.bls_mgpu_ctx.restype = vp
    lib.bls_mgpu_ctx.argtypes = [vp, ci]
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ci
    _lib = lib
    return lib


def _arr(a, w, name="array"):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != w:
        raise ValueError("%s must have shape (n, %d) uint64, got %r" % (name, w, a.shape))
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """One CUDA device + stream (bls_ctx).  Host-array methods take and return numpy uint64 arrays
    in the ABI layouts; `*_dev` methods take raw device pointers (ints) and a CUDA stream handle."""

    def __init__(self, device=0):
        self._lib = load()
        err = ctypes.c_int(0)
        self._ctx = self._lib.bls_ctx_create(int(device), ctypes.byref(err))
        if not self._ctx:
            raise BlsError("bls_ctx_create(device=%d) failed: %s" % (device, self._lib.bls_strerror(err.value).decode()))
        self.device = int(device)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.bls_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise BlsError("%s [%s]" % (self._lib.bls_strerror(rc).decode(),
                                        self._lib.bls_ctx_last_error(self._ctx).decode()))

    @property
    def sm_count(self):
        return self._lib.bls_ctx_sm_count(self._ctx)

    @property
    def launch_count(self):
        return int(self._lib.bls_ctx_launch_count(self._ctx))

    def trim(self, keep_bytes=0):
        """release the staging memory the context caches between host-buffer calls"""
        self._check(self._lib.bls_ctx_trim(self._ctx, int(keep_bytes)))

    def set_latency_path_limits(self, max_pairings, max_final_exps):
        """Batch sizes up to which pairing / final_exponentiation run one WARP per element (0: lane-pair kernels always)."""
        self._check(self._lib.bls_ctx_set_latency_path_limits(self._ctx, int(max_pairings), int(max_final_exps)))

    # ------------------------------------------------------------------ engine, host arrays
    def g2_prepare(self, q):
        q = _arr(q, W_G2A, "q")
        out = np.zeros((q.shape[0], W_G2P), dtype=np.uint64)
        self._check(self._lib.bls_g2_prepare_batch(self._ctx, _p(q), _p(out), q.shape[0]))
        return out

    def _pq(self, fn, p, q, wq):
        p, q = _arr(p, W_G1A, "p"), _arr(q, wq, "q")
        if p.shape[0] != q.shape[0]:
            raise ValueError("p and q must have the same length")
        out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
        self._check(fn(self._ctx, _p(p), _p(q), _p(out), p.shape[0]))
        return out

    def miller_loop(self, p, q):
        return self._pq(self._lib.bls_miller_loop_batch, p, q, W_G2A)

    def miller_loop_prepared(self, p, qp):
        return self._pq(self._lib.bls_miller_loop_prepared_batch, p, qp, W_G2P)

    def pairing(self, p, q):
        return self._pq(self._lib.bls_pairing_batch, p, q, W_G2A)

    def pairing_projective(self, p, q):
        """Engine::pairing on Jacobian inputs (the two into_affine conversions fused in front): (n,18),(n,36) -> (n,72)."""
        p, q = _arr(p, W_G1, "p"), _arr(q, W_G2, "q")
        if p.shape[0] != q.shape[0]:
            raise ValueError("p and q must have the same length")
        out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
        self._check(self._lib.bls_pairing_projective_batch(self._ctx, _p(p), _p(q), _p(out), p.shape[0]))
        return out

    def _shared_q(self, fn, p, q1):
        p, q1 = _arr(p, W_G1A, "p"), _arr(q1, W_G2P, "q1")
        if q1.shape[0] != 1:
            raise ValueError("q1 must be ONE prepared G2 point")
        out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
        self._check(fn(self._ctx, _p(p), _p(q1), _p(out), p.shape[0]))
        return out

    def miller_loop_shared_q(self, p, q1):
        """n Miller loops e(P_i, Q) against one prepared Q."""
        return self._shared_q(self._lib.bls_miller_loop_shared_q_batch, p, q1)

    def pairing_shared_q(self, p, q1):
        return self._shared_q(self._lib.bls_pairing_shared_q_batch, p, q1)

    def _multi(self, fn, p, q, wq):
        p, q = _arr(p, W_G1A, "p"), _arr(q, wq, "q")
        if p.shape[0] != q.shape[0]:
            raise ValueError("p and q must have the same length")
        out = np.zeros((1, W_FQ12), dtype=np.uint64)
        self._check(fn(self._ctx, _p(p), _p(q), p.shape[0], _p(out)))
        return out

    def multi_miller_loop(self, p, q):
        return self._multi(self._lib.bls_multi_miller_loop, p, q, W_G2A)

    def multi_miller_loop_prepared(self, p, qp):
        return self._multi(self._lib.bls_multi_miller_loop_prepared, p, qp, W_G2P)

    def pairing_product(self, p, q):
        """final_exponentiation(miller_loop(all pairs)) in one call -> ((1, 72) GT element, is_some)."""
        p, q = _arr(p, W_G1A, "p"), _arr(q, W_G2A, "q")
        if p.shape[0] != q.shape[0]:
            raise ValueError("p and q must have the same length")
        out = np.zeros((1, W_FQ12), dtype=np.uint64)
        ok = np.zeros(1, dtype=np.uint8)
        self._check(self._lib.bls_pairing_product(self._ctx, _p(p), _p(q), p.shape[0], _p(out), _p(ok)))
        return out, bool(ok[0])

    def final_exponentiation(self, f):
        f = _arr(f, W_FQ12, "f")
        out = np.zeros_like(f)
        ok = np.zeros(f.shape[0], dtype=np.uint8)
        self._check(self._lib.bls_final_exponentiation_batch(self._ctx, _p(f), _p(out), _p(ok), f.shape[0]))
        return out, ok

    def fq12_pow(self, a, k):
        """Fq12::pow(FrRepr) per element (lib.rs:306-324)."""
        a, k = _arr(a, W_FQ12, "a"), _arr(k, W_FR, "k")
        if a.shape[0] != k.shape[0]:
            raise ValueError("a and k must have the same length")
        out = np.zeros_like(a)
        self._check(self._lib.bls_fq12_pow_batch(self._ctx, _p(a), _p(k), _p(out), a.shape[0]))
        return out

    def fq12_product(self, f):
        f = _arr(f, W_FQ12, "f")
        out = np.zeros((1, W_FQ12), dtype=np.uint64)
        self._check(self._lib.bls_fq12_product(self._ctx, _p(f), f.shape[0], _p(out)))
        return out

    # ------------------------------------------------------------------ curves, host arrays
    def _wnaf(self, g2, bases, k, window, mode):
        w = W_G2 if g2 else W_G1
        bases, k = _arr(bases, w, "bases"), _arr(k, W_FR, "k")
        if bases.shape[0] != k.shape[0]:
            raise ValueError("bases and k must have the same length")
        out = np.zeros_like(bases)
        n = bases.shape[0]
        L = self._lib
        if mode == "mul":
            fn = L.bls_g2_mul_batch if g2 else L.bls_g1_mul_batch
            self._check(fn(self._ctx, _p(bases), _p(k), _p(out), n))
        elif window:
            fn = L.bls_g2_wnaf_mul_window_batch if g2 else L.bls_g1_wnaf_mul_window_batch
            self._check(fn(self._ctx, _p(bases), _p(k), _p(out), n, int(window)))
        else:
            fn = L.bls_g2_wnaf_mul_batch if g2 else L.bls_g1_wnaf_mul_batch
            self._check(fn(self._ctx, _p(bases), _p(k), _p(out), n))
        return out

    def decode(self, g2, data, compressed, checked=True):
        """EncodedPoint::into_affine(_unchecked) for n encodings back to back -> (affine rows, status bytes)."""
        size = (96 if g2 else 48) * (1 if compressed else 2)
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        if buf.size % size:
            raise ValueError("encoded data is not a multiple of %d bytes" % size)
        n = buf.size // size
        out = np.zeros((n, W_G2A if g2 else W_G1A), dtype=np.uint64)
        status = np.zeros(n, dtype=np.uint8)
        buf = np.ascontiguousarray(buf)
        fn = self._lib.bls_g2_decode_batch if g2 else self._lib.bls_g1_decode_batch
        self._check(fn(self._ctx, _p(buf), int(bool(compressed)), int(bool(checked)), _p(out), _p(status), n))
        return out, status

    def encode(self, g2, affine, compressed):
        """EncodedPoint::from_affine for n affine points -> bytes."""
        affine = _arr(affine, W_G2A if g2 else W_G1A, "affine")
        size = (96 if g2 else 48) * (1 if compressed else 2)
        out = np.zeros(affine.shape[0] * size, dtype=np.uint8)
        fn = self._lib.bls_g2_encode_batch if g2 else self._lib.bls_g1_encode_batch
        self._check(fn(self._ctx, _p(affine), int(bool(compressed)), _p(out), affine.shape[0]))
        return out.tobytes()

    def _wnaf_fixed(self, g2, base, window, k):
        """Wnaf::base(g, n).scalar(k_i): one shared table (window 2..16), every scalar against it."""
        w = W_G2 if g2 else W_G1
        base, k = _arr(base, w, "base"), _arr(k, W_FR, "k")
        if base.shape[0] != 1:
            raise ValueError("fixed-base mode takes exactly one base point")
        out = np.zeros((k.shape[0], w), dtype=np.uint64)
        fn = self._lib.bls_g2_wnaf_fixed_base_batch if g2 else self._lib.bls_g1_wnaf_fixed_base_batch
        self._check(fn(self._ctx, _p(base), int(window), _p(k), _p(out), k.shape[0]))
        return out

    def g1_wnaf_fixed_base(self, base, window, k): return self._wnaf_fixed(False, base, window, k)
    def g2_wnaf_fixed_base(self, base, window, k): return self._wnaf_fixed(True, base, window, k)

    def _wnaf_table(self, g2, base, window):
        w = W_G2 if g2 else W_G1
        base = _arr(base, w, "base")
        if not 2 <= int(window) <= 16:
            raise ValueError("window must be in 2..16")
        table = np.zeros((1 << (int(window) - 1), w), dtype=np.uint64)
        fn = self._lib.bls_g2_wnaf_table if g2 else self._lib.bls_g1_wnaf_table
        self._check(fn(self._ctx, _p(base), int(window), _p(table)))
        return table

    def g1_wnaf_table(self, base, window): return self._wnaf_table(False, base, window)
    def g2_wnaf_table(self, base, window): return self._wnaf_table(True, base, window)

    def g1_wnaf_mul(self, bases, k, window=0): return self._wnaf(False, bases, k, window, "wnaf")
    def g2_wnaf_mul(self, bases, k, window=0): return self._wnaf(True, bases, k, window, "wnaf")
    def g1_mul(self, bases, k): return self._wnaf(False, bases, k, 0, "mul")
    def g2_mul(self, bases, k): return self._wnaf(True, bases, k, 0, "mul")

    def g1_batch_normalization(self, v):
        v = _arr(v, W_G1, "v").copy()
        self._check(self._lib.bls_g1_batch_normalization(self._ctx, _p(v), v.shape[0]))
        return v

    def g2_batch_normalization(self, v):
        v = _arr(v, W_G2, "v").copy()
        self._check(self._lib.bls_g2_batch_normalization(self._ctx, _p(v), v.shape[0]))
        return v

    def g1_into_affine(self, v):
        v = _arr(v, W_G1, "v")
        out = np.zeros((v.shape[0], W_G1A), dtype=np.uint64)
        self._check(self._lib.bls_g1_into_affine_batch(self._ctx, _p(v), _p(out), v.shape[0]))
        return out

    def g2_into_affine(self, v):
        v = _arr(v, W_G2, "v")
        out = np.zeros((v.shape[0], W_G2A), dtype=np.uint64)
        self._check(self._lib.bls_g2_into_affine_batch(self._ctx, _p(v), _p(out), v.shape[0]))
        return out

    def _pt_op(self, g2, op, a, b):
        w, wa = (W_G2, W_G2A) if g2 else (W_G1, W_G1A)
        a = _arr(a, w, "a")
        if b is not None:
            b = _arr(b, wa if op == "add_mixed" else w, "b")
        out = np.zeros_like(a)
        fn = self._lib.bls_g2_op_batch if g2 else self._lib.bls_g1_op_batch
        self._check(fn(self._ctx, PT_OPS[op], _p(a), _p(b), _p(out), a.shape[0]))
        return out

    def g1_op(self, op, a, b=None): return self._pt_op(False, op, a, b)
    def g2_op(self, op, a, b=None): return self._pt_op(True, op, a, b)

    # ------------------------------------------------------------------ field tower, host arrays
    def field_op(self, degree, op, a, b=None):
        a = _arr(a, 6 * degree, "a")
        if b is not None:
            b = _arr(b, 6 * degree, "b")
        out = np.zeros_like(a)
        ok = np.zeros(a.shape[0], dtype=np.uint8)
        self._check(self._lib.bls_field_op_batch(self._ctx, degree, OPS[op], _p(a), _p(b), _p(out), _p(ok), a.shape[0]))
        return out, ok

    def pair_field_op(self, degree, op, a, b=None):
        """the same operation on the lane-pair tower of the pairing kernels (degree 2, 6, 12)"""
        a = _arr(a, 6 * degree, "a")
        if b is not None:
            b = _arr(b, 6 * degree, "b")
            if b.shape[0] != a.shape[0]:
                raise ValueError("a and b must have the same length")
        out = np.zeros_like(a)
        ok = np.zeros(a.shape[0], dtype=np.uint8)
        self._check(self._lib.bls_pair_field_op_batch(self._ctx, degree, OPS[op], _p(a), _p(b), _p(out), _p(ok), a.shape[0]))
        return out, ok

    # ------------------------------------------------------------------ device pointers
    def point_from_x(self, g2, x, greatest):
        """$affine::get_point_from_x for n x-coordinates (Montgomery limbs) -> (affine rows, is_some)."""
        x = _arr(x, W_FQ2 if g2 else W_FQ, "x")
        gr = np.ascontiguousarray(greatest, dtype=np.uint8).reshape(-1)
        if gr.shape[0] != x.shape[0]:
            raise ValueError("x and greatest must have the same length")
        out = np.zeros((x.shape[0], W_G2A if g2 else W_G1A), dtype=np.uint64)
        ok = np.zeros(x.shape[0], dtype=np.uint8)
        fn = self._lib.bls_g2_point_from_x_batch if g2 else self._lib.bls_g1_point_from_x_batch
        self._check(fn(self._ctx, _p(x), _p(gr), _p(out), _p(ok), x.shape[0]))
        return out, ok

    def affine_mul(self, g2, affine, k):
        """CurveAffine::mul for n (affine point, FrRepr) pairs -> Jacobian rows."""
        affine, k = _arr(affine, W_G2A if g2 else W_G1A, "affine"), _arr(k, W_FR, "k")
        if affine.shape[0] != k.shape[0]:
            raise ValueError("affine and k must have the same length")
        out = np.zeros((affine.shape[0], W_G2 if g2 else W_G1), dtype=np.uint64)
        fn = self._lib.bls_g2_affine_mul_batch if g2 else self._lib.bls_g1_affine_mul_batch
        self._check(fn(self._ctx, _p(affine), _p(k), _p(out), affine.shape[0]))
        return out

    def scale_by_cofactor(self, g2, affine):
        affine = _arr(affine, W_G2A if g2 else W_G1A, "affine")
        out = np.zeros((affine.shape[0], W_G2 if g2 else W_G1), dtype=np.uint64)
        fn = self._lib.bls_g2_scale_by_cofactor_batch if g2 else self._lib.bls_g1_scale_by_cofactor_batch
        self._check(fn(self._ctx, _p(affine), _p(out), affine.shape[0]))
        return out

    def fr_op(self, op, a, b=None):
        """Fr field operation (Montgomery-form (n,4) arrays; from_repr/into_repr convert) -> (values, ok)."""
        a = _arr(a, W_FR, "a")
        if b is not None:
            b = _arr(b, W_FR, "b")
        out = np.zeros_like(a)
        ok = np.zeros(a.shape[0], dtype=np.uint8)
        self._check(self._lib.bls_fr_op_batch(self._ctx, OPS[op], _p(a), _p(b) if b is not None else None, _p(out), _p(ok), a.shape[0]))
        return out, ok

    def call_dev(self, name, *args):
        """Raw access to a `*_dev` entry point: args are ints (device pointers / sizes / stream)."""
        self._check(getattr(self._lib, name)(self._ctx, *args))

    def multi_miller_scratch_bytes(self, n):
        return int(self._lib.bls_multi_miller_scratch_bytes(self._ctx, n))

    def fq12_product_scratch_bytes(self, n):
        return int(self._lib.bls_fq12_product_scratch_bytes(self._ctx, n))

    def batch_normalization_scratch_bytes(self, degree, n):
        return int(self._lib.bls_batch_normalization_scratch_bytes(self._ctx, degree, n))

    def imad_peak(self, variant=0, iters=2000):
        macs, ms = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._lib.bls_imad_peak(self._ctx, variant, iters, ctypes.byref(macs), ctypes.byref(ms)))
        return macs.value, ms.value


class MultiGpu:
    """Several devices of one node behind one call (bls_mgpu): an n-pair product or batch is split into contiguous
    shards, one host thread per device inside the library; the product's 576-byte partials meet on the first device."""

    def __init__(self, devices):
        self._lib = load()
        devices = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
        arr = (ctypes.c_int * len(devices))(*devices)
        err = ctypes.c_int(0)
        self._m = self._lib.bls_mgpu_create(arr, len(devices), ctypes.byref(err))
        if not self._m:
            raise BlsError("bls_mgpu_create(%r) failed: %s" % (devices, self._lib.bls_strerror(err.value).decode()))
        self.devices = devices

    def close(self):
        if getattr(self, "_m", None):
            self._lib.bls_mgpu_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            msgs = [self._lib.bls_ctx_last_error(self._lib.bls_mgpu_ctx(self._m, i)).decode() for i in range(len(self.devices))]
            raise BlsError("%s %r" % (self._lib.bls_strerror(rc).decode(), [m for m in msgs if m]))

    @property
    def device_count(self):
        return self._lib.bls_mgpu_device_count(self._m)

    def _pairs(self, p, q):
        p, q = _arr(p, W_G1A, "p"), _arr(q, W_G2A, "q")
        if p.shape[0] != q.shape[0]:
            raise ValueError("p and q must have the same length")
        return p, q

    def multi_miller_loop(self, p, q):
        p, q = self._pairs(p, q)
        out = np.zeros((1, W_FQ12), dtype=np.uint64)
        self._check(self._lib.bls_mgpu_multi_miller_loop(self._m, _p(p), _p(q), p.shape[0], _p(out)))
        return out

    def pairing_product(self, p, q):
        p, q = self._pairs(p, q)
        out = np.zeros((1, W_FQ12), dtype=np.uint64)
        ok = np.zeros(1, dtype=np.uint8)
        self._check(self._lib.bls_mgpu_pairing_product(self._m, _p(p), _p(q), p.shape[0], _p(out), _p(ok)))
        return out, bool(ok[0])

    def pairing(self, p, q):
        p, q = self._pairs(p, q)
        out = np.zeros((p.shape[0], W_FQ12), dtype=np.uint64)
        self._check(self._lib.bls_mgpu_pairing_batch(self._m, _p(p), _p(q), _p(out), p.shape[0]))
        return out

    def _wnaf(self, fn, w, bases, k):
        bases, k = _arr(bases, w, "bases"), _arr(k, W_FR, "k")
        if bases.shape[0] != k.shape[0]:
            raise ValueError("bases and k must have the same length")
        out = np.zeros_like(bases)
        self._check(fn(self._m, _p(bases), _p(k), _p(out), bases.shape[0]))
        return out

    def g1_wnaf_mul(self, bases, k):
        return self._wnaf(self._lib.bls_mgpu_g1_wnaf_mul_batch, W_G1, bases, k)

    def g2_wnaf_mul(self, bases, k):
        return self._wnaf(self._lib.bls_mgpu_g2_wnaf_mul_batch, W_G2, bases, k)

    def last_phase_ms(self):
        ms = (ctypes.c_double * 3)()
        self._check(self._lib.bls_mgpu_last_phase_ms(self._m, ms))
        return {"shards_ms": ms[0], "tail_ms": ms[1], "total_ms": ms[2]}

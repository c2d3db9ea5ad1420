"""pairing_b200 -- B200-native batched BLS12-381 pairing / wNAF engine.

Host-side mirror of the `pairing` crate's trait surface for the hot path (reference v0.14.2):

    Bls12.miller_loop / final_exponentiation / pairing      Engine            src/lib.rs:34-110
    G1 / G2  double, add_assign, add_assign_mixed, negate,  CurveProjective   src/lib.rs:114-181
             mul_assign, into_affine, batch_normalization,
             recommended_wnaf_for_scalar / _num_scalars
    G1Affine / G2Affine  prepare, pairing_with, into_projective  CurveAffine  src/lib.rs:185-234
    Wnaf  (scalar(k).base(g) and base(g, n).scalar(k))       Wnaf             src/wnaf.rs:75-179
    G1Compressed / G1Uncompressed / G2Compressed / G2Uncompressed  EncodedPoint  src/lib.rs:236-263

Everything is slice-shaped: the arguments are numpy uint64 arrays in the C-ABI layouts of
include/pairing_b200.h (one row per element), because the point of the GPU path is the batch.  All
arithmetic runs in the CUDA library (pairing_b200/lib/libpairing_b200.so); there is no CPU
fallback, and importing this package never touches the oracle.
"""
from ._native import (BlsError, Context, MultiGpu, LIB_PATH, SYMBOLS, W_FQ, W_FQ2, W_FQ6, W_FQ12, W_FR, W_G1,  # noqa: F401
                      W_G1A, W_G2, W_G2A, W_G2P, load)
from .engine import (Bls12, G1, G1Affine, G1Compressed, G1Uncompressed, G2, G2Affine, G2Compressed, G2Prepared,  # noqa: F401
                     G2Uncompressed, Wnaf, default_context)

__all__ = ["Bls12", "G1", "G2", "G1Affine", "G2Affine", "G2Prepared", "Wnaf", "G1Compressed", "G1Uncompressed",
           "G2Compressed", "G2Uncompressed", "Context", "MultiGpu", "BlsError", "default_context", "load"]

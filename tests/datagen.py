"""Deterministic synthetic inputs: a SplitMix64 stream, uniform field
elements / scalars by rejection sampling (mirrors Fq::rand fq.rs:723-736 and Fr::rand
fr.rs:255-268), subgroup points as scalar multiples of the generators.

Test-side helper: point generation uses the oracle (tests may), bench.py generates its points with
the product's own kernels instead.
"""
import numpy as np

import bls_model as m
import oracle_lib as o

SEED0 = 0x5DBE62598D313D76
MASK64 = (1 << 64) - 1


def splitmix64(seed, n):
    """n u64 words of the SplitMix64 stream started at `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _rand_below(n, seed, limbs, top_mask, bound):
    """(n, limbs) u64: uniform integers < bound by rejection sampling."""
    out = np.zeros((n, limbs), dtype=np.uint64)
    filled = 0
    rnd = 0
    bl = np.array(m.limbs64(bound, limbs), dtype=np.uint64)
    while filled < n:
        need = (n - filled) * 2 + 16
        w = splitmix64(seed + 0x1000 * rnd, need * limbs).reshape(need, limbs)
        rnd += 1
        w[:, limbs - 1] &= np.uint64(top_mask)
        # lexicographic compare from the top limb
        lt = np.zeros(need, dtype=bool)
        eq = np.ones(need, dtype=bool)
        for j in range(limbs - 1, -1, -1):
            lt |= eq & (w[:, j] < bl[j])
            eq &= w[:, j] == bl[j]
        good = w[lt]
        take = min(len(good), n - filled)
        out[filled:filled + take] = good[:take]
        filled += take
    return out


def rand_fq(n, seed):
    """(n,6) canonical residues < q (read as Montgomery-form elements they are uniform too)."""
    return _rand_below(n, seed, 6, MASK64 >> 3, m.Q)


def rand_field(n, degree, seed):
    return rand_fq(n * degree, seed).reshape(n, 6 * degree)


def rand_scalars(n, seed, edge_cases=True):
    """(n,4) canonical scalars < r, with the window-threshold edge cases of SURVEY 8d first."""
    k = _rand_below(n, seed ^ 0xABCDEF, 4, MASK64 >> 1, m.R_ORDER)
    if edge_cases:
        edges = [0, 1, m.R_ORDER - 1, 1 << 33, 1 << 129, (1 << 130) - 1, (1 << 36) - 5, (1 << 102) - 3,
                 (1 << 32) + 1, (1 << 102) + 7, 2, 3]
        for i, e in enumerate(edges[:n]):
            k[i] = np.array(m.limbs64(e, 4), dtype=np.uint64)
    return k


def g1_points(n, seed, infinity_at=()):
    """Non-normalised Jacobian G1 points [a_i] g1 (as G::rand yields, Z != 1) -> (n,18)."""
    g1, _ = o.generators()
    base = np.repeat(o.g1_from_affine(g1), n, 0)
    a = rand_scalars(n, seed ^ 0x1111, edge_cases=False)
    a[:, 0] |= np.uint64(2)          # never 0 or 1: keeps Z != 1 and the point finite
    pts = o.g1_op("mul", base, k=a, threads=o.default_threads())
    for i in infinity_at:
        pts[i] = 0
        pts[i, 6:12] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)   # (0, 1, 0)
    return pts


def g2_points(n, seed, infinity_at=()):
    _, g2 = o.generators()
    base = np.repeat(o.g2_from_affine(g2), n, 0)
    b = rand_scalars(n, seed ^ 0x2222, edge_cases=False)
    b[:, 0] |= np.uint64(2)
    pts = o.g2_op("mul", base, k=b, threads=o.default_threads())
    for i in infinity_at:
        pts[i] = 0
        pts[i, 12:18] = np.array(m.limbs64(m.MONT_R), dtype=np.uint64)
    return pts


def g1_affine_points(n, seed, infinity_at=()):
    return o.g1_into_affine(g1_points(n, seed, infinity_at))


def g2_affine_points(n, seed, infinity_at=()):
    return o.g2_into_affine(g2_points(n, seed, infinity_at))

"""CPU tests of the host-side logic: window heuristics, sharding, the gloo world_size-2 exchange of
the sharded multi-Miller product (local products computed by the oracle here -- there is no GPU)."""
import os
import socket

import numpy as np
import pytest

import bls_model as m
import datagen as dg
import oracle_lib as o
from pairing_b200 import engine
from pairing_b200 import dist as pdist


def test_num_bits_and_windows_match_reference_rules():
    """ec.rs:895-905 (G1: 4 if bits >= 130, 3 if >= 34, else 2) and 1586-1596 (G2: 103 / 37)"""
    vals = [0, 1, 2, 3, (1 << 33) - 1, 1 << 33, (1 << 34) - 1, 1 << 36, (1 << 37) - 1, 1 << 102, (1 << 103) - 1,
            1 << 103, 1 << 129, (1 << 130) - 1, 1 << 130, m.R_ORDER - 1]
    k = np.array([m.limbs64(v, 4) for v in vals], dtype=np.uint64)
    assert engine._num_bits(k).tolist() == [v.bit_length() for v in vals]
    assert engine.G1.recommended_wnaf_for_scalar(k).tolist() == [m.g1_recommended_wnaf_for_scalar(v) for v in vals]
    assert engine.G2.recommended_wnaf_for_scalar(k).tolist() == [m.g2_recommended_wnaf_for_scalar(v) for v in vals]
    lib = o.lib()
    import ctypes
    for row, v in zip(k, vals):
        assert lib.oracle_g1_window_for_scalar(row.ctypes.data_as(ctypes.c_void_p)) == m.g1_recommended_wnaf_for_scalar(v)
        assert lib.oracle_g2_window_for_scalar(row.ctypes.data_as(ctypes.c_void_p)) == m.g2_recommended_wnaf_for_scalar(v)


def test_recommended_wnaf_for_num_scalars():
    """ec.rs:907-921 / 1598-1612"""
    for n in (0, 1, 2, 3, 4, 7, 8, 21, 44, 121, 274, 564, 1631, 3129, 7934, 62570, 10**6):
        assert engine.G1.recommended_wnaf_for_num_scalars(n) == m.recommended_wnaf_for_num_scalars(m.G1_NUM_SCALARS_REC, n)
        assert engine.G2.recommended_wnaf_for_num_scalars(n) == m.recommended_wnaf_for_num_scalars(m.G2_NUM_SCALARS_REC, n)
    assert engine.G1.recommended_wnaf_for_num_scalars(1) == 4 and engine.G1.recommended_wnaf_for_num_scalars(10**6) == 16
    assert engine.G2.recommended_wnaf_for_num_scalars(10**6) == 15


def test_into_projective_rows():
    p = dg.g1_affine_points(5, 3, infinity_at=(2,))
    assert np.array_equal(engine.G1Affine.into_projective(p), o.g1_from_affine(p))
    q = dg.g2_affine_points(5, 4, infinity_at=(0,))
    assert np.array_equal(engine.G2Affine.into_projective(q), o.g2_from_affine(q))


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [pdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_wnaf_form_matches_model():
    """wnaf.rs:18-43 digits: C oracle == big-int model, for every window the reference tests (2..13)"""
    ks = dg.rand_scalars(24, 5)
    for w in range(2, 14):
        for row in ks:
            k = m.from_limbs64(row)
            assert o.wnaf_form(row, w) == m.wnaf_form(k, w)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = dg.g1_affine_points(n, 31, infinity_at=(1,))
    q = dg.g2_affine_points(n, 32, infinity_at=(n - 2,))
    lo, hi = pdist.shard_range(n, rank, world)

    def local_product(ps, qs):
        return torch.from_numpy(o.multi_miller_product(ps, qs, 1).view(np.int64))

    def merge(parts):
        f = parts.numpy().view(np.uint64)
        acc = f[:1].copy()
        for i in range(1, f.shape[0]):
            acc = o.fq12_op("mul", acc, f[i:i + 1])[0]
        return torch.from_numpy(acc.view(np.int64))

    got = pdist.multi_miller_loop_sharded(local_product, merge, p[lo:hi], q[lo:hi])
    ret[rank] = got.numpy().view(np.uint64).tobytes()
    dist.destroy_process_group()


def test_sharded_multi_miller_gloo_world2():
    """N > 1 path on CPU: 2 gloo ranks each reduce their shard, all-gather the 576-byte partials and
    merge; every rank ends with the reference's single-accumulator Miller value."""
    import torch.multiprocessing as mp
    n, world = 10, 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, n, ret)) for r in range(world)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(120)
        assert p_.exitcode == 0
    p = dg.g1_affine_points(n, 31, infinity_at=(1,))
    q = dg.g2_affine_points(n, 32, infinity_at=(n - 2,))
    want = o.multi_miller_loop(p, q).tobytes()        # the literal Engine::miller_loop over all pairs
    assert ret[0] == want and ret[1] == want


class _OracleEngine:
    """the three DeviceEngine methods pairing_product_sharded drives, computed by the oracle (there is no GPU here)"""

    def multi_miller_loop(self, p, q):
        import torch
        return torch.from_numpy(o.multi_miller_product(np.asarray(p), np.asarray(q), 1).view(np.int64))

    def fq12_product_tail(self, parts, final_exp=False):
        import torch
        f = parts.numpy().view(np.uint64)
        acc = f[:1].copy()
        for i in range(1, f.shape[0]):
            acc = o.fq12_op("mul", acc, f[i:i + 1])[0]
        ok = np.ones(1, dtype=np.uint8)
        if final_exp:
            acc, ok = o.final_exponentiation(acc, 1)
        return torch.from_numpy(acc.view(np.int64)), torch.from_numpy(ok)

    def pairing_product(self, p, q):
        return self.fq12_product_tail(self.multi_miller_loop(p, q), final_exp=True)


def _worker_product(rank, world, port, n, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = dg.g1_affine_points(n, 33, infinity_at=(2,))
    q = dg.g2_affine_points(n, 34)
    lo, hi = pdist.shard_range(n, rank, world)
    gt, ok = pdist.pairing_product_sharded(_OracleEngine(), p[lo:hi], q[lo:hi])
    ret[rank] = (gt.numpy().view(np.uint64).tobytes(), int(ok[0]))
    dist.destroy_process_group()


def test_sharded_pairing_product_gloo_world2():
    """BASELINE configs[2] in its stated form on 2 gloo ranks: ONE product sharded over the ranks, 576-byte partials
    all-gathered, one final exponentiation == final_exponentiation(miller_loop(all pairs)) of the reference"""
    import torch.multiprocessing as mp
    n, world = 9, 2
    port = _free_port()
    ret = mp.Manager().dict()
    procs = [mp.get_context("spawn").Process(target=_worker_product, args=(r, world, port, n, ret)) for r in range(world)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(180)
        assert p_.exitcode == 0
    p = dg.g1_affine_points(n, 33, infinity_at=(2,))
    q = dg.g2_affine_points(n, 34)
    want, wok = o.final_exponentiation(o.multi_miller_loop(p, q), 1)
    assert ret[0] == ret[1] == (want.tobytes(), int(wok[0]))


def test_shard_ranges_partition_every_size():
    """dist.shard_range (and mgpu.cu's shard_of, the same formula): contiguous, exhaustive, sizes within one of each other"""
    for n in (0, 1, 7, 8, 9, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            r = [pdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_carry_schedule_of_the_montgomery_products_on_the_relaxed_range():
    """tools/emulate_fp.py replays fp_mul / fp_mul2 (pairing_b200/csrc/fp.cuh) word by word on operands in [0, 2q]: the
    result is congruent, stays <= 2q without a conditional subtraction, and every carry the PTX drops is zero."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "tools", "emulate_fp.py")], text=True)
    assert "fp_mul:" in out and "fp_mul2:" in out and "cases ok" in out


def test_rust_shim_binds_every_host_entry_point():
    """rust/src/ffi.rs declares every host-buffer function of include/pairing_b200.h (the `_dev` variants, scratch-size
    helpers and the measurement hook are for the bench harness, not for the crate)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "pairing_b200.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(bls_[a-z0-9_]+)\s*\(", hdr)))
    ffi = open(os.path.join(root, "rust", "src", "ffi.rs")).read()
    skip = ("_dev", "_scratch_bytes", "bls_imad_peak", "bls_ctx_device", "bls_ctx_sm_count", "bls_ctx_launch_count", "bls_field_op_batch",
            "bls_pair_field_op_batch", "bls_ctx_set_latency_path_limits", "bls_ctx_trim", "bls_mgpu_ctx", "bls_mgpu_last_phase_ms")
    missing = [n for n in names if not n.endswith(skip[:2]) and n not in skip and ("fn %s(" % n) not in ffi]
    assert not missing, missing


def test_bench_reference_arm_runs_on_cpu_and_prints_the_contract_line():
    """`bench.py --impl reference` (the C restatement on the host cores) needs no GPU and prints one JSON line with the
    contract's keys; under torchrun only rank 0 prints."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                                  text=True, timeout=300)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairings/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                                  text=True, timeout=60, env=env)
    assert out.strip() == ""


def test_line_pair_product_is_two_mul_by_014():
    """The multi-pairing kernels multiply the line values of two pairs with each other first (pair_tower.cuh:
    p12_mul_by_line_pair -- 6 Fq2 products for sparse x sparse, 17 for f x the 5-coefficient result).  The big-int model
    of exactly those formulas equals two mul_by_014 (fq12.rs:34-48) on random and degenerate operands, i.e. the kernel
    changes the order of the multiplications, not the value of f."""
    import random
    import bls_model as m
    rnd = random.Random(0xE11)
    r2 = lambda: (rnd.randrange(m.Q), rnd.randrange(m.Q))
    r6 = lambda: (r2(), r2(), r2())
    add, sub, mul = m.fq2_add, m.fq2_sub, m.fq2_mul

    def line_pair(f, l, k):
        m00, m11, m44 = mul(l[0], k[0]), mul(l[1], k[1]), mul(l[2], k[2])
        c01 = sub(sub(mul(add(l[0], l[1]), add(k[0], k[1])), m00), m11)
        d1 = sub(sub(mul(add(l[0], l[2]), add(k[0], k[2])), m00), m44)
        d2 = sub(sub(mul(add(l[1], l[2]), add(k[1], k[2])), m11), m44)
        lm0 = (add(m00, m.fq2_mul_by_nonresidue(m44)), c01, m11)
        aa = m.fq6_mul(f[0], lm0)
        bb = m.fq6_mul_by_nonresidue(m.fq6_mul_by_01(f[1], d1, d2))
        s = m.fq6_mul(m.fq6_add(f[1], f[0]), (lm0[0], add(lm0[1], d1), add(lm0[2], d2)))
        return (m.fq6_add(m.fq6_mul_by_nonresidue(bb), aa), m.fq6_sub(m.fq6_sub(s, aa), bb))

    one, zero = (1, 0), (0, 0)
    cases = [((r6(), r6()), (r2(), r2(), r2()), (r2(), r2(), r2())) for _ in range(12)]
    cases.append(((r6(), r6()), (r2(), r2(), r2()), (one, zero, zero)))          # a skipped pair: the line is one
    cases.append(((r6(), r6()), (one, zero, zero), (one, zero, zero)))
    cases.append((((one, zero, zero), (zero, zero, zero)), (r2(), r2(), r2()), (r2(), zero, r2())))
    for f, l, k in cases:
        assert line_pair(f, l, k) == m.fq12_mul_by_014(m.fq12_mul_by_014(f, *l), *k)


def test_generated_squaring_schedule():
    """tools/gen_fp_sqr.py --check: the committed fp_sqr_gen.cuh is what the generator emits, and the instruction
    list replayed on integers squares correctly on [0, 2q], stays <= 2q and never drops a carry"""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_fp_sqr.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "222 wide MACs" in r.stdout


def test_division_step_inversion_host_build_matches_pow():
    """pairing_b200/csrc/fp_inv_gcd.cuh (the Fq inversion every CUDA kernel runs; fq.rs:849-902 is a binary extended
    Euclid with the same unique result) is plain integer C: compiled for the host here and compared with pow(a, -1, q)
    on edge operands, short operands of every bit length and random ones; the embedded 30-bit-limb constants are
    recomputed."""
    import random
    import re
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "pairing_b200", "csrc", "fp_inv_gcd.cuh")).read()
    def limbs(name):
        body = re.search(r"#define %s\s*\\\n\s*\{(.*?)\}" % name, hdr, re.S).group(1).replace("\\", "")
        return sum(int(t, 16) << (30 * i) for i, t in enumerate(body.replace(",", " ").split()))
    R = 1 << 384
    assert limbs("BLS_GCD_Q30") == m.Q and limbs("BLS_GCD_R2_30") == R * R % m.Q
    assert int(re.search(r"QINV30 = (0x[0-9a-f]+)u", hdr).group(1), 16) == pow(m.Q, -1, 1 << 30)
    rng = random.Random(0xDEC0DE)
    vals = [0, 1, 2, 3, m.Q - 1, m.Q - 2, (m.Q + 1) // 2, R % m.Q, R * R % m.Q]
    vals += [rng.randrange(1 << k) % m.Q for k in range(1, 382) for _ in range(2)]
    vals += [rng.randrange(m.Q) for _ in range(3000)]
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "inv")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(root, "tests", "cpp", "test_fp_inv_gcd.cpp")],
                              stderr=subprocess.DEVNULL)
        out = subprocess.run([exe], input="".join("%096x\n" % v for v in vals), capture_output=True, text=True, check=True).stdout.split("\n")
    assert len(out) >= len(vals)
    for v, line in zip(vals, out):
        got, batches = line.split()
        assert int(got, 16) == (pow(v, -1, m.Q) * R * R % m.Q if v else 0), hex(v)
        assert int(batches) <= 38          # (49 * 381 + 57) / 17 = 1101 division steps bound the loop: 37 batches of 30


def test_mul_assign_is_wnaf_exp_over_the_bits():
    """k_pt_mul runs CurveProjective::mul_assign (ec.rs:534-553) on the wNAF kernel's decoupled-lane runner: the claim that its
    operation sequence is wnaf_exp's (wnaf.rs:49-71) with the scalar's bits as digits and a one-entry table, on the model --
    the same Jacobian TRIPLE, not just the same point."""
    import random
    rng = random.Random(41)
    for F, gen in ((m._F1, m.G1_GEN_AFFINE), (m._F2, m.G2_GEN_AFFINE)):
        base = m.pt_double(F, m.pt_from_affine(F, gen))               # Z != 1
        for k in (0, 1, 2, 3, m.R_ORDER - 1, (1 << 255) - 1, rng.randrange(m.R_ORDER), rng.randrange(1 << 64)):
            bits = [(k >> i) & 1 for i in range(k.bit_length())]
            assert m.wnaf_exp(F, [base], bits) == m.pt_mul(F, base, k)
        zero = m.pt_zero(F)
        assert m.wnaf_exp(F, [zero], [1, 0, 1, 1]) == m.pt_mul(F, zero, 0b1101)

"""Device-resident throughput of every path of SURVEY.md section 8 (configs 2-5) + a bit-exact spot check
against the oracle.  Tuning helper (bench.py is the contract benchmark).
    python tools/bench_paths.py [--log2 18] [--skip pairing,...]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import bench
from pairing_b200.device import DeviceEngine
import oracle_lib as o

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=18)
ap.add_argument("--pair-log2", type=int, default=16)
ap.add_argument("--skip", default="")
ap.add_argument("--check", type=int, default=64)
args = ap.parse_args()
skip = set(args.skip.split(","))
eng = DeviceEngine(device=0)
ctx = eng.ctx
peak, _ = ctx.imad_peak(0, 4000)
print("IMAD.WIDE peak %.3f T MAC32/s" % (peak / 1e12))
n = 1 << args.log2
npair = 1 << args.pair_log2
pa, qa, g1_jac, ks = bench.make_inputs(eng, max(n, npair), bench.SEED, torch, np)
# non-normalised Jacobian bases (Z != 1), as G::rand yields: double the normalised points once
def denorm(eng, jac, g2):
    W = 36 if g2 else 18
    out = torch.empty_like(jac)
    two = torch.zeros((jac.shape[0], 4), dtype=torch.int64, device=jac.device); two[:, 0] = 3
    return (eng.g2_wnaf_mul if g2 else eng.g1_wnaf_mul)(jac, two, 2, out)
g1b = denorm(eng, g1_jac, False)
g2_jac = torch.zeros((qa.shape[0], 36), dtype=torch.int64, device=qa.device)
g2_jac[:, :24] = qa[:, :24]; g2_jac[:, 24:30] = g1_jac[:1, 12:18]  # z = one
g2b = denorm(eng, g2_jac[:n].contiguous(), True)
torch.cuda.synchronize()

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def report(name, units, ms, mac_per_unit):
    rate = units / (ms * 1e-3)
    print("%-34s n=%8d  %9.3f ms  %12.0f /s  %6.2f T MAC32/s  frac %.3f" % (name, units, ms, rate, rate * mac_per_unit / 1e12, rate * mac_per_unit / peak))

C = args.check
if "pairing" not in skip:
    out = torch.empty((npair, 72), dtype=torch.int64, device=eng.device)
    ms = timeit(lambda: eng.pairing(pa[:npair], qa[:npair], out))
    report("pairing (config 2)", npair, ms, 20621 * 300)
    want = o.pairing(pa[:C].cpu().numpy().view(np.uint64), qa[:C].cpu().numpy().view(np.uint64), o.default_threads())
    assert np.array_equal(out[:C].cpu().numpy().view(np.uint64), want), "pairing mismatch"
    ms = timeit(lambda: eng.miller_loop_batch(pa[:npair], qa[:npair], out))
    report("miller_loop batch", npair, ms, 5156 * 300)
    f = out.clone(); out2 = torch.empty_like(out)
    ms = timeit(lambda: eng.final_exponentiation(f, out2))
    report("final_exponentiation batch", npair, ms, 13705 * 300)
if "mm" not in skip:
    nm = min(n, pa.shape[0])
    res = []
    ms = timeit(lambda: res.append(eng.multi_miller_loop(pa[:nm], qa[:nm])))
    report("multi_miller_loop (config 3)", nm, ms, 4684 * 300)
    print("  multi-Miller value checksum %016x" % int(res[-1].cpu().numpy().view(np.uint64).sum(dtype=np.uint64)))
    want = o.multi_miller_loop(pa[:C].cpu().numpy().view(np.uint64), qa[:C].cpu().numpy().view(np.uint64)) if hasattr(o, "multi_miller_loop") else None
    if want is not None:
        got = eng.multi_miller_loop(pa[:C].contiguous(), qa[:C].contiguous()).cpu().numpy().view(np.uint64)
        assert np.array_equal(got.reshape(-1), np.asarray(want).reshape(-1)), "multi-miller mismatch"
if "g1" not in skip:
    wout = torch.empty_like(g1b[:n])
    b, kk = g1b[:n].contiguous(), ks[:n].contiguous()
    ms = timeit(lambda: eng.g1_wnaf_mul(b, kk, 0, wout))
    report("G1 wNAF mul (config 4)", n, ms, 2570 * 300)
    want = o.g1_op("wnaf", b[:C].cpu().numpy().view(np.uint64), k=kk[:C].cpu().numpy().view(np.uint64), threads=o.default_threads())
    assert np.array_equal(wout[:C].cpu().numpy().view(np.uint64), want), "G1 wNAF mismatch"
    ms2 = timeit(lambda: eng.g1_batch_normalization_(wout.clone()))
    report("G1 batch_normalization (+clone)", n, ms2, 7 * 300)
    report("G1 wNAF + normalisation", n, ms + ms2, 2577 * 300)
if "g2" not in skip:
    n2 = min(n, 1 << 20)
    b, kk = g2b[:n2].contiguous(), ks[:n2].contiguous()
    wout = torch.empty_like(b)
    ms = timeit(lambda: eng.g2_wnaf_mul(b, kk, 0, wout))
    report("G2 wNAF mul (config 5)", n2, ms, 6212 * 300)
    want = o.g2_op("wnaf", b[:C].cpu().numpy().view(np.uint64), k=kk[:C].cpu().numpy().view(np.uint64), threads=o.default_threads())
    assert np.array_equal(wout[:C].cpu().numpy().view(np.uint64), want), "G2 wNAF mismatch"
    np2 = min(n2, 1 << 15)
    prep = torch.empty((np2, 68 * 36 + 1), dtype=torch.int64, device=eng.device)
    q2 = qa[:np2].contiguous()
    ms = timeit(lambda: eng.g2_prepare(q2, prep))
    report("G2 prepare (config 5)", np2, ms, 1760 * 300)
    want = o.g2_prepare(q2[:8].cpu().numpy().view(np.uint64))
    assert np.array_equal(prep[:8].cpu().numpy().view(np.uint64), want), "G2 prepare mismatch"
print("all spot checks bit-exact")

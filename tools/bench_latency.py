"""Latency path vs throughput path (tuning helper): device-timed pairing batches of n = 1 .. 16384 on the warp-cooperative
kernel and on the lane-pair kernel, the single final exponentiation, and the phases of a 2^20-pair product.
    python tools/bench_latency.py [--mm-log2 20]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import bench
from pairing_b200.device import DeviceEngine

ap = argparse.ArgumentParser()
ap.add_argument("--mm-log2", type=int, default=20)
ap.add_argument("--only-mm", action="store_true")
ap.add_argument("--mm-sweep", action="store_true")
args = ap.parse_args()
eng = DeviceEngine(device=0)
ctx = eng.ctx
N = 1 << 14
pa, qa, _, _ = bench.make_inputs(eng, N, bench.SEED, torch, np)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = torch.empty((N, 72), dtype=torch.int64, device=eng.device)
print("%8s %12s %12s" % ("n", "wide ms", "lane-pair ms"))
for n in (() if args.only_mm else (1, 10, 100, 500, 1000, 2000, 3000, 4096, 6000, 8192, 16384)):
    p, q = pa[:n].contiguous(), qa[:n].contiguous()
    ctx.set_latency_path_limits(1 << 30, 1 << 30)
    w = timeit(lambda: eng.pairing(p, q, out[:n]))
    ref = out[:n].clone()
    ctx.set_latency_path_limits(0, 0)
    l = timeit(lambda: eng.pairing(p, q, out[:n]), reps=2)
    assert torch.equal(ref, out[:n]), "paths differ at n=%d" % n
    print("%8d %12.3f %12.3f" % (n, w, l))
f = eng.miller_loop_batch(pa[:4096].contiguous(), qa[:4096].contiguous())
for n in (() if args.only_mm else (1, 100, 1000, 4096)):
    ctx.set_latency_path_limits(1 << 30, 1 << 30)
    w = timeit(lambda: eng.final_exponentiation(f[:n].contiguous()))
    ctx.set_latency_path_limits(0, 0)
    l = timeit(lambda: eng.final_exponentiation(f[:n].contiguous()), reps=2)
    print("final_exp n=%5d  wide %8.3f ms   lane-pair %8.3f ms" % (n, w, l))
ctx.set_latency_path_limits(4096, 4096)
for cnt in (8, 296):
    part = f[:cnt].contiguous()
    print("product_tail count=%3d: %.3f ms without, %.3f ms with the final exponentiation" % (
        cnt, timeit(lambda: eng.fq12_product_tail(part, False)), timeit(lambda: eng.fq12_product_tail(part, True))))
nm = 1 << args.mm_log2
pm = pa.repeat(nm // N, 1).contiguous(); qm = qa.repeat(nm // N, 1).contiguous()
eng._buf("mm", ctx.multi_miller_scratch_bytes(nm))
for frac in (() if args.mm_sweep else (1, 2, 4, 8)):
    n = nm // frac
    a = timeit(lambda: eng.multi_miller_loop(pm[:n], qm[:n]), reps=3)
    b = timeit(lambda: eng.pairing_product(pm[:n], qm[:n]), reps=3)
    print("product of 2^%d / %d pairs: multi_miller_loop %.3f ms (%.2f M pairs/s), + final exponentiation %.3f ms" % (args.mm_log2, frac, a, n / a / 1e3, b))

if args.mm_sweep:     # trip counts (pairs per lane pair) 1 .. 14, odd and even, and the 2^20 case
    T = ctx.sm_count * 2 * 64
    big = 1 << 20
    pm = pa.repeat(big // N, 1).contiguous(); qm = qa.repeat(big // N, 1).contiguous()
    eng._buf("mm", ctx.multi_miller_scratch_bytes(big))
    for per in ((6, 7, 8, 14) if args.only_mm else (1, 2, 3, 4, 5, 6, 7, 8, 13, 14)):
        n = T * per
        a = timeit(lambda: eng.multi_miller_loop(pm[:n], qm[:n]), reps=3)
        print("trips %2d (n = %7d): multi_miller_loop %8.3f ms  (%.3f ms per trip)" % (per, n, a, a / per))
    a = timeit(lambda: eng.multi_miller_loop(pm, qm), reps=3)
    print("n = 2^20: multi_miller_loop %8.3f ms" % a)

"""Diagnostic: the lane-pair pairing / final-exponentiation kernels on a small batch with and without infinity pairs,
compared with the oracle.  Run under `timeout` (tuning helper for new kernel builds).
    python tools/diag_pairing.py [n] [inf]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
import oracle_lib as o
from pairing_b200.device import DeviceEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
inf = len(sys.argv) > 2 and sys.argv[2] == "inf"
eng = DeviceEngine(device=0)
eng.ctx.set_latency_path_limits(0, 0)
pa, qa, g1j, ks = bench.make_inputs(eng, n, bench.SEED, torch, np)
if inf:
    pa[5, 12] = 1; qa[9, 24] = 1
print("inputs ready", flush=True)
ml = eng.miller_loop_batch(pa, qa); torch.cuda.synchronize(); print("miller ok", flush=True)
fe, ok = eng.final_exponentiation(ml); torch.cuda.synchronize(); print("final exp ok", flush=True)
gt = eng.pairing(pa, qa); torch.cuda.synchronize(); print("pairing ok", flush=True)
C = min(n, 64)
want = o.pairing(pa[:C].cpu().numpy().view(np.uint64), qa[:C].cpu().numpy().view(np.uint64), o.default_threads())
print("final_exp == oracle:", np.array_equal(fe[:C].cpu().numpy().view(np.uint64), want), " pairing == oracle:", np.array_equal(gt[:C].cpu().numpy().view(np.uint64), want), " equal:", bool(torch.equal(gt, fe)))

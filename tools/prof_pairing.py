"""Small driver for ncu: generate n pairs on the device, run the pairing kernel twice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
from pairing_b200.device import DeviceEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
what = sys.argv[2] if len(sys.argv) > 2 else "pairing"
eng = DeviceEngine(device=0)
pa, qa, g1j, ks = bench.make_inputs(eng, n, bench.SEED, torch, np)
out = torch.empty((n, 72), dtype=torch.int64, device=eng.device)
for _ in range(2):
    if what == "pairing":
        eng.pairing(pa, qa, out)
    elif what == "wnaf":
        eng.g1_wnaf_mul(g1j, ks)
    elif what == "mm":
        eng.multi_miller_loop(pa, qa)
torch.cuda.synchronize()
print("done", what, n)

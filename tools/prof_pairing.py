"""Small driver for ncu: generate n inputs on the device, run one path twice.
    python tools/prof_pairing.py <n> <pairing|pairing_wide|wnaf|g2wnaf|mm|product|sharedq|fixed|pow|finalexp>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
from pairing_b200.device import DeviceEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
what = sys.argv[2] if len(sys.argv) > 2 else "pairing"
eng = DeviceEngine(device=0)
gen = max(min(n, 1 << 16), 1 << 10)
pa, qa, g1j, ks = bench.make_inputs(eng, gen, bench.SEED, torch, np)
def tile(t, m):
    return t.repeat((m + t.shape[0] - 1) // t.shape[0], 1)[:m].contiguous()
pa, qa, g1j, ks = tile(pa, n), tile(qa, n), tile(g1j, n), tile(ks, n)
out = torch.empty((n, 72), dtype=torch.int64, device=eng.device)
if what == "pairing":
    eng.ctx.set_latency_path_limits(0, 0)
if what == "pairing_wide":
    eng.ctx.set_latency_path_limits(1 << 30, 1 << 30)
if what in ("sharedq",):
    q1 = eng.g2_prepare(qa[:1].contiguous())
if what == "fixed":
    table = eng.wnaf_table(g1j[5:6].contiguous(), 16)
if what == "g2wnaf":
    q2 = torch.zeros((n, 36), dtype=torch.int64, device=eng.device)
    q2[:, :24] = qa[:, :24]; q2[:, 24:30] = g1j[:1, 12:18]
if what in ("pow", "finalexp"):
    eng.ctx.set_latency_path_limits(0, 0)
    f = eng.pairing(pa, qa) if what == "pow" else eng.miller_loop_batch(pa, qa)      # GT operands for the powers (cyclotomic path)
for _ in range(2):
    if what in ("pairing", "pairing_wide"):
        eng.pairing(pa, qa, out)
    elif what == "wnaf":
        eng.g1_wnaf_mul(g1j, ks)
    elif what == "g2wnaf":
        eng.g2_wnaf_mul(q2, ks)
    elif what == "mm":
        eng.multi_miller_loop(pa, qa)
    elif what == "product":
        eng.pairing_product(pa, qa)
    elif what == "sharedq":
        eng.pairing_shared_q(pa, q1, out)
    elif what == "fixed":
        eng.wnaf_fixed_base(table, 16, ks)
    elif what == "pow":
        eng.fq12_pow(f, ks)
    elif what == "finalexp":
        eng.final_exponentiation(f)
torch.cuda.synchronize()
print("done", what, n)

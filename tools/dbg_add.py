import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np
import pairing_b200._native as nat, oracle_lib as o, datagen as dg
a = dg.g1_points(32, 22); b = dg.g1_points(32, 23)
cases = {"plain": (a, b)}
a2 = a.copy(); a2[0] = 0; a2[0, 6:12] = np.array(__import__("bls_model").limbs64(__import__("bls_model").MONT_R), dtype=np.uint64)
cases["a_inf"] = (a2, b)
cases["b_inf"] = (b, a2)
cases["equal"] = (a, a.copy())
cases["neg"] = (a, o.g1_op("negate", a))
for name, (x, y) in cases.items():
    try:
        ctx = nat.Context(0)
        got = ctx.g1_op("add", x, y)
        print(name, "ok" if np.array_equal(got, o.g1_op("add", x, y)) else "MISMATCH", flush=True)
    except Exception as e:
        print(name, "EXC", str(e)[:150], flush=True)
        break

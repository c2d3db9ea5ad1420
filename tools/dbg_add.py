import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np
import pairing_b200._native as nat, oracle_lib as o, datagen as dg
n = 128
a = dg.g1_points(n, 22, infinity_at=(0, 5)); b = dg.g1_points(n, 23, infinity_at=(1, 5))
def run(name, op, x, y):
    try:
        ctx = nat.Context(0)
        got = ctx.g1_op(op, x, y)
        print(name, op, "ok" if np.array_equal(got, o.g1_op(op, x, y)) else "MISMATCH", flush=True)
        return True
    except Exception as e:
        print(name, op, "EXC", str(e)[:120], flush=True)
        return False
ok = run("inf-mix", "add", a, b)
b1 = b.copy(); b1[10] = a[10]
ok = ok and run("+equal", "add", a, b1)
b2 = b1.copy(); b2[11] = o.g1_op("negate", a[11:12])[0]
ok = ok and run("+neg", "add", a, b2)
b3 = b2.copy(); b3[12] = o.g1_op("double", a[12:13])[0]
ok = ok and run("+dbl", "add", a, b3)
ok = ok and run("all", "sub", a, b3)
ok = ok and run("plain", "sub", dg.g1_points(n, 1), dg.g1_points(n, 2))

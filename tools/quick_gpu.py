"""Quick GPU probe: integer-multiply peaks + rough throughput of the main kernels (host-API timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import pairing_b200._native as nat
import oracle_lib as o

ctx = nat.Context(0)
print("sm_count", ctx.sm_count)
for v, name in ((0, "IMAD.WIDE.U32 independent"), (1, "fp_mul chain (300 MAC32/mul)"), (2, "IMAD 32-bit")):
    macs, ms = ctx.imad_peak(v, 4000)
    print("peak[%d] %-32s %8.3f T MAC/s  (%.2f ms)" % (v, name, macs / 1e12, ms))
g1, g2 = o.generators()
for n in (1 << 12, 1 << 14, 1 << 16):
    P = np.repeat(g1, n, 0); Q = np.repeat(g2, n, 0)
    ctx.pairing(P[:256], Q[:256])
    t = time.time(); out = ctx.pairing(P, Q); dt = time.time() - t
    print("pairing n=%d: %.1f ms -> %.0f pairings/s (host API, incl. copies)" % (n, dt * 1e3, n / dt))
    t = time.time(); out = ctx.miller_loop(P, Q); dt = time.time() - t
    print("miller  n=%d: %.1f ms -> %.0f /s" % (n, dt * 1e3, n / dt))

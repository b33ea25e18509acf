"""Scratch GPU probe (not a test): first timings per stage + IMAD peak.  python tests/gpu_probe.py [log2n]"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-nullifier-sig_b200"))
import numpy as np
import plume_b200 as P

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << lg
ctx = P.PlumeContext(0)
print("imad peak LP/s: %.3e" % ctx.measure_imad_peak(4096))
rng = np.random.default_rng(1)
msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
for ver in (1, 2):
    o = ctx.sign_batch(ver, msgs, sk, r)   # warm
    ctx.set_profiling(True)
    t = time.time(); o = ctx.sign_batch(ver, msgs, sk, r); dt = time.time() - t
    print("sign v%d n=%d: %.1f ms wall e2e -> %.3e sig/s; status ok=%d" % (ver, n, dt * 1e3, n / dt, int((o["status"] == 0).sum())))
    tot = 0
    for st in ("sign_fixed", "sign_h2c", "sign_varbase", "sign_final", "binv"):
        ms, k = ctx.stage_ms(st); tot += ms
        print("   %-13s %8.3f ms  (%d launches)" % (st, ms, k))
    print("   kernels total %.3f ms -> %.3e sig/s" % (tot, n / (tot * 1e-3)))
    ctx.set_profiling(True)
    t = time.time(); ok = ctx.verify_batch(ver, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"]); dt = time.time() - t
    print("verify v%d: %.1f ms wall -> %.3e /s; ok=%d" % (ver, dt * 1e3, n / dt, int(ok.sum())))
    tot = 0
    for st in ("verify_h2c", "verify_mul_b", "verify_mul_a", "verify_final", "binv"):
        ms, k = ctx.stage_ms(st); tot += ms
        print("   %-13s %8.3f ms  (%d launches)" % (st, ms, k))
    print("   kernels total %.3f ms -> %.3e ver/s" % (tot, n / (tot * 1e-3)))
ctx.set_profiling(False)

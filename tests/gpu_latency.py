"""Scratch GPU probe (not a test): wall-clock latency of small batches through the host-pointer API.  python tests/gpu_latency.py"""
import sys, time
sys.path.insert(0, "zk-nullifier-sig_b200"); sys.path.insert(0, "oracle")
import numpy as np, plume_b200 as P
ctx = P.PlumeContext(0)
rng = np.random.default_rng(3)
for n in (1, 32, 1024, 32768):
    msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
    r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
    o = ctx.sign_batch(1, msgs, sk, r)
    t = time.perf_counter()
    for _ in range(10): o = ctx.sign_batch(1, msgs, sk, r)
    ts = (time.perf_counter() - t) / 10
    ok = ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    t = time.perf_counter()
    for _ in range(10): ok = ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    tv = (time.perf_counter() - t) / 10
    print("n=%6d  sign %.3f ms (%.0f/s)  verify %.3f ms (%.0f/s)  all ok %s" % (n, ts * 1e3, n / ts, tv * 1e3, n / tv, bool(ok.all())))

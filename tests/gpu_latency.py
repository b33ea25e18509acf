"""Scratch GPU probe (not a test): wall-clock latency of small batches through the host-pointer API, with the small-batch
kernels (k_team.cu) on and off, and the per-stage device times of a batch of one.  python tests/gpu_latency.py"""
import os, sys, time
sys.path.insert(0, "zk-nullifier-sig_b200"); sys.path.insert(0, "oracle")
import numpy as np, plume_b200 as P

STAGES = ["sign_fixed", "sign_h2c", "sign_tab", "sign_varbase", "sign_final", "verify_h2c", "verify_tab_b", "verify_mul_b",
          "verify_mul_a", "verify_final", "binv"]


def make(team_max, window=16):
    os.environ["PLUME_TEAM_MAX"] = str(team_max)
    try:
        return P.PlumeContext(0, fixed_window_bits=window)
    finally:
        del os.environ["PLUME_TEAM_MAX"]


rng = np.random.default_rng(3)
# (the library caps PLUME_TEAM_MAX at 4 096: above that size both configurations run the throughput kernels)
CONFIGS = (("team kernels", 4096),) if os.environ.get("LAT_ONLY_TEAM") else (("team kernels", 4096), ("throughput kernels", 0))
SIZES = (1, 1024, 4096) if os.environ.get("LAT_ONLY_TEAM") else (1, 32, 1024, 4096, 8192)
for label, team_max in CONFIGS:
    ctx = make(team_max)
    for n in SIZES:
        msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
        r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
        o = ctx.sign_batch(1, msgs, sk, r)
        reps = 20
        t = time.perf_counter()
        for _ in range(reps): o = ctx.sign_batch(1, msgs, sk, r)
        ts = (time.perf_counter() - t) / reps
        ok = ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
        t = time.perf_counter()
        for _ in range(reps): ok = ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
        tv = (time.perf_counter() - t) / reps
        print("%-18s n=%6d  sign %.3f ms  verify %.3f ms  all ok %s" % (label, n, ts * 1e3, tv * 1e3, bool(ok.all())), flush=True)
    # stage times of a batch of one (events serialise the two streams of the verifier: a breakdown, not the latency)
    msgs = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    sk = rng.integers(0, 256, (1, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
    ctx.set_profiling(True)
    for _ in range(10):
        o = ctx.sign_batch(1, msgs, sk, sk)
        ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    print("  stage us (batch of one):", "  ".join("%s %.0f" % (s, 1e3 * ctx.stage_ms(s)[0] / max(1, ctx.stage_ms(s)[1])) for s in STAGES), flush=True)
    ctx.set_profiling(False)
    ctx.close()

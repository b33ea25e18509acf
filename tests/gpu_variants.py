"""Scratch GPU experiment driver (not a test): times the stage kernels of several builds of the library.
    python tests/gpu_variants.py [log2n] lib1.so lib2.so ...
Each build runs in its own process (PLUME_B200_LIB); outputs are checked against the C oracle on a sample."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%(root)r, "zk-nullifier-sig_b200")); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, plume_b200 as P, c_oracle
lg = %(lg)d; n = 1 << lg
ctx = P.PlumeContext(0)
rng = np.random.default_rng(1)
msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
res = {"lib": os.environ.get("PLUME_B200_LIB", "default")}
for rep in range(2):
    o = ctx.sign_batch(1, msgs, sk, r)
    ok = ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
ctx.set_profiling(True)
for rep in range(3):
    o = ctx.sign_batch(1, msgs, sk, r)
    ok = ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
for st in ("sign_fixed", "sign_h2c", "sign_varbase", "sign_final", "verify_h2c", "verify_muls", "verify_mul_a", "verify_tab_b", "verify_mul_b", "verify_final", "binv"):
    ms, k = ctx.stage_ms(st); res[st] = round(ms / 3, 3)          # per batch of n items (3 timed repetitions)
res["verify_muls"] = round(res["verify_muls"] + res["verify_mul_a"] + res["verify_tab_b"] + res["verify_mul_b"], 3)
res["sign_ms"] = round(res["sign_fixed"] + res["sign_h2c"] + res["sign_varbase"] + res["sign_final"] + res["binv"] * 3 / 5, 3)
res["verify_ms"] = round(res["verify_h2c"] + res["verify_muls"] + res["verify_final"] + res["binv"] * 2 / 5, 3)
res["sign_per_s"] = round(n / res["sign_ms"] * 1e3); res["verify_per_s"] = round(n / res["verify_ms"] * 1e3)
m = 256
want = c_oracle.sign_batch(1, msgs[:m], sk[:m], r[:m], threads=16)
res["bit_exact"] = bool(all(np.array_equal(o[k][:m], want[k]) for k in want)) and bool(ok.all())
print("RESULT " + json.dumps(res))
'''

lg = int(sys.argv[1])
for lib in sys.argv[2:]:
    env = dict(os.environ)
    if lib != "default":
        env["PLUME_B200_LIB"] = os.path.join(ROOT, "zk-nullifier-sig_b200", lib)
    p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lg": lg}], env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    print(lib, line[0][7:] if line else ("FAILED: " + p.stderr[-600:]))

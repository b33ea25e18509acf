"""Scratch GPU probe (not a test): a small pass over every entry point, meant to be run under compute-sanitizer.
    compute-sanitizer --tool memcheck python tests/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zk-nullifier-sig_b200"))
import numpy as np
import plume_b200 as P

n = 300
rng = np.random.default_rng(4)
ctx = P.PlumeContext(0, 8)          # small generator table: the sanitizer is slow
msgs = [bytes(rng.integers(0, 256, int(rng.integers(0, 90)), dtype=np.uint8)) for _ in range(n)]
sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
sk[3] = 0; r[5] = 255
for ver in (1, 2):
    o = ctx.sign_batch(ver, msgs, sk, r)
    ok = ctx.verify_batch(ver, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    z = np.zeros_like(o["pk"])
    ctx.verify_batch(ver, msgs, z, o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    ctx.verify_batch(ver, msgs, o["pk"], z, o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    a = ctx.ark_sign_batch(ver, msgs, o["pk"], sk, r)
    ctx.ark_verify_batch(ver, msgs, o["pk"], a["nullifier"], a["digest_private"], a["s"], a["r_point"], a["hashed_to_curve_r"])
    s1 = ctx.sign_batch_sec1(ver, msgs, sk, r)
    ctx.verify_batch_sec1(ver, msgs, s1["pk"], s1["nullifier"], s1["c"], s1["s"], s1["r_point"], s1["hashed_to_curve_r"])
    print("v%d: %d signed, %d verified" % (ver, int((o["status"] == 0).sum()), int(ok.sum())))
w = ctx.hash_to_curve_witness_batch(msgs)
h = ctx.hash_to_curve_batch(msgs)
assert np.array_equal(w["h"], h)
ctx.registers_batch(o["c"])
c33 = ctx.points_compress(o["pk"])
ctx.points_decompress(c33)
ctx.close()
print("done")

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
pts = ctx.fixed_base_mul_batch(sk)
ctx.hash_to_curve_pk_batch(msgs, ctx.points_compress(pts))
# fixed 32-byte and 65-byte records (the fixed-layout b0 path, word and byte loads), and a batch above the small-batch
# threshold (no second stream) next to the ones above (second stream for G*s - pk*c)
fixed = rng.integers(0, 256, (n, 32), dtype=np.uint8)
o = ctx.sign_batch(1, fixed, sk, r)
ctx.verify_batch(1, fixed, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
ctx.hash_to_curve_batch(rng.integers(0, 256, (n, 65), dtype=np.uint8))
# the small-batch (team) kernels with part of a warp and part of a team's block unused, and the library's self test
for k in (1, 3, 33):
    o = ctx.sign_batch(1, fixed[:k], sk[:k], r[:k])
    ctx.verify_batch(1, fixed[:k], o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
ctx.self_test()
# the throughput kernels on the same small batch (PLUME_TEAM_MAX=0), as a 2^20 batch runs them
os.environ["PLUME_TEAM_MAX"] = "0"
t = P.PlumeContext(0, 8)
del os.environ["PLUME_TEAM_MAX"]
o = t.sign_batch(1, msgs, sk, r)
t.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
t.close()
big = 9000
bm = rng.integers(0, 256, (big, 32), dtype=np.uint8)
bs = rng.integers(0, 256, (big, 32), dtype=np.uint8); bs[:, 0] &= 0x7F
o = ctx.sign_batch(2, bm, bs, bs)
assert ctx.verify_batch(2, bm, o["pk"], o["nullifier"], o["c"], o["s"]).all()
ctx.close()
# the multi-device context (one sub-context per visible GPU; worker threads)
import torch
m = P.PlumeContext(list(range(torch.cuda.device_count())), 8)
o = m.sign_batch(1, msgs, sk, r)
m.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
m.close()
print("done")

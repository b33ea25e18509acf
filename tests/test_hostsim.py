"""Host logic of the CUDA sources, without a GPU: the same .cuh files compiled for the CPU with the
PTX carry-chain wrappers emulated (tests/hostsim).  Checks the limb-level field/scalar arithmetic, the
GLV split and Booth recoding, SHA-256, the SSWU map, both scalar-multiplication paths and the full
staged sign / verify / hash_to_curve pipelines against the oracles, including the edge cases the
reference tests (empty and ragged messages) and those it leaves unpinned (identity points, bad scalars)."""
import ctypes
import hashlib
import random

import numpy as np

import _hostsim as H
import c_oracle
import plume_ref as R

P, N = R.P, R.N
LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
BETA = 0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE


def _pt64(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def test_field_arithmetic_limb_level():
    rnd = random.Random(1)
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, P - 2, 2**255, 2**32 + 977, 2**256 - 2**32 - 978, 0xFFFFFFFF,
            2**224 - 1, (2**256 - 1) ^ (2**128), 2**256 - 2**32 - 977 + 976, (1 << 256) - (1 << 224)]
    vals = edge + [rnd.randrange(2**256) for _ in range(120)]
    for a in vals:
        for b in edge + rnd.sample(vals, 6):
            assert H.fe_op(0, a, b) % P == a * b % P
            assert H.fe_op(2, a, b) % P == (a + b) % P
            assert H.fe_op(3, a, b) % P == (a - b) % P
        assert H.fe_op(1, a) % P == a * a % P
        assert H.fe_op(5, a) == a % P
        assert H.fe_op(8, a) % P == (-a) % P
        for k, op in ((1, 12), (2, 13), (3, 11)):
            assert H.fe_op(op, a) % P == (a << k) % P
        for k in (0, 1, 2, 3, 8, 11, 1771, 65535, 65536):
            assert H.fe_op(6, a, k) % P == a * k % P
        assert bool(H.lib().hs_fe_is_zero(H.limbs(a))) == (a % P == 0)
    for a in vals[:24]:
        if a % P:
            assert H.fe_op(4, a) % P == pow(a, -1, P)
        assert H.fe_op(7, a) % P == pow(a, (P - 3) // 4, P)
        assert H.fe_op(9, a) % P == pow(a, (P + 1) // 4, P)


def test_division_step_inversion():
    """inv.cuh fe_inv_var (Bernstein-Yang division steps, 30 at a time): against pow(x, -1, p) on edge values (0 and p give 0
    through the Fermat fallback), non-canonical representatives, powers of two, values with long runs, and random ones."""
    rnd = random.Random(77)
    vals = [0, P, 1, 2, 3, P - 1, P - 2, P + 1, 2**256 - 1, 2**255, 2**32 + 977, 977, 2**32, 0xFFFFFFFF, (P + 1) // 2, (P - 1) // 2,
            2**224 - 1, (1 << 256) - (1 << 224), 0x5555555555555555555555555555555555555555555555555555555555555555 % 2**256,
            0xAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA]
    vals += [1 << k for k in range(0, 256, 7)] + [(1 << k) - 1 for k in range(2, 256, 11)] + [P - (1 << k) for k in range(0, 250, 13)]
    vals += [rnd.randrange(2**256) for _ in range(20000)] + [rnd.randrange(2**64) for _ in range(200)]
    vals += [rnd.getrandbits(256) & rnd.getrandbits(256) & rnd.getrandbits(256) for _ in range(500)]          # sparse
    vals += [(rnd.getrandbits(256) | rnd.getrandbits(256) | rnd.getrandbits(256)) for _ in range(500)]        # dense
    vals += [pow(a, -1, P) for a in (2, 3, 2**30, 2**60, 2**255)] + [P - pow(2, -k, P) for k in (1, 30, 31, 60)]
    for a in vals:
        got = H.fe_op(14, a) % P
        assert got == (pow(a, -1, P) if a % P else 0), hex(a)


def test_scalar_arithmetic_and_glv():
    rnd = random.Random(2)
    sv = [0, 1, 2, N - 1, N - 2, N // 2, N // 2 + 1, 2**128, 2**129 - 1, LAMBDA, N - LAMBDA] + [rnd.randrange(N) for _ in range(100)]
    for a in sv:
        for b in sv[:10]:
            assert H.sc_op(0, a, b) == a * b % N
            assert H.sc_op(1, a, b) == (a + b) % N
        assert H.sc_op(2, a) == (-a) % N
    for a in (N, N + 1, 2**256 - 1, 5):
        assert H.sc_op(3, a) == a % N
    for _ in range(100):
        x = rnd.choice([rnd.randrange(2**512), 2**512 - 1, (N - 1) ** 2, 2**512 - rnd.randrange(2**20)])
        out = (ctypes.c_uint32 * 8)()
        H.lib().hs_sc_reduce512(H.limbs(x, 16), out)
        assert H.val(out) == x % N
    # the lattice constants used on the device
    assert pow(LAMBDA, 3, N) == 1 and pow(BETA, 3, P) == 1
    assert R.pt_mul(R.G, LAMBDA) == (BETA * R.GX % P, R.GY)
    for k in sv + [LAMBDA - 1, LAMBDA + 1]:
        out = (ctypes.c_uint32 * 78)()
        H.lib().hs_glv(H.limbs(k), out)
        m1, n1, m2, n2 = H.val(out[0:5]), out[5], H.val(out[6:11]), out[11]
        k1, k2 = (-m1 if n1 else m1), (-m2 if n2 else m2)
        assert (k1 + k2 * LAMBDA - k) % N == 0
        assert m1 < 2**129 and m2 < 2**129
        d1 = [ctypes.c_int32(out[12 + i]).value for i in range(33)]
        d2 = [ctypes.c_int32(out[45 + i]).value for i in range(33)]
        assert sum(d * 16 ** (32 - i) for i, d in enumerate(d1)) == m1
        assert sum(d * 16 ** (32 - i) for i, d in enumerate(d2)) == m2
        assert max(map(abs, d1 + d2)) <= 8


def test_sha256_stream():
    rnd = random.Random(3)
    for n in (0, 1, 55, 56, 57, 63, 64, 65, 99, 119, 120, 198, 255):
        m = bytes(rnd.randrange(256) for _ in range(n))
        out = (ctypes.c_uint8 * 32)()
        H.lib().hs_sha256(m, n, out)
        assert bytes(out) == hashlib.sha256(m).digest()


def test_map_to_curve_and_scalar_muls():
    rnd = random.Random(4)
    for u in [0, 1, 2, P - 1, 5] + [rnd.randrange(P) for _ in range(12)]:
        out = (ctypes.c_uint8 * 64)()
        H.lib().hs_map_to_curve(H.limbs(u), out)
        assert bytes(out) == _pt64(R.iso_map(R.map_to_curve_sswu(u)))
    for k in [1, 2, 3, N - 1, N - 2, 2**128, 255, 256, 4095, 4096, LAMBDA, N - LAMBDA] + [rnd.randrange(1, N) for _ in range(6)]:
        for w in (8, 11):
            out = (ctypes.c_uint8 * 64)()
            H.lib().hs_fb_mul(k.to_bytes(32, "big"), w, out)
            assert bytes(out) == c_oracle.mul_g(k)
        base = c_oracle.mul_g(rnd.randrange(1, N))
        out = (ctypes.c_uint8 * 64)()
        H.lib().hs_vb_mul(base, k.to_bytes(32, "big"), out)
        assert bytes(out) == c_oracle.mul(base, k)
    out = (ctypes.c_uint8 * 64)()
    H.lib().hs_vb_mul(_pt64(R.G), (0).to_bytes(32, "big"), out)
    assert bytes(out) == bytes(64)


def test_signed_comb_scalar_multiplication():
    """mul.cuh comb_*: k * P through the signed comb the signer uses for h^r and h^sk (teeth shared between scalars);
    odd and even magnitudes, zero, the GLV edge values, and scalars whose halves are tiny or maximal."""
    rnd = random.Random(12)
    ks = [0, 1, 2, 3, 4, N - 1, N - 2, 2**32, 2**33, 2**33 - 1, 2**66, 2**99 + 1, 2**128, 2**131 % N, LAMBDA, LAMBDA + 1, N - LAMBDA,
          (LAMBDA * 2) % N, 0xFFFFFFFF, 2**255 % N] + [rnd.randrange(N) for _ in range(40)]
    for t, k in enumerate(ks):
        base = c_oracle.mul_g(rnd.randrange(1, N)) if t % 3 else _pt64(R.G)
        out = (ctypes.c_uint8 * 64)()
        H.lib().hs_comb_mul(base, k.to_bytes(32, "big"), out)
        assert bytes(out) == (c_oracle.mul(base, k) if k else bytes(64)), hex(k)


def test_team_shares_add_up():
    """The per-lane shares of the small-batch kernels (stages_team.cuh): the GLV halves of the windowed ladder and of the
    signed comb, and the two window ranges of the generator walk, summed, are k * P."""
    rnd = random.Random(31)
    ks = [0, 1, 2, 3, N - 1, N - 2, 2**128, 2**128 - 1, LAMBDA, LAMBDA + 1, N - LAMBDA, (LAMBDA * 2) % N, 0xFFFFFFFF, 2**255 % N]
    ks += [rnd.randrange(N) for _ in range(16)]
    for t, k in enumerate(ks):
        base = c_oracle.mul_g(rnd.randrange(1, N)) if t % 3 else _pt64(R.G)
        want = c_oracle.mul(base, k) if k else bytes(64)
        for fn in (H.lib().hs_vb_mul_halves, H.lib().hs_comb_mul_halves):
            out = (ctypes.c_uint8 * 64)()
            fn(base, k.to_bytes(32, "big"), out)
            assert bytes(out) == want, hex(k)
        for w in (8, 11):
            out = (ctypes.c_uint8 * 64)()
            H.lib().hs_fb_mul_split(k.to_bytes(32, "big"), w, out)
            assert bytes(out) == (c_oracle.mul_g(k) if k else bytes(64)), hex(k)


def test_h2c_pipeline(golden):
    msgs = [b"abc", b"", bytes(golden["h2c_preimage62"]["preimage"])] + [bytes([i]) * i for i in (1, 54, 55, 56, 63, 64, 65, 127, 128, 200)]
    out = H.h2c_batch(msgs)
    assert bytes(out[0]).hex() == golden["h2c_abc"]["x"] + golden["h2c_abc"]["y"]
    assert np.array_equal(out, c_oracle.h2c_batch(msgs))


def _check_sign_verify(ver, msgs, sks, rs, golden=None):
    skb = b"".join(x.to_bytes(32, "big") for x in sks); rb = b"".join(x.to_bytes(32, "big") for x in rs)
    want = c_oracle.sign_batch(ver, msgs, skb, rb, threads=2)
    for comb in (2, 1, 0):   # the shipped form (comb table kernel + ladder kernel), the comb as one kernel, the windowed ladder (-DPLUME_SIGN_WINDOWED)
        got = H.sign_batch(ver, msgs, skb, rb, gw=8, binv_threads=5, comb=comb)
        for k in ("status", "pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r"):
            assert np.array_equal(got[k], want[k]), (k, comb)
    return got


def test_sign_verify_pipeline(golden):
    k = golden["sign_kat"]
    msg = k["message_ascii"].encode()
    for ver in (1, 2):
        o = _check_sign_verify(ver, [msg], [int(k["sk"]["hex"], 16)], [int(k["r"]["hex"], 16)])
        assert bytes(o["c"][0]).hex() == k["v%d_c" % ver]["hex"] and bytes(o["s"][0]).hex() == k["v%d_s" % ver]["hex"]
    rnd = random.Random(6)
    n = 20
    msgs = [bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 29, 32, 55, 56, 63, 64, 65, 100, 200]))) for _ in range(n)]
    sks = [rnd.randrange(1, N) for _ in range(n)]
    rs = [rnd.randrange(1, N) for _ in range(n)]
    sks[3] = 0; rs[4] = N; sks[5] = N - 1; rs[5] = 1; sks[6] = 1; rs[6] = N - 1
    for ver in (1, 2):
        o = _check_sign_verify(ver, msgs, sks, rs)
        good = [i for i in range(n) if o["status"][i] == 0]
        assert len(good) == n - 2
        sel = lambda key: np.ascontiguousarray(o[key][good])
        gm = [msgs[i] for i in good]
        ok = H.verify_batch(ver, gm, sel("pk"), sel("nullifier"), sel("c"), sel("s"), sel("r_point"), sel("hashed_to_curve_r"))
        assert ok.all()
        # negative tests the reference lacks: one flipped bit per item, every field in turn
        fields = ["pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r"]
        tam = {f: sel(f).copy() for f in fields}
        for j in range(len(good)):
            f = fields[j % 6]
            tam[f][j, rnd.randrange(tam[f].shape[1])] ^= 1 << rnd.randrange(8)
        ok = H.verify_batch(ver, gm, tam["pk"], tam["nullifier"], tam["c"], tam["s"], tam["r_point"], tam["hashed_to_curve_r"])
        want = c_oracle.verify_batch(ver, gm, tam["pk"], tam["nullifier"], tam["c"], tam["s"], tam["r_point"], tam["hashed_to_curve_r"])
        assert np.array_equal(ok, want)
        if ver == 1:
            assert not ok.any()


def test_verify_identity_points():
    """AffinePoint can be the identity and verify() does not reject it: the device code
    follows the same path as the oracle (one-byte encodings, identity operands in the group law)."""
    rnd = random.Random(8)
    msg = b"identity"
    sk, r = rnd.randrange(1, N), rnd.randrange(1, N)
    skb, rb = sk.to_bytes(32, "big"), r.to_bytes(32, "big")
    for ver in (1, 2):
        o = c_oracle.sign_batch(ver, [msg], skb, rb)
        z = np.zeros((1, 64), dtype=np.uint8)
        for pk, nul, rp, hr in ((z, o["nullifier"], o["r_point"], o["hashed_to_curve_r"]), (o["pk"], z, o["r_point"], o["hashed_to_curve_r"]),
                                (o["pk"], o["nullifier"], z, z), (z, z, z, z)):
            got = H.verify_batch(ver, [msg], pk, nul, o["c"], o["s"], rp, hr)
            want = c_oracle.verify_batch(ver, [msg], pk, nul, o["c"], o["s"], rp, hr)
            assert np.array_equal(got, want)
    # a forged V2 "signature" with pk = nullifier = identity: R' = s*G, z' = s*h, c = H(00 || enc || enc) -- accepted by the
    # reference's verify as written (lib.rs:93-145 never rejects identity inputs); both implementations agree
    s = rnd.randrange(1, N)
    h = R.hash_to_curve_bytes(msg + b"\x00")
    c = int.from_bytes(R.c_sha256_vec_signal([None, R.pt_mul(R.G, s), R.pt_mul(h, s)]), "big") % N
    z = np.zeros((1, 64), dtype=np.uint8)
    cb = np.frombuffer(c.to_bytes(32, "big"), dtype=np.uint8).reshape(1, 32)
    sb = np.frombuffer(s.to_bytes(32, "big"), dtype=np.uint8).reshape(1, 32)
    assert R.verify(2, msg, None, None, c, s)
    assert c_oracle.verify_batch(2, [msg], z, z, cb, sb)[0] == 1
    assert H.verify_batch(2, [msg], z, z, cb, sb, z, z)[0] == 1


def test_sec1_compressed_points():
    """33-byte slots <-> affine points, device code (host-sim) vs both oracles."""
    rnd = random.Random(9)
    slots, want = [], []
    pts = [R.pt_mul(R.G, rnd.randrange(1, N)) for _ in range(12)] + [None]
    for p in pts:
        slots.append(R.compress33(p))
    for _ in range(20):       # random x: about half are not on the curve
        slots.append(bytes([rnd.choice([2, 3])]) + rnd.randrange(P).to_bytes(32, "big"))
    slots.append(b"\x02" + P.to_bytes(32, "big"))                      # x = p: non-canonical
    slots.append(b"\x03" + (2**256 - 1).to_bytes(32, "big"))
    slots.append(b"\x04" + slots[0][1:])                               # bad prefix
    slots.append(b"\x00" + b"\x01" + bytes(31))                        # identity prefix with junk
    n = len(slots)
    blob = np.frombuffer(b"".join(slots), dtype=np.uint8).copy()
    out64 = np.zeros((n, 64), dtype=np.uint8); ok = np.zeros(n, dtype=np.uint8); back = np.zeros((n, 33), dtype=np.uint8)
    H.lib().hs_sec1_roundtrip(n, H._p(blob), H._p(out64), H._p(ok), H._p(back))
    for i, b in enumerate(slots):
        p, good = R.decompress33(b)
        o64, g2 = c_oracle.decompress33(b)
        assert bool(ok[i]) == bool(good) == bool(g2), i
        assert bytes(out64[i]) == _pt64(p) == o64, i
        if good:
            assert bytes(back[i]) == b == c_oracle.compress33(o64)

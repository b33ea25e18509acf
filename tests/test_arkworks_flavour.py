"""the arkworks twin's semantics (rust-arkworks/src/lib.rs:229-278 sign_with_r, tests.rs:28-78
verify_non_zk).  CPU part: the Python restatement pinned to the reference's own vectors
(rust-arkworks/src/tests.rs:267-300) and the host-sim build of the kernel sources against it, including what only
this flavour allows (zero scalars -> identity points hashed as the byte 00, a public key that is an input, c reduced
mod n).  GPU part: the same through the C ABI."""
import random

import numpy as np
import pytest

import _hostsim as H
import plume_ref as R

N = R.N


def _pt64(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def _b32(x):
    return x.to_bytes(32, "big")


def _cases(seed, count):
    """(msg, pk, sk, r): ordinary keypairs, a pk that does not belong to sk, zero sk / zero r, ragged messages."""
    rnd = random.Random(seed)
    out = []
    for i in range(count):
        sk, r = rnd.randrange(1, N), rnd.randrange(1, N)
        pk = R.pt_mul(R.G, sk)
        msg = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 29, 32, 55, 56, 64, 100])))
        if i % 7 == 3:
            pk = R.pt_mul(R.G, rnd.randrange(1, N))      # the keypair is taken as given: pk need not be sk * G
        if i % 7 == 4:
            sk = 0
        if i % 7 == 5:
            r = 0
        if i % 11 == 6:
            sk, r = 0, 0
        out.append((msg, pk, sk, r))
    return out


def _oracle_sign(version, cases):
    return [R.ark_sign_with_r(version, m, pk, sk, r) for m, pk, sk, r in cases]


def _check_sign(version, cases, got):
    for i, ((st, want), (m, pk, sk, r)) in enumerate(zip(_oracle_sign(version, cases), cases)):
        assert int(got["status"][i]) == st, i
        if st:
            continue
        assert bytes(got["nullifier"][i]) == _pt64(want["nullifier"]), i
        assert bytes(got["digest_private"][i]) == _b32(want["digest_private"]), i
        assert bytes(got["s"][i]) == _b32(want["s"]), i
        assert bytes(got["r_point"][i]) == _pt64(want["r_point"]), i
        assert bytes(got["hashed_to_curve_r"][i]) == _pt64(want["hashed_to_curve_r"]), i


def _verify_inputs(version, cases):
    """valid signatures of the cases plus tampered copies; returns arrays and the oracle's verdicts"""
    rows, want = [], []
    rnd = random.Random(99)
    for (m, pk, sk, r) in cases:
        st, o = R.ark_sign_with_r(version, m, pk, sk, r)
        assert st == 0
        base = dict(msg=m, pk=pk, nul=o["nullifier"], c=o["digest_private"], s=o["s"], rp=o["r_point"], z=o["hashed_to_curve_r"])
        rows.append(base)
        t = dict(base)
        which = rnd.choice(["s", "c", "nul", "rp", "z", "msg", "pk"])
        if which in ("s", "c"):
            t[which] = (t[which] + 1) % N
        elif which == "msg":
            t["msg"] = t["msg"] + b"!"
        else:
            t[which] = R.pt_add(t[which], R.G)
        rows.append(t)
    for x in rows:
        want.append(R.ark_verify_non_zk(version, x["msg"], x["pk"], x["nul"], x["c"], x["s"], x["rp"], x["z"]))
    arr = lambda f, w: np.frombuffer(b"".join(f(x) for x in rows), dtype=np.uint8).reshape(len(rows), w)
    return ([x["msg"] for x in rows], arr(lambda x: _pt64(x["pk"]), 64), arr(lambda x: _pt64(x["nul"]), 64),
            arr(lambda x: _b32(x["c"]), 32), arr(lambda x: _b32(x["s"]), 32), arr(lambda x: _pt64(x["rp"]), 64),
            arr(lambda x: _pt64(x["z"]), 64)), want


def test_oracle_pinned_to_arkworks_vectors(golden):
    k = golden["sign_kat"]
    msg, sk, r = k["message_ascii"].encode(), int(k["sk"]["hex"], 16), int(k["r"]["hex"], 16)
    pk = R.pt_mul(R.G, sk)
    for ver in (1, 2):
        st, o = R.ark_sign_with_r(ver, msg, pk, sk, r)
        assert st == 0
        assert "%064x" % o["digest_private"] == golden["arkworks_c_s"]["v%d_c" % ver]["hex"]   # tests.rs:281-299
        assert "%064x" % o["s"] == golden["arkworks_c_s"]["v%d_s" % ver]["hex"]
        inter = golden["intermediates"]                                                       # tests.rs:189-263
        assert _pt64(o["r_point"]).hex() == inter["g_r"]["x"] + inter["g_r"]["y"]
        assert _pt64(o["hashed_to_curve_r"]).hex() == inter["h_r"]["x"] + inter["h_r"]["y"]
        assert _pt64(o["nullifier"]).hex() == inter["h_sk"]["x"] + inter["h_sk"]["y"]
        # test_sign_and_verify (tests.rs:139-167): what sign produces, verify_non_zk accepts
        assert R.ark_verify_non_zk(ver, msg, pk, o["nullifier"], o["digest_private"], o["s"], o["r_point"], o["hashed_to_curve_r"]) is True
        assert R.ark_verify_non_zk(ver, msg, pk, o["nullifier"], o["digest_private"], (o["s"] + 1) % N, o["r_point"], o["hashed_to_curve_r"]) is False
    assert R.ark_sign_with_r(1, msg, R.INF, sk, r)[0] == R.ST_BAD_PK
    # zero scalars are legal Fr values: identity points, hashed as the single byte 00 (lib.rs:112-118)
    st, o = R.ark_sign_with_r(2, msg, pk, sk, 0)
    assert st == 0 and o["r_point"] is R.INF and o["hashed_to_curve_r"] is R.INF and o["s"] == sk * o["digest_private"] % N


def test_hostsim_ark_sign_matches_oracle():
    cases = _cases(5, 22)
    for ver in (1, 2):
        got = H.ark_sign_batch(ver, [c[0] for c in cases], b"".join(_pt64(c[1]) for c in cases), b"".join(_b32(c[2]) for c in cases),
                               b"".join(_b32(c[3]) for c in cases))
        _check_sign(ver, cases, got)
    # rejected inputs: scalar >= n, identity / off-curve public key
    bad = [(b"m", R.G, N, 1), (b"m", R.G, 1, N + 5), (b"m", R.INF, 1, 1)]
    got = H.ark_sign_batch(1, [c[0] for c in bad], b"".join(_pt64(c[1]) for c in bad[:2]) + bytes(64),
                           b"".join(_b32(c[2]) for c in bad), b"".join(_b32(c[3]) for c in bad))
    assert list(got["status"]) == [R.ST_BAD_SK, R.ST_BAD_R, R.ST_BAD_PK]
    off = bytearray(_pt64(R.G)); off[63] ^= 1
    got = H.ark_sign_batch(1, [b"m"], bytes(off), _b32(1), _b32(1))
    assert list(got["status"]) == [R.ST_BAD_PK]


def test_hostsim_ark_verify_matches_oracle():
    cases = _cases(6, 12)
    for ver in (1, 2):
        (msgs, pk, nul, c, s, rp, z), want = _verify_inputs(ver, cases)
        ok = H.ark_verify_batch(ver, msgs, pk, nul, c, s, rp, z)
        assert [bool(v) for v in ok] == [bool(w) for w in want]
        # a keypair whose pk is not sk * G signs fine but does not verify; the consistent ones do
        assert 4 <= sum(bool(w) for w in want) < len(cases)


@pytest.mark.gpu
def test_gpu_ark_sign_and_verify_match_oracle(gpu_ctx):
    import plume_b200.arkworks as ark
    cases = _cases(7, 96)
    for ver in (1, 2):
        got = gpu_ctx.ark_sign_batch(ver, [c[0] for c in cases], b"".join(_pt64(c[1]) for c in cases),
                                     b"".join(_b32(c[2]) for c in cases), b"".join(_b32(c[3]) for c in cases))
        _check_sign(ver, cases, got)
        (msgs, pk, nul, c, s, rp, z), want = _verify_inputs(ver, cases[:40])
        ok = gpu_ctx.ark_verify_batch(ver, msgs, pk, nul, c, s, rp, z)
        assert [bool(v) for v in ok] == [bool(w) for w in want]
    # the reference-shaped calls (batches of one) on the reference's fixed vector
    sk = 0x519b423d715f8b581f4fa8ee59f4771a5b44c8130b4e3eacca54a56dda72b464
    r = 0x93b9323b629f251b8f3fc2dd11f4672c5544e8230d493eceea98a90bda789808
    msg = b"An example app message string"
    pk = R.pt_mul(R.G, sk)
    pub, priv = ark.sign_with_r((pk, sk), msg, r, ark.PlumeVersion.V1, ctx=gpu_ctx)
    assert "%064x" % priv.digest_private == "c6a7fc2c926ddbaf20731a479fb6566f2daa5514baae5223fe3b32edbce83254"
    assert "%064x" % pub.s == "e69f027d84cb6fe5f761e333d12e975fb190d163e8ea132d7de0bd6079ba28ca"
    assert ark.verify_non_zk((pub, priv), pk, msg, ark.PlumeVersion.V1, ctx=gpu_ctx)
    pub.s = (pub.s + 1) % N
    assert not ark.verify_non_zk((pub, priv), pk, msg, ark.PlumeVersion.V1, ctx=gpu_ctx)
    with pytest.raises(ark.HashToCurveError):
        ark.sign_with_r((None, sk), msg, r, ark.PlumeVersion.V2, ctx=gpu_ctx)


@pytest.mark.gpu
def test_gpu_ark_large_batch_roundtrip(gpu_ctx):
    """2^16 items: every arkworks-flavour signature verifies, and tampering s makes all fail."""
    n = 1 << 16
    rng = np.random.default_rng(11)
    msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
    r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
    k = gpu_ctx.sign_batch(2, msgs, sk, r)                 # k256 flavour gives the matching public keys
    for ver in (1, 2):
        o = gpu_ctx.ark_sign_batch(ver, msgs, k["pk"], sk, r)
        assert (o["status"] == 0).all()
        if ver == 2:   # same nullifier / s as the k256 flavour whenever c < n (always, for random data)
            assert np.array_equal(o["nullifier"], k["nullifier"]) and np.array_equal(o["s"], k["s"]) and np.array_equal(o["digest_private"], k["c"])
        ok = gpu_ctx.ark_verify_batch(ver, msgs, k["pk"], o["nullifier"], o["digest_private"], o["s"], o["r_point"], o["hashed_to_curve_r"])
        assert ok.all()
        bad = o["s"].copy(); bad[:, 31] ^= 1
        assert not gpu_ctx.ark_verify_batch(ver, msgs, k["pk"], o["nullifier"], o["digest_private"], bad, o["r_point"], o["hashed_to_curve_r"]).any()

"""Pins both oracles (oracle/plume_ref.py, oracle/plume_oracle.c) to every golden vector the
reference's own tests hold for the PLUME hot path (tests/golden/reference_vectors.json; sources
cited inside), and to each other on random inputs."""
import hashlib
import random

import numpy as np

import c_oracle
import plume_ref as R


def _pt64(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def test_sign_kat_and_intermediates(golden):
    # rust-k256/tests/signing.rs:9-21,48-64 and rust-arkworks/src/tests.rs:180-299
    k = golden["sign_kat"]
    msg = k["message_ascii"].encode()
    sk = bytes.fromhex(k["sk"]["hex"]); r = bytes.fromhex(k["r"]["hex"])
    inter = golden["intermediates"]
    for ver in (1, 2):
        st, ref = R.sign(ver, msg, int.from_bytes(sk, "big"), int.from_bytes(r, "big"))
        o = c_oracle.sign_batch(ver, [msg], sk, r)
        assert st == 0 and o["status"][0] == 0
        for cs in ("c", "s"):
            want = k["v%d_%s" % (ver, cs)]["hex"]
            assert want == golden["arkworks_c_s"]["v%d_%s" % (ver, cs)]["hex"]
            assert "%064x" % ref[cs] == want
            assert bytes(o[cs][0]).hex() == want
        for name, key in (("pk", "pk"), ("g_r", "r_point"), ("h_r", "hashed_to_curve_r"), ("h_sk", "nullifier")):
            want = inter[name]["x"] + inter[name]["y"]
            assert _pt64(ref[key]).hex() == want
            assert bytes(o[key][0]).hex() == want
        assert _pt64(ref["h"]).hex() == inter["h"]["x"] + inter["h"]["y"]


def test_verify_kat(golden):
    # rust-k256/tests/verification.rs:25-107: the assembled V1 and V2 signatures verify
    k = golden["sign_kat"]
    msg = k["message_ascii"].encode()
    sk = bytes.fromhex(k["sk"]["hex"]); r = bytes.fromhex(k["r"]["hex"])
    for ver in (1, 2):
        o = c_oracle.sign_batch(ver, [msg], sk, r)
        if ver == 1:
            assert bytes(o["c"][0]).hex() == golden["verify_kat_c_v1"]["hex"]
        ok = c_oracle.verify_batch(ver, [msg], o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
        assert ok[0] == 1
        st, ref = R.sign(ver, msg, int.from_bytes(sk, "big"), int.from_bytes(r, "big"))
        assert R.verify(ver, msg, ref["pk"], ref["nullifier"], ref["c"], ref["s"],
                        ref["r_point"] if ver == 1 else None, ref["hashed_to_curve_r"] if ver == 1 else None)


def test_h2c_kats(golden):
    # "abc": rust-k256/tests/verification.rs:283-294; empty: rust-arkworks/src/secp256k1/tests.rs:87-156;
    # 62-byte preimage: circuits/circom/test/javascript/test/hashToCurve.test.ts:5-19
    out = c_oracle.h2c_batch([b"abc", b"", bytes(golden["h2c_preimage62"]["preimage"])])
    assert bytes(out[0]).hex() == golden["h2c_abc"]["x"] + golden["h2c_abc"]["y"]
    assert _pt64(R.hash_to_curve_bytes(b"abc")).hex() == golden["h2c_abc"]["x"] + golden["h2c_abc"]["y"]
    e = golden["h2c_empty"]
    assert int.from_bytes(bytes(out[1][:32]), "big") == int(e["px_dec"])
    assert int.from_bytes(bytes(out[1][32:]), "big") == int(e["py_dec"])
    u = R.hash_to_field2(b"")
    assert u[0] == int(e["u0_dec"]) and "%064x" % u[0] == e["rfc_comment"]["u[0]"] and "%064x" % u[1] == e["rfc_comment"]["u[1]"]
    for i, q in enumerate((R.iso_map(R.map_to_curve_sswu(u[0])), R.iso_map(R.map_to_curve_sswu(u[1])))):
        assert "%064x" % q[0] == e["rfc_comment"]["Q%d.x" % i] and "%064x" % q[1] == e["rfc_comment"]["Q%d.y" % i]
    pp = golden["h2c_preimage62"]
    assert bytes(out[2]).hex() == pp["x"] + pp["y"]
    # the 62-byte preimage is message || enc(pk) of the sign KAT
    k = golden["sign_kat"]
    assert bytes(pp["preimage"])[:29] == k["message_ascii"].encode()


def test_sec1_vectors(golden):
    # rust-arkworks/src/tests/test_vectors.rs:1-504 (k*G, k = 0..99) and rust-k256/src/lib.rs:177-183
    for t in golden["sec1_kG"]["vectors"]:
        p64 = c_oracle.mul_g(t["k"]) if t["k"] else bytes(64)
        assert c_oracle.encode_pt(p64).hex() == t["compressed"]
        ref = R.pt_mul(R.G, t["k"])
        assert R.encode_pt(ref).hex() == t["compressed"]
        assert R.encode_pt_uncompressed(ref).hex() == t["uncompressed"]
        if t["k"]:
            assert "04" + p64.hex() == t["uncompressed"]
    assert c_oracle.encode_pt(_pt64(R.G)).hex() == golden["encode_pt_G"]["hex"]
    assert golden["dst"]["ascii"].encode() == R.DST


def test_scalar_roundtrip(golden):
    # rust-k256/src/lib.rs:199-208: 32 BE bytes -> scalar -> bytes is the identity for values < n
    v = bytes.fromhex(golden["scalar_roundtrip"]["hex"])
    assert int.from_bytes(v, "big") < R.N
    assert (int.from_bytes(v, "big") % R.N).to_bytes(32, "big") == v


def test_sha_and_xmd_against_hashlib():
    rnd = random.Random(3)
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 200):
        m = bytes(rnd.randrange(256) for _ in range(n))
        assert c_oracle.sha256(m) == hashlib.sha256(m).digest()
        assert c_oracle.expand_message_xmd(m, 96) == R.expand_message_xmd(m, R.DST, 96)


def test_c_oracle_vs_python_random():
    rnd = random.Random(11)
    n = 40
    msgs = [bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 31, 32, 33, 55, 56, 64, 100, 300]))) for _ in range(n)]
    sks = [rnd.randrange(1, R.N) for _ in range(n)]
    rs = [rnd.randrange(1, R.N) for _ in range(n)]
    sks[0] = 0; rs[1] = 0; sks[2] = R.N; rs[3] = R.N + 5; sks[4] = 1; rs[4] = R.N - 1
    skb = b"".join(x.to_bytes(32, "big") for x in sks); rb = b"".join(x.to_bytes(32, "big") for x in rs)
    for ver in (1, 2):
        o = c_oracle.sign_batch(ver, msgs, skb, rb, threads=4)
        for i in range(n):
            st, ref = R.sign(ver, msgs[i], sks[i], rs[i])
            assert o["status"][i] == st
            if st:
                assert not o["pk"][i].any() and not o["s"][i].any()
                continue
            for key in ("pk", "nullifier", "r_point", "hashed_to_curve_r"):
                assert bytes(o[key][i]) == _pt64(ref[key])
            assert int.from_bytes(bytes(o["c"][i]), "big") == ref["c"]
            assert int.from_bytes(bytes(o["s"][i]), "big") == ref["s"]
        good = [i for i in range(n) if o["status"][i] == 0]
        sel = lambda k: np.ascontiguousarray(o[k][good])
        ok = c_oracle.verify_batch(ver, [msgs[i] for i in good], sel("pk"), sel("nullifier"), sel("c"), sel("s"),
                                   sel("r_point"), sel("hashed_to_curve_r"), threads=4)
        assert ok.all()


def test_verify_edge_semantics():
    """Edge cases no reference test pins (SURVEY.md 8c 'parity unpinned'): the two oracles agree."""
    rnd = random.Random(5)
    msg = b"edge"
    sk, r = rnd.randrange(1, R.N), rnd.randrange(1, R.N)
    st, ref = R.sign(2, msg, sk, r)
    pk, nul, c, s = ref["pk"], ref["nullifier"], ref["c"], ref["s"]

    def both(ver, pk, nul, c, s, rp=None, hr=None):
        a = R.verify(ver, msg, pk, nul, c, s, rp, hr)
        b = c_oracle.verify_batch(ver, [msg], _pt64(pk), _pt64(nul), c.to_bytes(32, "big"), s.to_bytes(32, "big"),
                                  _pt64(rp), _pt64(hr))[0]
        assert bool(a) == bool(b)
        return bool(a)

    assert both(2, pk, nul, c, s)
    assert not both(2, None, nul, c, s)           # identity pk: encodes to one byte, still well-formed
    assert not both(2, pk, None, c, s)            # identity nullifier
    assert not both(2, pk, nul, (c + 1) % R.N or 1, s)
    # malformed inputs are rejected by the C oracle (the reference's types cannot hold them)
    bad = bytearray(_pt64(pk)); bad[63] ^= 1
    assert c_oracle.verify_batch(2, [msg], bytes(bad), _pt64(nul), c.to_bytes(32, "big"), s.to_bytes(32, "big"))[0] == 0
    assert c_oracle.verify_batch(2, [msg], _pt64(pk), _pt64(nul), bytes(32), s.to_bytes(32, "big"))[0] == 0
    assert c_oracle.verify_batch(2, [msg], _pt64(pk), _pt64(nul), R.N.to_bytes(32, "big"), s.to_bytes(32, "big"))[0] == 0

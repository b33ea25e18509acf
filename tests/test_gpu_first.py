"""First GPU parity checks: golden vectors of the reference and random small batches against the
pure-Python oracle, all through the C ABI (ctypes -> libplume_b200.so)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pt(b):
    import plume_b200
    return plume_b200.point_from_bytes(b)


def test_sign_kat(gpu_ctx, golden):
    k = golden["sign_kat"]
    msg = k["message_ascii"].encode()
    sk = bytes.fromhex(k["sk"]["hex"]); r = bytes.fromhex(k["r"]["hex"])
    inter = golden["intermediates"]
    for ver in (1, 2):
        o = gpu_ctx.sign_batch(ver, [msg], sk, r)
        assert o["status"][0] == 0
        assert bytes(o["c"][0]).hex() == k["v%d_c" % ver]["hex"]
        assert bytes(o["s"][0]).hex() == k["v%d_s" % ver]["hex"]
        for name, key in (("pk", "pk"), ("g_r", "r_point"), ("h_r", "hashed_to_curve_r"), ("h_sk", "nullifier")):
            assert bytes(o[key][0]).hex() == inter[name]["x"] + inter[name]["y"], name
        ok = gpu_ctx.verify_batch(ver, [msg], o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
        assert ok[0] == 1


def test_h2c_kats(gpu_ctx, golden):
    out = gpu_ctx.hash_to_curve_batch([b"abc", b"", bytes(golden["h2c_preimage62"]["preimage"])])
    assert bytes(out[0]).hex() == golden["h2c_abc"]["x"] + golden["h2c_abc"]["y"]
    e = golden["h2c_empty"]
    assert int.from_bytes(bytes(out[1][:32]), "big") == int(e["px_dec"])
    assert int.from_bytes(bytes(out[1][32:]), "big") == int(e["py_dec"])
    assert bytes(out[2]).hex() == golden["h2c_preimage62"]["x"] + golden["h2c_preimage62"]["y"]


def test_random_vs_python_oracle(gpu_ctx):
    import plume_ref as R
    rnd = random.Random(7)
    n = 48
    msgs = [bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 29, 32, 55, 56, 63, 64, 65, 100, 200]))) for _ in range(n)]
    sks = [rnd.randrange(1, R.N) for _ in range(n)]
    rs = [rnd.randrange(1, R.N) for _ in range(n)]
    sks[3] = 0; rs[4] = R.N; sks[5] = R.N - 1; rs[5] = 1
    skb = b"".join(x.to_bytes(32, "big") for x in sks); rb = b"".join(x.to_bytes(32, "big") for x in rs)
    for ver in (1, 2):
        o = gpu_ctx.sign_batch(ver, msgs, skb, rb)
        for i in range(n):
            st, ref = R.sign(ver, msgs[i], sks[i], rs[i])
            assert o["status"][i] == st, i
            if st:
                assert not o["pk"][i].any() and not o["c"][i].any()
                continue
            assert _pt(o["pk"][i]) == ref["pk"]
            assert _pt(o["nullifier"][i]) == ref["nullifier"]
            assert _pt(o["r_point"][i]) == ref["r_point"]
            assert _pt(o["hashed_to_curve_r"][i]) == ref["hashed_to_curve_r"]
            assert int.from_bytes(bytes(o["c"][i]), "big") == ref["c"]
            assert int.from_bytes(bytes(o["s"][i]), "big") == ref["s"]
        good = [i for i in range(n) if o["status"][i] == 0]
        sel = lambda k: np.ascontiguousarray(o[k][good])
        ok = gpu_ctx.verify_batch(ver, [msgs[i] for i in good], sel("pk"), sel("nullifier"), sel("c"), sel("s"),
                                  sel("r_point"), sel("hashed_to_curve_r"))
        assert ok.all()
        # flip one bit somewhere in every field in turn
        fields = ["pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r"]
        tam = {k: sel(k).copy() for k in fields}
        for j in range(len(good)):
            f = fields[j % 6]
            tam[f][j, rnd.randrange(tam[f].shape[1])] ^= 1 << rnd.randrange(8)
        ok = gpu_ctx.verify_batch(ver, [msgs[i] for i in good], tam["pk"], tam["nullifier"], tam["c"], tam["s"],
                                  tam["r_point"], tam["hashed_to_curve_r"])
        for j, i in enumerate(good):
            pts = [_pt(tam[k][j]) for k in ("pk", "nullifier", "r_point", "hashed_to_curve_r")]
            use = pts if ver == 1 else pts[:2]
            valid = all(p is None or (p[0] < R.P and p[1] < R.P and R.on_curve(p)) for p in use)
            c = int.from_bytes(bytes(tam["c"][j]), "big"); s = int.from_bytes(bytes(tam["s"][j]), "big")
            valid = valid and 1 <= c < R.N and 1 <= s < R.N
            exp = valid and R.verify(ver, msgs[i], pts[0], pts[1], c, s, pts[2] if ver == 1 else None, pts[3] if ver == 1 else None)
            assert bool(ok[j]) == bool(exp), (ver, j, fields[j % 6])

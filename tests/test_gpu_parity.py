"""GPU parity tests proper: the CUDA path through the C ABI (ctypes -> libplume_b200.so) against the C oracle
on identical inputs, bit-exact, plus size-independent properties at BASELINE.json's full batch size.

Covers the cases the reference tests (fixed vector, empty / 3-byte / 29-byte messages) and the ones it leaves
unpinned: ragged and long messages, out-of-range scalars, identity / off-curve / non-canonical
points, c or s outside [1, n-1], chunk boundaries of the host API, device-pointer vs host-pointer entry points."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
P = 2**256 - 2**32 - 977
FIELDS = ("pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r")


def _rand_scalars(rng, n):
    a = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    a[:, 0] &= 0x7F          # < 2^255 < n, and non-zero with overwhelming probability
    return a


def _sign_both(ctx, ver, msgs, sk, r):
    import c_oracle
    got = ctx.sign_batch(ver, msgs, sk, r)
    want = c_oracle.sign_batch(ver, msgs, sk, r, threads=os.cpu_count() or 1)
    for k in ("status",) + FIELDS:
        assert np.array_equal(got[k], want[k]), "sign v%d: %s differs" % (ver, k)
    return got


def _verify_both(ctx, ver, msgs, o, expect=None):
    import c_oracle
    got = ctx.verify_batch(ver, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    want = c_oracle.verify_batch(ver, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"],
                                 threads=os.cpu_count() or 1)
    assert np.array_equal(got, want), "verify v%d differs from the oracle" % ver
    if expect is not None:
        assert np.array_equal(got, expect)
    return got


def test_ragged_messages_and_bad_scalars(gpu_ctx):
    rnd = random.Random(21)
    rng = np.random.default_rng(21)
    lens = [0, 1, 2, 3, 29, 31, 32, 33, 54, 55, 56, 57, 63, 64, 65, 118, 119, 120, 127, 128, 129, 200, 255, 256, 1000, 4097]
    msgs = [bytes(rnd.randrange(256) for _ in range(L)) for L in lens * 4]
    n = len(msgs)
    sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
    def put(a, i, v): a[i] = np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8)
    put(sk, 0, 0); put(r, 1, 0); put(sk, 2, N); put(r, 3, N); put(sk, 4, 2**256 - 1); put(r, 5, 2**256 - 1)
    put(sk, 6, 1); put(r, 6, N - 1); put(sk, 7, N - 1); put(r, 7, 1); put(sk, 8, 0); put(r, 8, 0)
    for ver in (1, 2):
        o = _sign_both(gpu_ctx, ver, msgs, sk, r)
        assert list(o["status"][:9]) == [2, 1, 2, 1, 2, 1, 0, 0, 1]
        good = np.flatnonzero(o["status"] == 0)
        sub = {k: np.ascontiguousarray(o[k][good]) for k in FIELDS}
        _verify_both(gpu_ctx, ver, [msgs[i] for i in good], sub, expect=np.ones(len(good), dtype=np.uint8))


def test_fixed_length_records(gpu_ctx):
    rng = np.random.default_rng(22)
    for mlen in (1, 32, 64, 65, 100):
        n = 300
        msgs = rng.integers(0, 256, (n, mlen), dtype=np.uint8)
        sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
        o = _sign_both(gpu_ctx, 1, msgs, sk, r)
        _verify_both(gpu_ctx, 1, msgs, o, expect=np.ones(n, dtype=np.uint8))


def test_tampered_and_malformed_verify_inputs(gpu_ctx):
    rnd = random.Random(23)
    rng = np.random.default_rng(23)
    n = 600
    msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
    for ver in (1, 2):
        o = _sign_both(gpu_ctx, ver, msgs, sk, r)
        t = {k: o[k].copy() for k in FIELDS}
        for i in range(n):
            kind = i % 12
            if kind < 6:                                   # one flipped bit in one field
                f = FIELDS[kind]
                t[f][i, rnd.randrange(t[f].shape[1])] ^= 1 << rnd.randrange(8)
            elif kind == 6:                                # identity pk
                t["pk"][i] = 0
            elif kind == 7:                                # identity nullifier
                t["nullifier"][i] = 0
            elif kind == 8:                                # c = 0 / c >= n
                t["c"][i] = np.frombuffer((0 if i % 24 < 12 else N).to_bytes(32, "big"), dtype=np.uint8)
            elif kind == 9:                                # s >= n
                t["s"][i] = 0xFF
            elif kind == 10:                               # non-canonical x (x + p) when it fits in 256 bits
                x = int.from_bytes(bytes(t["pk"][i][:32]), "big")
                if x + P < 2**256:
                    t["pk"][i][:32] = np.frombuffer((x + P).to_bytes(32, "big"), dtype=np.uint8)
            # kind 11: untouched, must still verify
        got = _verify_both(gpu_ctx, ver, msgs, t)
        assert got[11::12].all()
        assert not got[8::12].any() and not got[9::12].any()
        if ver == 1:
            assert not got[0::12].any() and not got[4::12].any() and not got[5::12].any()
        # all-identity "signatures"
        z = {k: np.zeros_like(o[k]) for k in FIELDS}
        z["c"] = o["c"]; z["s"] = o["s"]
        _verify_both(gpu_ctx, ver, msgs, z)


def test_identity_forgery_agrees_with_oracle(gpu_ctx):
    """pk = nullifier = identity with c = H(00 || enc(s*G) || enc(s*h)) passes the reference's verify as written
    (lib.rs:93-145 never rejects identity inputs); the CUDA path must say the same as the oracle."""
    import c_oracle
    import plume_ref as R
    rnd = random.Random(24)
    msgs, cs, ss = [], [], []
    for _ in range(8):
        m = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 70)))
        s = rnd.randrange(1, N)
        h = R.hash_to_curve_bytes(m + b"\x00")
        c = int.from_bytes(R.c_sha256_vec_signal([None, R.pt_mul(R.G, s), R.pt_mul(h, s)]), "big") % N
        msgs.append(m); cs.append(c.to_bytes(32, "big")); ss.append(s.to_bytes(32, "big"))
    n = len(msgs)
    z = np.zeros((n, 64), dtype=np.uint8)
    c = np.frombuffer(b"".join(cs), dtype=np.uint8).reshape(n, 32); s = np.frombuffer(b"".join(ss), dtype=np.uint8).reshape(n, 32)
    got = gpu_ctx.verify_batch(2, msgs, z, z, c, s)
    want = c_oracle.verify_batch(2, msgs, z, z, c, s)
    assert np.array_equal(got, want) and got.all()


def test_hash_to_curve_batch(gpu_ctx):
    import c_oracle
    rnd = random.Random(25)
    msgs = [bytes(rnd.randrange(256) for _ in range(L)) for L in [0, 1, 3, 29, 32, 62, 63, 64, 65, 66, 127, 128, 129, 500] * 20]
    assert np.array_equal(gpu_ctx.hash_to_curve_batch(msgs), c_oracle.h2c_batch(msgs, threads=os.cpu_count() or 1))
    fixed = np.random.default_rng(25).integers(0, 256, (6000, 65), dtype=np.uint8)
    for n in (1, 3, 33, 4096, 4097, 6000):     # two lanes per item up to 4 096 items, one thread per item above
        assert np.array_equal(gpu_ctx.hash_to_curve_batch(fixed[:n]), c_oracle.h2c_batch(fixed[:n], threads=os.cpu_count() or 1)), n


def test_chunk_boundaries_of_host_api():
    """A context with a tiny chunk size walks both lanes several times; results must not depend on it."""
    import c_oracle
    import plume_b200
    os.environ["PLUME_CHUNK_ITEMS"] = "1000"
    os.environ["PLUME_FIXED_WINDOW"] = "9"      # also exercises a window width that does not divide 256
    try:
        ctx = plume_b200.PlumeContext(0)
    finally:
        del os.environ["PLUME_CHUNK_ITEMS"], os.environ["PLUME_FIXED_WINDOW"]
    try:
        assert ctx.chunk_items == 1000
        rng = np.random.default_rng(26)
        rnd = random.Random(26)
        for n in (1, 999, 1000, 1001, 2000, 4321):
            msgs = [bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 80))) for _ in range(n)]
            sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
            o = _sign_both(ctx, 2, msgs, sk, r)
            _verify_both(ctx, 2, msgs, o, expect=np.ones(n, dtype=np.uint8))
        assert ctx.sign_batch(1, [], np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8))["status"].shape == (0,)
    finally:
        ctx.close()


def test_64k_batch_bit_exact(gpu_ctx):
    """2^16 items, every output byte compared with the CPU oracle (SURVEY.md 8d's sample size)."""
    rng = np.random.default_rng(27)
    n = 1 << 16
    msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
    for ver in (1, 2):
        o = _sign_both(gpu_ctx, ver, msgs, sk, r)
        _verify_both(gpu_ctx, ver, msgs, o, expect=np.ones(n, dtype=np.uint8))


def test_full_size_properties_and_device_api(gpu_ctx):
    """BASELINE size (2^20): sign -> verify round trip is all-true, a tampered stripe is all-false, results are
    deterministic, and the device-pointer entry points produce the same bytes as the host-pointer ones."""
    import torch
    rng = np.random.default_rng(28)
    n = 1 << 20
    msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
    o = gpu_ctx.sign_batch(1, msgs, sk, r)
    assert (o["status"] == 0).all()
    ok = gpu_ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], o["s"], o["r_point"], o["hashed_to_curve_r"])
    assert ok.all()
    o2 = gpu_ctx.sign_batch(1, msgs, sk, r)
    assert all(np.array_equal(o[k], o2[k]) for k in FIELDS)
    bad = o["s"].copy()
    bad[::1024, 31] ^= 1
    ok = gpu_ctx.verify_batch(1, msgs, o["pk"], o["nullifier"], o["c"], bad, o["r_point"], o["hashed_to_curve_r"])
    assert not ok[::1024].any() and ok.sum() == n - len(ok[::1024])
    # device-pointer API on a torch stream, chunk by chunk
    dev = torch.device("cuda", 0)
    D = {k: torch.from_numpy(v).to(dev) for k, v in (("msgs", msgs), ("sk", sk), ("r", r))}
    for k in FIELDS:
        D[k] = torch.empty((n, o[k].shape[1]), dtype=torch.uint8, device=dev)
    D["status"] = torch.empty(n, dtype=torch.uint8, device=dev)
    D["ok"] = torch.empty(n, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    ch = gpu_ctx.chunk_items
    for i0 in range(0, n, ch):
        cn = min(ch, n - i0)
        p = lambda t: t.data_ptr() + i0 * (t.shape[1] if t.dim() > 1 else 1)
        gpu_ctx.sign_batch_device(1, cn, p(D["msgs"]), 0, 32, p(D["sk"]), p(D["r"]), p(D["pk"]), p(D["nullifier"]), p(D["c"]),
                                  p(D["s"]), p(D["r_point"]), p(D["hashed_to_curve_r"]), p(D["status"]), st.cuda_stream)
        gpu_ctx.verify_batch_device(1, cn, p(D["msgs"]), 0, 32, p(D["pk"]), p(D["nullifier"]), p(D["c"]), p(D["s"]),
                                    p(D["r_point"]), p(D["hashed_to_curve_r"]), p(D["ok"]), st.cuda_stream)
    st.synchronize()
    for k in FIELDS:
        assert np.array_equal(D[k].cpu().numpy(), o[k]), k
    assert bool(D["ok"].all().item()) and int(D["status"].sum().item()) == 0


def test_reference_style_api(gpu_ctx, golden):
    """The mirror of the reference's own test (rust-k256/tests/signing.rs:23-64): a mock RNG whose fill_bytes
    returns the fixed nonce, PlumeSignature::sign_v1 / sign_v2, then verify()."""
    import plume_b200
    k = golden["sign_kat"]
    R = bytes.fromhex(k["r"]["hex"])

    class Mock:
        def fill_bytes(self, dest):
            assert len(dest) == len(R) == 32
            dest[:] = R

    sk = plume_b200.SecretKey.from_bytes(bytes.fromhex(k["sk"]["hex"]))
    msg = k["message_ascii"].encode()
    sig = plume_b200.PlumeSignature.sign_v1(sk, msg, Mock(), ctx=gpu_ctx)
    assert "%064x" % sig.c == k["v1_c"]["hex"] and "%064x" % sig.s == k["v1_s"]["hex"]
    assert sig.v1specific is not None and sig.verify()
    sig = plume_b200.PlumeSignature.sign_v2(sk, msg, Mock(), ctx=gpu_ctx)
    assert "%064x" % sig.c == k["v2_c"]["hex"] and "%064x" % sig.s == k["v2_s"]["hex"]
    assert sig.v1specific is None and sig.verify()
    sig.s = (sig.s + 1) % N or 1
    assert not sig.verify()
    h = plume_b200.hash_to_curve(msg, sig.pk, ctx=gpu_ctx)
    inter = golden["intermediates"]["h"]
    assert "%064x" % h[0] == inter["x"] and "%064x" % h[1] == inter["y"]


def test_cpp_host_mirror(golden, tmp_path):
    """zk-nullifier-sig_b200/host/plume.hpp compiled with g++ against the C ABI: the reference's own signing test
    (mock RNG -> sign_v1 / sign_v2 -> c, s; verify) and a batch."""
    import subprocess
    import plume_b200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(plume_b200.LIB_PATH)
    exe = str(tmp_path / "test_signing")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), "-I", os.path.join(libdir, "host"),
                    os.path.join(root, "tests", "cpp", "test_signing.cpp"), "-o", exe, "-L", libdir, "-lplume_b200",
                    "-Wl,-rpath," + libdir], check=True)
    k = golden["sign_kat"]
    out = subprocess.run([exe, k["message_ascii"], k["sk"]["hex"], k["r"]["hex"]], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "v1 %s %s 1 1" % (k["v1_c"]["hex"], k["v1_s"]["hex"])
    assert out[1] == "v2 %s %s 1 0" % (k["v2_c"]["hex"], k["v2_s"]["hex"])
    # the serde-JSON writer of the C++ mirror produces the same text as the Python mirror's (and reads back there)
    class Mock:
        def fill_bytes(self, buf):
            buf[:] = bytes.fromhex(k["r"]["hex"])
    sk = plume_b200.SecretKey.from_bytes(bytes.fromhex(k["sk"]["hex"]))
    msg = k["message_ascii"].encode()
    assert out[2] == "json1 " + plume_b200.PlumeSignature.sign_v1(sk, msg, Mock()).to_json()
    assert out[3] == "json2 " + plume_b200.PlumeSignature.sign_v2(sk, msg, Mock()).to_json()
    assert plume_b200.PlumeSignature.from_json(out[2][6:]).verify()
    assert out[4] == "tampered 0" and out[5] == "zero-sk rejected" and out[6] == "batch 1000"
    assert out[7] == "ark v1 %s %s 1 0" % (k["v1_c"]["hex"], k["v1_s"]["hex"])      # rust-arkworks/src/tests.rs:281-299
    assert out[8] == "ark v2 %s %s 1 0" % (k["v2_c"]["hex"], k["v2_s"]["hex"])
    assert out[9] == "identity-pk rejected"


def test_sec1_compressed_api(gpu_ctx):
    """33-byte SEC1 slots in and out (compress / decompress / sign_sec1 / verify_sec1) against the
    oracles: encodings of k*G from the reference's table, random and malformed slots, and a sign->verify round trip."""
    import c_oracle
    import plume_ref as R
    rnd = random.Random(31)
    rng = np.random.default_rng(31)
    # decompress: valid points, identity, random x (half off-curve), non-canonical x, bad prefixes
    slots = [R.compress33(R.pt_mul(R.G, k)) for k in range(0, 40)]
    slots += [bytes([rnd.choice([2, 3])]) + rnd.randrange(P).to_bytes(32, "big") for _ in range(400)]
    slots += [b"\x02" + P.to_bytes(32, "big"), b"\x03" + (2**256 - 1).to_bytes(32, "big"), b"\x04" + slots[1][1:],
              b"\x00" + b"\x01" + bytes(31), b"\x01" + bytes(32), b"\xff" * 33]
    blob = np.frombuffer(b"".join(slots), dtype=np.uint8)
    pts, ok = gpu_ctx.points_decompress(blob)
    for i, b in enumerate(slots):
        want, good = c_oracle.decompress33(b)
        assert int(ok[i]) == good and bytes(pts[i]) == want, i
    good = np.flatnonzero(ok)
    back = gpu_ctx.points_compress(np.ascontiguousarray(pts[good]))
    for j, i in enumerate(good):
        assert bytes(back[j]) == slots[i]
    # sign with compressed outputs == compress(sign outputs); verify on compressed inputs
    n = 3000
    msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
    for ver in (1, 2):
        o64 = gpu_ctx.sign_batch(ver, msgs, sk, r)
        o33 = gpu_ctx.sign_batch_sec1(ver, msgs, sk, r)
        assert np.array_equal(o33["status"], o64["status"]) and np.array_equal(o33["c"], o64["c"]) and np.array_equal(o33["s"], o64["s"])
        for k in ("pk", "nullifier", "r_point", "hashed_to_curve_r"):
            for i in range(0, n, 97):
                assert bytes(o33[k][i]) == c_oracle.compress33(bytes(o64[k][i])), (k, i)
        ok = gpu_ctx.verify_batch_sec1(ver, msgs, o33["pk"], o33["nullifier"], o33["c"], o33["s"], o33["r_point"], o33["hashed_to_curve_r"])
        assert ok.all()
        t = {k: o33[k].copy() for k in ("pk", "nullifier", "r_point", "hashed_to_curve_r")}
        t["pk"][0::4, 0] ^= 1                      # wrong parity: decodes to -pk
        t["nullifier"][1::4, 5] ^= 0x40            # different x (may or may not be on the curve)
        t["r_point"][2::4, 0] = 7                  # bad prefix
        ok = gpu_ctx.verify_batch_sec1(ver, msgs, t["pk"], t["nullifier"], o33["c"], o33["s"], t["r_point"], t["hashed_to_curve_r"])
        # oracle: decode each slot, then verify
        for i in list(range(0, 40)) + list(range(n - 40, n)):
            dec = [c_oracle.decompress33(bytes(t[k][i])) for k in ("pk", "nullifier", "r_point", "hashed_to_curve_r")]
            use = dec if ver == 1 else dec[:2]
            want = 0
            if all(g for _, g in use):
                want = int(c_oracle.verify_batch(ver, msgs[i:i + 1], dec[0][0], dec[1][0], o33["c"][i:i + 1], o33["s"][i:i + 1],
                                                 dec[2][0], dec[3][0])[0])
            assert int(ok[i]) == want, (ver, i)
        assert ok[3::4].all() and not ok[0::4].any()
        if ver == 1:
            assert not ok[2::4].any()


def test_small_batch_kernels_every_size(gpu_ctx):
    """Batches of at most PLUME_TEAM_MAX (4 096) items run the small-batch kernels (k_team.cu: 2 or 4 lanes per item, points
    added by shuffles).  Sizes that leave whole warps, part of a warp and part of a team's block unused, both versions, ragged
    messages; 4 097 items is the first size back on the throughput kernels."""
    rnd = random.Random(41)
    rng = np.random.default_rng(41)
    for n in (1, 2, 3, 5, 8, 31, 32, 33, 63, 64, 65, 127, 1000, 4096, 4097):
        msgs = [bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 32, 33, 65, 90]))) for _ in range(n)]
        sk, r = _rand_scalars(rng, n), _rand_scalars(rng, n)
        if n >= 3:
            sk[1] = 0                                        # a rejected item in the middle of a warp
            r[n - 1] = 0xFF
        for ver in (1, 2):
            o = _sign_both(gpu_ctx, ver, msgs, sk, r)
            bad = o["s"].copy()
            if n >= 2:
                bad[n // 2, 31] ^= 1
            t = dict(o); t["s"] = bad
            got = _verify_both(gpu_ctx, ver, msgs, t)
            want = (o["status"] == 0)
            if n >= 2:
                want[n // 2] = False
            assert np.array_equal(got.astype(bool), want)


@pytest.fixture(scope="module")
def throughput_ctx():
    """A context with the small-batch kernels switched off: every batch, however small, on the throughput kernels."""
    import plume_b200
    os.environ["PLUME_TEAM_MAX"] = "0"
    try:
        ctx = plume_b200.PlumeContext(0, fixed_window_bits=16)
    finally:
        del os.environ["PLUME_TEAM_MAX"]
    yield ctx
    ctx.close()


def test_small_batches_on_the_throughput_kernels(throughput_ctx):
    """The edge cases above run small batches, i.e. the small-batch kernels; the same cases with those switched off keep the
    throughput kernels (what a 2^20 batch runs) under the same edge-case coverage."""
    test_ragged_messages_and_bad_scalars(throughput_ctx)
    test_fixed_length_records(throughput_ctx)
    test_tampered_and_malformed_verify_inputs(throughput_ctx)
    test_identity_forgery_agrees_with_oracle(throughput_ctx)
    test_hash_to_curve_batch(throughput_ctx)
    import test_arkworks_flavour as A
    A.test_gpu_ark_sign_and_verify_match_oracle(throughput_ctx)
    throughput_ctx.self_test()

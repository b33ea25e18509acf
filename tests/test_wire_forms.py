"""Host-side wire helpers that need no GPU: the serde-JSON form of PlumeSignature (rust-k256/src/lib.rs:66,83; field
encodings per k256 0.13 -- unpinned by the reference, see plume_b200/api.py) and the SEC1-DER scalar framing of the JS
binding (javascript/src/lib.rs:97-117).  The GPU halves (public keys, decompression, verify) are in tests/test_gpu_context.py."""
import json

import pytest

import plume_ref as R


def _sig(golden, v1):
    import plume_b200
    k = golden["sign_kat"]
    sk, r = int(k["sk"]["hex"], 16), int(k["r"]["hex"], 16)
    msg = k["message_ascii"].encode()
    st, o = R.sign(1 if v1 else 2, msg, sk, r)
    assert st == 0
    v1f = plume_b200.PlumeSignatureV1Fields(o["r_point"], o["hashed_to_curve_r"]) if v1 else None
    return plume_b200.PlumeSignature(msg, o["pk"], o["nullifier"], o["c"], o["s"], v1f), k


def test_serde_json_writer_matches_reference_values(golden):
    for v1 in (True, False):
        sig, k = _sig(golden, v1)
        d = json.loads(sig.to_json())
        assert list(d) == ["message", "pk", "nullifier", "c", "s", "v1specific"]          # declaration order, lib.rs:67-80
        assert bytes(d["message"]) == k["message_ascii"].encode()
        assert d["c"] == k["v1_c" if v1 else "v2_c"]["hex"].upper() and len(d["c"]) == 64
        assert d["s"] == k["v1_s" if v1 else "v2_s"]["hex"].upper()
        inter = golden["intermediates"]
        for name, key in (("pk", "pk"), ("nullifier", "h_sk")):
            assert d[name] == ("%02X" % (2 + int(inter[key]["y"], 16) % 2)) + inter[key]["x"].upper()
        if v1:
            assert list(d["v1specific"]) == ["r_point", "hashed_to_curve_r"]               # lib.rs:84-89
            assert d["v1specific"]["r_point"][2:] == inter["g_r"]["x"].upper()
            assert d["v1specific"]["hashed_to_curve_r"][2:] == inter["h_r"]["x"].upper()
        else:
            assert d["v1specific"] is None
        assert sig.to_json() == json.dumps(d, separators=(",", ":"))


def test_sec1_der_scalar_reader():
    import plume_b200
    k = 0x519B423D715F8B581F4FA8EE59F4771A5B44C8130B4E3EACCA54A56DDA72B464
    body = bytes.fromhex("0201010420") + k.to_bytes(32, "big")
    pub = R.pt_mul(R.G, k)
    bits = b"\x00\x04" + pub[0].to_bytes(32, "big") + pub[1].to_bytes(32, "big")
    with_pub = body + b"\xa1\x44\x03\x42" + bits
    params = b"\xa0\x07" + bytes.fromhex("06052b8104000a")
    for der in (b"\x30" + bytes([len(body)]) + body,                                        # private key only
                b"\x30" + bytes([len(with_pub)]) + with_pub,                                # what to_sec1_der writes (109 bytes)
                b"\x30" + bytes([len(body + params)]) + body + params):                     # with the optional curve OID
        assert plume_b200.scalar_from_sec1_der(der, check_public_key=False) == k
    assert len(b"\x30" + bytes([len(with_pub)]) + with_pub) == 109
    bad = [b"\x31" + bytes([len(body)]) + body,                                             # not a SEQUENCE
           b"\x30" + bytes([len(body)]) + body[:-1],                                        # truncated
           b"\x30" + bytes([len(body)]) + bytes.fromhex("0201020420") + k.to_bytes(32, "big"),   # version 2
           b"\x30\x25" + bytes.fromhex("0201010420") + bytes(32),                           # zero scalar
           b"\x30\x25" + bytes.fromhex("0201010420") + R.N.to_bytes(32, "big"),             # scalar = n
           b"\x30" + bytes([len(body) + 9]) + body + b"\xa0\x07" + bytes.fromhex("06052b81040022")]   # another curve's OID
    for der in bad:
        with pytest.raises(ValueError):
            plume_b200.scalar_from_sec1_der(der, check_public_key=False)

"""ctypes wrapper of oracle/libplume_oracle.so (TEST INFRASTRUCTURE ONLY -- see plume_oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Builds the library with `make -C oracle` on first use if it is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libplume_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "plume_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.plume_oracle_encode_pt.restype = ctypes.c_size_t
        _lib = L
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _u8(a, shape):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8).reshape(shape)


def _msgs(msgs):
    if isinstance(msgs, (list, tuple)):
        offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
        if msgs:
            offs[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8) if msgs else np.zeros(0, dtype=np.uint8)
        return np.ascontiguousarray(blob), offs, 0, len(msgs)
    a = np.ascontiguousarray(msgs, dtype=np.uint8)
    return a, None, a.shape[1], a.shape[0]


def sign_batch(version, msgs, sk, r, threads=1):
    blob, offs, mlen, n = _msgs(msgs)
    sk = _u8(sk, (n, 32)); r = _u8(r, (n, 32))
    o = {k: np.zeros((n, w), dtype=np.uint8) for k, w in
         (("pk", 64), ("nullifier", 64), ("c", 32), ("s", 32), ("r_point", 64), ("hashed_to_curve_r", 64))}
    o["status"] = np.zeros(n, dtype=np.uint8)
    lib().plume_oracle_sign_batch(version, ctypes.c_size_t(n), _p(blob), _p(offs), ctypes.c_size_t(mlen), _p(sk), _p(r),
                                  _p(o["pk"]), _p(o["nullifier"]), _p(o["c"]), _p(o["s"]), _p(o["r_point"]),
                                  _p(o["hashed_to_curve_r"]), _p(o["status"]), int(threads))
    return o


def verify_batch(version, msgs, pk, nullifier, c, s, r_point=None, hashed_to_curve_r=None, threads=1):
    blob, offs, mlen, n = _msgs(msgs)
    pk = _u8(pk, (n, 64)); nullifier = _u8(nullifier, (n, 64)); c = _u8(c, (n, 32)); s = _u8(s, (n, 32))
    rp = None if r_point is None else _u8(r_point, (n, 64))
    hr = None if hashed_to_curve_r is None else _u8(hashed_to_curve_r, (n, 64))
    ok = np.zeros(n, dtype=np.uint8)
    lib().plume_oracle_verify_batch(version, ctypes.c_size_t(n), _p(blob), _p(offs), ctypes.c_size_t(mlen), _p(pk), _p(nullifier),
                                    _p(c), _p(s), _p(rp), _p(hr), _p(ok), int(threads))
    return ok


def h2c_batch(msgs, threads=1):
    blob, offs, mlen, n = _msgs(msgs)
    out = np.zeros((n, 64), dtype=np.uint8)
    lib().plume_oracle_h2c_batch(ctypes.c_size_t(n), _p(blob), _p(offs), ctypes.c_size_t(mlen), _p(out), int(threads))
    return out


def mul_g(k):
    out = (ctypes.c_uint8 * 64)()
    lib().plume_oracle_mul_g(k.to_bytes(32, "big"), out)
    return bytes(out)


def mul(p64, k):
    out = (ctypes.c_uint8 * 64)()
    lib().plume_oracle_mul(bytes(p64), k.to_bytes(32, "big"), out)
    return bytes(out)


def encode_pt(p64):
    out = (ctypes.c_uint8 * 33)()
    n = lib().plume_oracle_encode_pt(bytes(p64), out)
    return bytes(out[:n])


def expand_message_xmd(msg, n):
    out = (ctypes.c_uint8 * n)()
    lib().plume_oracle_expand_message_xmd(bytes(msg), ctypes.c_size_t(len(msg)), ctypes.c_size_t(n), out)
    return bytes(out)


def sha256(msg):
    out = (ctypes.c_uint8 * 32)()
    lib().plume_oracle_sha256(bytes(msg), ctypes.c_size_t(len(msg)), out)
    return bytes(out)


def compress33(p64):
    out = (ctypes.c_uint8 * 33)()
    lib().plume_oracle_compress33(bytes(p64), out)
    return bytes(out)


def decompress33(b33):
    out = (ctypes.c_uint8 * 64)()
    ok = lib().plume_oracle_decompress33(bytes(b33), out)
    return bytes(out), int(ok)

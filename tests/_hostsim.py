"""Builds and wraps tests/hostsim/libplume_hostsim.so: the kernel sources (zk-nullifier-sig_b200/csrc/*.cuh)
compiled for the host with the PTX wrappers emulated (see tests/hostsim/hostsim.cpp).  Test-only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
EXTRA = [d for d in os.environ.get("PLUME_HOSTSIM_DEFINES", "").split() if d]   # e.g. "PLUME_SQR2 PLUME_MUL2": experiment builds
LIB = os.path.join(ROOT, "tests", "hostsim", "libplume_hostsim%s.so" % ("_" + "_".join(EXTRA).replace("=", "") if EXTRA else ""))
CSRC = os.path.join(ROOT, "zk-nullifier-sig_b200", "csrc")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            subprocess.run(["g++", "-O2", "-std=c++17", "-DPLUME_HOSTSIM"] + ["-D" + d for d in EXTRA] + ["-fPIC", "-shared", "-I", CSRC, "-o", LIB, SRC], check=True)
        _lib = ctypes.CDLL(LIB)
    return _lib


def limbs(x, n=8):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(a):
    return sum(int(a[i]) << (32 * i) for i in range(len(a)))


def fe_op(op, a, b=0):
    out = (ctypes.c_uint32 * 8)()
    lib().hs_fe_op(op, limbs(a), limbs(b), out)
    return val(out)


def sc_op(op, a, b=0):
    out = (ctypes.c_uint32 * 8)()
    lib().hs_sc_op(op, limbs(a), limbs(b), out)
    return val(out)


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _msgs(msgs):
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    if msgs:
        offs[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8)
    return np.ascontiguousarray(blob), offs


def ark_sign_batch(version, msgs, pk, sk, r, gw=8, binv_threads=3):
    """arkworks flavour: pk is an input; returns nullifier, digest_private, s, r_point, hashed_to_curve_r, status."""
    n = len(msgs)
    blob, offs = _msgs(msgs)
    pk = np.frombuffer(pk, dtype=np.uint8).copy(); sk = np.frombuffer(sk, dtype=np.uint8).copy(); r = np.frombuffer(r, dtype=np.uint8).copy()
    o = {k: np.zeros((n, w), dtype=np.uint8) for k, w in
         (("nullifier", 64), ("digest_private", 32), ("s", 32), ("r_point", 64), ("hashed_to_curve_r", 64))}
    o["status"] = np.zeros(n, dtype=np.uint8)
    lib().hs_sign_batch(1, 2, version, n, _p(blob), _p(offs), 0, _p(sk), _p(r), _p(pk), _p(o["nullifier"]), _p(o["digest_private"]),
                        _p(o["s"]), _p(o["r_point"]), _p(o["hashed_to_curve_r"]), _p(o["status"]), gw, binv_threads)
    return o


def ark_verify_batch(version, msgs, pk, nullifier, digest_private, s, r_point, hashed_to_curve_r, gw=8, binv_threads=3):
    n = len(msgs)
    blob, offs = _msgs(msgs)
    a = [np.ascontiguousarray(x, dtype=np.uint8) for x in (pk, nullifier, digest_private, s, r_point, hashed_to_curve_r)]
    ok = np.zeros(n, dtype=np.uint8)
    lib().hs_verify_batch(1, version, n, _p(blob), _p(offs), 0, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), _p(ok),
                          gw, binv_threads, 0)
    return ok


def sign_batch(version, msgs, sk, r, gw=8, binv_threads=3, comb=True):
    n = len(msgs)
    blob, offs = _msgs(msgs)
    sk = np.frombuffer(sk, dtype=np.uint8).copy(); r = np.frombuffer(r, dtype=np.uint8).copy()
    o = {k: np.zeros((n, w), dtype=np.uint8) for k, w in
         (("pk", 64), ("nullifier", 64), ("c", 32), ("s", 32), ("r_point", 64), ("hashed_to_curve_r", 64))}
    o["status"] = np.zeros(n, dtype=np.uint8)
    lib().hs_sign_batch(0, int(comb), version, n, _p(blob), _p(offs), 0, _p(sk), _p(r), _p(o["pk"]), _p(o["nullifier"]), _p(o["c"]), _p(o["s"]),
                        _p(o["r_point"]), _p(o["hashed_to_curve_r"]), _p(o["status"]), gw, binv_threads)
    return o


def verify_batch(version, msgs, pk, nullifier, c, s, r_point, hashed_to_curve_r, gw=8, binv_threads=3, fused=False):
    n = len(msgs)
    blob, offs = _msgs(msgs)
    a = [np.ascontiguousarray(x, dtype=np.uint8) for x in (pk, nullifier, c, s, r_point, hashed_to_curve_r)]
    ok = np.zeros(n, dtype=np.uint8)
    lib().hs_verify_batch(0, version, n, _p(blob), _p(offs), 0, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), _p(ok),
                          gw, binv_threads, int(fused))
    return ok


def h2c_batch(msgs, binv_threads=3):
    n = len(msgs)
    blob, offs = _msgs(msgs)
    out = np.zeros((n, 64), dtype=np.uint8)
    lib().hs_h2c_batch(n, _p(blob), _p(offs), 0, _p(out), binv_threads)
    return out


def h2c_witness_batch(msgs, binv_threads=3):
    n = len(msgs)
    blob, offs = _msgs(msgs)
    o = {"u": np.zeros((n, 2, 32), dtype=np.uint8), "q": np.zeros((n, 2, 64), dtype=np.uint8),
         "gx1_square": np.zeros((n, 2), dtype=np.uint8), "h": np.zeros((n, 64), dtype=np.uint8),
         "hints": np.zeros((n, 2, 3, 32), dtype=np.uint8)}
    lib().hs_h2c_witness_batch(n, _p(blob), _p(offs), 0, _p(o["u"]), _p(o["q"]), _p(o["gx1_square"]), _p(o["h"]), binv_threads,
                               _p(o["hints"]))
    return o


def registers(values32):
    a = np.ascontiguousarray(values32, dtype=np.uint8)
    n = a.size // 32
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().hs_registers(n, _p(a), _p(out))
    return out

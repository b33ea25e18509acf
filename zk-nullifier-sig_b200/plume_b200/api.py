"""Host-side mirror of the reference's operator interface for the PLUME hot path, over the C ABI.

Reference (Rust, crate `plume_rustcrypto`, /root/reference/rust-k256):
  PlumeSignature{message, pk, nullifier, c, s, v1specific}        src/lib.rs:67-80
  PlumeSignatureV1Fields{r_point, hashed_to_curve_r}              src/lib.rs:84-89
  PlumeSignature::verify / sign_v1 / sign_v2                      src/lib.rs:93,149,154
  PlumeSigner::new + try_sign_with_rng                            src/randomizedsigner.rs:38,43
  DST                                                             src/lib.rs:61

Same names, argument meaning and error behaviour; the batch entry points are the additions.
Points are (x, y) tuples of ints or None for the identity; scalars are ints.  All compute goes
through libplume_b200.so on the GPU -- there is no Python/CPU implementation in this package.
"""
import ctypes

import numpy as np

from . import _lib

DST = b"QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_"  # rust-k256/src/lib.rs:61
ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141

STATUS_TEXT = {  # the `expect` messages of the reference at the places it panics
    1: "nonce r outside [1, n-1]",
    2: "secret key outside [1, n-1]",
    3: "it should be impossible to get the hash equal to zero",                      # randomizedsigner.rs:91
    4: "something is terribly wrong if the nonce is equal to negated product of the secret and the hash",  # :95
    5: "something is drammatically wrong if the input hashed to the identity",       # :61
}


class PlumeError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _as_u8(a, shape=None):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8)
    if shape is not None:
        a = a.reshape(shape)
    return a


def pack_messages(msgs):
    """list of bytes -> (blob u8[total], offsets u64[n+1])"""
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    if msgs:
        offs[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    blob = np.frombuffer(b"".join(msgs), dtype=np.uint8) if msgs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(blob), offs


class PlumeContext:
    """A signer/verifier context: one GPU (plume_ctx_create) or, when `device` is a list of ordinals, several GPUs of
    this process behind one context (plume_ctx_create_multi: the host-pointer batch calls range-split over them)."""

    def __init__(self, device=0, fixed_window_bits=0, _handle=None):
        self._lib = _lib.load()
        self._owned = _handle is None
        if _handle is not None:
            self._h = ctypes.c_void_p(_handle)
            self.device = device
            return
        h = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (ctypes.c_int * len(device))(*[int(d) for d in device])
            rc = self._lib.plume_ctx_create_multi(ctypes.byref(h), devs, len(device), int(fixed_window_bits))
            what = "plume_ctx_create_multi"
        else:
            rc = self._lib.plume_ctx_create(ctypes.byref(h), int(device), int(fixed_window_bits))
            what = "plume_ctx_create"
        if rc != 0:
            raise PlumeError("%s failed (%d): %s" % (what, rc, self._lib.plume_last_error(None).decode()))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._lib.plume_ctx_destroy(self._h)
            self._h = None

    @property
    def device_count(self):
        return self._lib.plume_ctx_device_count(self._h)

    def sub(self, i):
        """The i-th per-device context of a multi-device context (borrowed: closing it is a no-op)."""
        h = self._lib.plume_ctx_sub(self._h, int(i))
        if not h:
            raise PlumeError("no sub-context %d" % i)
        dev = self.device[i] if isinstance(self.device, (list, tuple)) else self.device
        return PlumeContext(dev, _handle=h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise PlumeError("%s failed (%d): %s" % (what, rc, self._lib.plume_last_error(self._h).decode()))

    def self_test(self):
        """plume_self_test: the reference's known-answer vectors through this context's own batch calls (every device of a
        multi-device context); raises PlumeError naming the field that differs."""
        self._check(self._lib.plume_self_test(self._h), "plume_self_test")

    @property
    def chunk_items(self):
        return self._lib.plume_ctx_chunk_items(self._h)

    @property
    def launch_count(self):
        return self._lib.plume_ctx_launch_count(self._h)

    def set_profiling(self, on):
        self._check(self._lib.plume_ctx_set_profiling(self._h, 1 if on else 0), "plume_ctx_set_profiling")

    def stage_ms(self, stage):
        n = ctypes.c_uint64(0)
        ms = self._lib.plume_ctx_stage_ms(self._h, stage.encode(), ctypes.byref(n))
        return ms, n.value

    def measure_imad_peak(self, iters=4096):
        v = ctypes.c_double(0)
        self._check(self._lib.plume_measure_imad_peak(self._h, iters, ctypes.byref(v)), "plume_measure_imad_peak")
        return v.value

    def measure_imad_rates(self, iters=4096):
        """(plain IMAD.WIDE.U32, carry-chain IMAD.WIDE.U32.X) limb products per second."""
        p, c = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._lib.plume_measure_imad_rates(self._h, iters, ctypes.byref(p), ctypes.byref(c)), "plume_measure_imad_rates")
        return p.value, c.value

    def debug_fe_op(self, op, a, b):
        """plume_debug_fe_op on u32[n,8] little-endian limb arrays."""
        a = np.ascontiguousarray(a, dtype=np.uint32); b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.empty_like(a)
        self._check(self._lib.plume_debug_fe_op(self._h, op, a.shape[0], _ptr(a), _ptr(b), _ptr(out)), "plume_debug_fe_op")
        return out

    # ---- batch entry points on host (numpy) buffers ---------------------------------------------------
    @staticmethod
    def _msgs(msgs, msg_len):
        """msgs: list of bytes | (blob, offsets) | u8 array [n, msg_len]"""
        if isinstance(msgs, (list, tuple)) and (len(msgs) == 0 or isinstance(msgs[0], (bytes, bytearray))):
            blob, offs = pack_messages(list(msgs))
            return blob, offs, 0, len(msgs)
        if isinstance(msgs, tuple):
            blob, offs = msgs
            return _as_u8(blob), np.ascontiguousarray(offs, dtype=np.uint64), 0, len(offs) - 1
        a = _as_u8(msgs)
        if a.ndim != 2:
            raise ValueError("fixed-length messages must be a [n, msg_len] u8 array")
        return a, None, a.shape[1], a.shape[0]

    def sign_batch(self, version, msgs, sk, r, out=None):
        """plume_sign_batch.  sk, r: u8[n,32] big-endian.  Returns dict of u8 arrays + status."""
        blob, offs, mlen, n = self._msgs(msgs, None)
        sk = _as_u8(sk, (n, 32))
        r = _as_u8(r, (n, 32))
        o = out or {}
        for k, w in (("pk", 64), ("nullifier", 64), ("c", 32), ("s", 32), ("r_point", 64), ("hashed_to_curve_r", 64)):
            if k not in o:
                o[k] = np.empty((n, w), dtype=np.uint8)
        if "status" not in o:
            o["status"] = np.empty(n, dtype=np.uint8)
        rc = self._lib.plume_sign_batch(self._h, version, n, _ptr(blob), _ptr(offs), mlen, _ptr(sk), _ptr(r),
                                        _ptr(o["pk"]), _ptr(o["nullifier"]), _ptr(o["c"]), _ptr(o["s"]),
                                        _ptr(o["r_point"]), _ptr(o["hashed_to_curve_r"]), _ptr(o["status"]))
        self._check(rc, "plume_sign_batch")
        return o

    def verify_batch(self, version, msgs, pk, nullifier, c, s, r_point=None, hashed_to_curve_r=None, out=None):
        blob, offs, mlen, n = self._msgs(msgs, None)
        pk = _as_u8(pk, (n, 64)); nullifier = _as_u8(nullifier, (n, 64))
        c = _as_u8(c, (n, 32)); s = _as_u8(s, (n, 32))
        rp = None if r_point is None else _as_u8(r_point, (n, 64))
        hr = None if hashed_to_curve_r is None else _as_u8(hashed_to_curve_r, (n, 64))
        ok = out if out is not None else np.empty(n, dtype=np.uint8)
        rc = self._lib.plume_verify_batch(self._h, version, n, _ptr(blob), _ptr(offs), mlen, _ptr(pk), _ptr(nullifier),
                                          _ptr(c), _ptr(s), _ptr(rp), _ptr(hr), _ptr(ok))
        self._check(rc, "plume_verify_batch")
        return ok

    # ---- circuit-input side ----------------------------------------------------------------
    def hash_to_curve_witness_batch(self, msgs):
        """plume_hash_to_curve_witness_batch: dict u [n,2,32], q [n,2,64], gx1_square [n,2], h [n,64]."""
        blob, offs, mlen, n = self._msgs(msgs, None)
        o = {"u": np.empty((n, 2, 32), dtype=np.uint8), "q": np.empty((n, 2, 64), dtype=np.uint8),
             "gx1_square": np.empty((n, 2), dtype=np.uint8), "h": np.empty((n, 64), dtype=np.uint8),
             "hints": np.empty((n, 2, 3, 32), dtype=np.uint8)}   # per u: gx1_sqrt, gx2_sqrt, y_pos (declared convention)
        rc = self._lib.plume_hash_to_curve_witness_batch(self._h, n, _ptr(blob), _ptr(offs), mlen, _ptr(o["u"]), _ptr(o["q"]),
                                                         _ptr(o["gx1_square"]), _ptr(o["h"]), _ptr(o["hints"]))
        self._check(rc, "plume_hash_to_curve_witness_batch")
        return o

    def registers_batch(self, values32):
        """plume_registers_batch: u8[..., 32] big-endian -> u64[..., 4] (circuits/circom/utils.ts scalarToCircuitValue)."""
        a = _as_u8(values32)
        n = a.size // 32
        out = np.empty((n, 4), dtype=np.uint64)
        self._check(self._lib.plume_registers_batch(self._h, n, _ptr(a), _ptr(out)), "plume_registers_batch")
        return out.reshape(a.shape[:-1] + (4,)) if a.ndim > 1 else out

    def fixed_base_mul_batch(self, scalars32):
        """plume_fixed_base_mul_batch: u8[n,32] big-endian scalars -> u8[n,64] points k*G."""
        a = _as_u8(scalars32)
        n = a.size // 32
        a = a.reshape(n, 32)
        out = np.empty((n, 64), dtype=np.uint8)
        self._check(self._lib.plume_fixed_base_mul_batch(self._h, n, _ptr(a), _ptr(out)), "plume_fixed_base_mul_batch")
        return out

    def debug_read_arena(self, lane, cap=1 << 26):
        """(device arena bytes, pinned staging bytes) of lane 0/1 after the streams drained -- test hook."""
        d = np.zeros(cap, dtype=np.uint8); h = np.zeros(cap, dtype=np.uint8)
        nd, nh = ctypes.c_size_t(0), ctypes.c_size_t(0)
        self._check(self._lib.plume_debug_read_arena(self._h, lane, _ptr(d), _ptr(h), cap, ctypes.byref(nd), ctypes.byref(nh)),
                    "plume_debug_read_arena")
        return d[:nd.value], h[:nh.value]

    # ---- arkworks flavour (rust-arkworks/src/lib.rs:229-278, tests.rs:28-78) -----------------------------
    def ark_sign_batch(self, version, msgs, pk, sk, r):
        """plume_ark_sign_batch.  pk: u8[n,64] (input); sk, r: u8[n,32] big-endian Fr (zero allowed)."""
        blob, offs, mlen, n = self._msgs(msgs, None)
        pk = _as_u8(pk, (n, 64)); sk = _as_u8(sk, (n, 32)); r = _as_u8(r, (n, 32))
        o = {k: np.empty((n, w), dtype=np.uint8) for k, w in
             (("nullifier", 64), ("digest_private", 32), ("s", 32), ("r_point", 64), ("hashed_to_curve_r", 64))}
        o["status"] = np.empty(n, dtype=np.uint8)
        rc = self._lib.plume_ark_sign_batch(self._h, version, n, _ptr(blob), _ptr(offs), mlen, _ptr(pk), _ptr(sk), _ptr(r),
                                            _ptr(o["nullifier"]), _ptr(o["digest_private"]), _ptr(o["s"]), _ptr(o["r_point"]),
                                            _ptr(o["hashed_to_curve_r"]), _ptr(o["status"]))
        self._check(rc, "plume_ark_sign_batch")
        return o

    def ark_verify_batch(self, version, msgs, pk, nullifier, digest_private, s, r_point, hashed_to_curve_r):
        blob, offs, mlen, n = self._msgs(msgs, None)
        a = [_as_u8(x, (n, w)) for x, w in ((pk, 64), (nullifier, 64), (digest_private, 32), (s, 32), (r_point, 64),
                                             (hashed_to_curve_r, 64))]
        ok = np.empty(n, dtype=np.uint8)
        rc = self._lib.plume_ark_verify_batch(self._h, version, n, _ptr(blob), _ptr(offs), mlen, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]),
                                              _ptr(a[3]), _ptr(a[4]), _ptr(a[5]), _ptr(ok))
        self._check(rc, "plume_ark_verify_batch")
        return ok

    # ---- SEC1-compressed wire form (33-byte slots) ----------------------------------------------------
    def points_compress(self, pts64):
        a = _as_u8(pts64); n = a.size // 64; a = a.reshape(n, 64)
        out = np.empty((n, 33), dtype=np.uint8)
        self._check(self._lib.plume_points_compress_batch(self._h, n, _ptr(a), _ptr(out)), "plume_points_compress_batch")
        return out

    def points_decompress(self, pts33):
        a = _as_u8(pts33); n = a.size // 33; a = a.reshape(n, 33)
        out = np.empty((n, 64), dtype=np.uint8); ok = np.empty(n, dtype=np.uint8)
        self._check(self._lib.plume_points_decompress_batch(self._h, n, _ptr(a), _ptr(out), _ptr(ok)), "plume_points_decompress_batch")
        return out, ok

    def points_compress_device(self, n, in64, out33, stream=0):
        vp = ctypes.c_void_p
        self._check(self._lib.plume_points_compress_batch_device(self._h, n, vp(in64), vp(out33), vp(stream or None)), "plume_points_compress_batch_device")

    def points_decompress_device(self, n, in33, out64, ok, stream=0):
        vp = ctypes.c_void_p
        self._check(self._lib.plume_points_decompress_batch_device(self._h, n, vp(in33), vp(out64), vp(ok), vp(stream or None)), "plume_points_decompress_batch_device")

    def sign_batch_sec1(self, version, msgs, sk, r):
        blob, offs, mlen, n = self._msgs(msgs, None)
        sk = _as_u8(sk, (n, 32)); r = _as_u8(r, (n, 32))
        o = {k: np.empty((n, w), dtype=np.uint8) for k, w in
             (("pk", 33), ("nullifier", 33), ("c", 32), ("s", 32), ("r_point", 33), ("hashed_to_curve_r", 33))}
        o["status"] = np.empty(n, dtype=np.uint8)
        rc = self._lib.plume_sign_batch_sec1(self._h, version, n, _ptr(blob), _ptr(offs), mlen, _ptr(sk), _ptr(r), _ptr(o["pk"]),
                                             _ptr(o["nullifier"]), _ptr(o["c"]), _ptr(o["s"]), _ptr(o["r_point"]),
                                             _ptr(o["hashed_to_curve_r"]), _ptr(o["status"]))
        self._check(rc, "plume_sign_batch_sec1")
        return o

    def verify_batch_sec1(self, version, msgs, pk33, nullifier33, c, s, r_point33=None, hashed_to_curve_r33=None):
        blob, offs, mlen, n = self._msgs(msgs, None)
        pk = _as_u8(pk33, (n, 33)); nul = _as_u8(nullifier33, (n, 33)); c = _as_u8(c, (n, 32)); s = _as_u8(s, (n, 32))
        rp = None if r_point33 is None else _as_u8(r_point33, (n, 33))
        hr = None if hashed_to_curve_r33 is None else _as_u8(hashed_to_curve_r33, (n, 33))
        ok = np.empty(n, dtype=np.uint8)
        rc = self._lib.plume_verify_batch_sec1(self._h, version, n, _ptr(blob), _ptr(offs), mlen, _ptr(pk), _ptr(nul), _ptr(c),
                                               _ptr(s), _ptr(rp), _ptr(hr), _ptr(ok))
        self._check(rc, "plume_verify_batch_sec1")
        return ok

    # ---- device-pointer entry points (raw addresses, e.g. torch tensors' data_ptr()) -----------------
    def sign_batch_device(self, version, n, msgs, msg_offsets, msg_len, sk, r, pk, nullifier, c, s,
                          r_point, hashed_to_curve_r, status, stream=0):
        vp = ctypes.c_void_p
        rc = self._lib.plume_sign_batch_device(self._h, version, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(sk), vp(r),
                                               vp(pk), vp(nullifier), vp(c), vp(s), vp(r_point or None),
                                               vp(hashed_to_curve_r or None), vp(status), vp(stream or None))
        self._check(rc, "plume_sign_batch_device")

    def verify_batch_device(self, version, n, msgs, msg_offsets, msg_len, pk, nullifier, c, s, r_point,
                            hashed_to_curve_r, ok, stream=0):
        vp = ctypes.c_void_p
        rc = self._lib.plume_verify_batch_device(self._h, version, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(pk),
                                                 vp(nullifier), vp(c), vp(s), vp(r_point or None),
                                                 vp(hashed_to_curve_r or None), vp(ok), vp(stream or None))
        self._check(rc, "plume_verify_batch_device")

    def hash_to_curve_batch_device(self, n, msgs, msg_offsets, msg_len, out, stream=0):
        vp = ctypes.c_void_p
        rc = self._lib.plume_hash_to_curve_batch_device(self._h, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(out),
                                                        vp(stream or None))
        self._check(rc, "plume_hash_to_curve_batch_device")

    # raw host addresses (e.g. pinned torch tensors): same C entry points as sign_batch / verify_batch
    def sign_batch_ptr(self, version, n, msgs, msg_offsets, msg_len, sk, r, pk, nullifier, c, s, r_point,
                       hashed_to_curve_r, status):
        vp = ctypes.c_void_p
        rc = self._lib.plume_sign_batch(self._h, version, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(sk), vp(r), vp(pk),
                                        vp(nullifier), vp(c), vp(s), vp(r_point or None), vp(hashed_to_curve_r or None),
                                        vp(status))
        self._check(rc, "plume_sign_batch")

    def sign_batch_sec1_ptr(self, version, n, msgs, msg_offsets, msg_len, sk, r, pk, nullifier, c, s, r_point, hashed_to_curve_r, status):
        vp = ctypes.c_void_p
        rc = self._lib.plume_sign_batch_sec1(self._h, version, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(sk), vp(r), vp(pk),
                                             vp(nullifier), vp(c), vp(s), vp(r_point or None), vp(hashed_to_curve_r or None), vp(status))
        self._check(rc, "plume_sign_batch_sec1")

    def verify_batch_sec1_ptr(self, version, n, msgs, msg_offsets, msg_len, pk, nullifier, c, s, r_point, hashed_to_curve_r, ok):
        vp = ctypes.c_void_p
        rc = self._lib.plume_verify_batch_sec1(self._h, version, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(pk), vp(nullifier),
                                               vp(c), vp(s), vp(r_point or None), vp(hashed_to_curve_r or None), vp(ok))
        self._check(rc, "plume_verify_batch_sec1")

    def verify_batch_ptr(self, version, n, msgs, msg_offsets, msg_len, pk, nullifier, c, s, r_point, hashed_to_curve_r, ok):
        vp = ctypes.c_void_p
        rc = self._lib.plume_verify_batch(self._h, version, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(pk),
                                          vp(nullifier), vp(c), vp(s), vp(r_point or None), vp(hashed_to_curve_r or None),
                                          vp(ok))
        self._check(rc, "plume_verify_batch")

    def hash_to_curve_batch_ptr(self, n, msgs, msg_offsets, msg_len, out):
        vp = ctypes.c_void_p
        rc = self._lib.plume_hash_to_curve_batch(self._h, n, vp(msgs), vp(msg_offsets or None), msg_len, vp(out))
        self._check(rc, "plume_hash_to_curve_batch")

    def hash_to_curve_pk_batch(self, msgs, pk33):
        """plume_hash_to_curve_pk_batch: h_i = hash_to_curve(m_i, pk_i) with pk as 33-byte SEC1 slots (utils.rs:11-20)."""
        blob, offs, mlen, n = self._msgs(msgs, None)
        pk = _as_u8(pk33, (n, 33))
        o = np.empty((n, 64), dtype=np.uint8)
        rc = self._lib.plume_hash_to_curve_pk_batch(self._h, n, _ptr(blob), _ptr(offs), mlen, _ptr(pk), _ptr(o))
        self._check(rc, "plume_hash_to_curve_pk_batch")
        return o

    def hash_to_curve_batch(self, msgs, out=None):
        blob, offs, mlen, n = self._msgs(msgs, None)
        o = out if out is not None else np.empty((n, 64), dtype=np.uint8)
        rc = self._lib.plume_hash_to_curve_batch(self._h, n, _ptr(blob), _ptr(offs), mlen, _ptr(o))
        self._check(rc, "plume_hash_to_curve_batch")
        return o


_default_ctx = None


def default_context():
    """The context behind the single-signature calls (sign_v1 / sign_v2 / verify / hash_to_curve without an explicit ctx).
    Latency, not throughput, matters there, so it takes the small generator table (16-bit windows: 64 MiB, built in
    milliseconds) instead of the 3.2 GB one of a batch context; its workspaces grow with the batches it actually sees."""
    global _default_ctx
    if _default_ctx is None:
        ctx = PlumeContext(0, fixed_window_bits=16)
        ctx.self_test()          # once per process: a GPU / driver / build that disagrees with the reference's vectors stops here
        _default_ctx = ctx
    return _default_ctx


# ---- conversions -----------------------------------------------------------------------------------------
def point_to_bytes(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def point_from_bytes(b):
    b = bytes(b)
    if b == bytes(64):
        return None
    return (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))


def encode_pt(p):
    """SEC1 compressed (rust-k256/src/utils.rs:23-25)."""
    return b"\x00" if p is None else bytes([2 + (p[1] & 1)]) + p[0].to_bytes(32, "big")


# ---- SEC1-DER scalars: the JS wire form of c and s (javascript/src/lib.rs:97-117) -------------------------------------
# `SecretKey::from(scalar).to_sec1_der()` of elliptic-curve 0.13 / sec1 0.7 (external crates, not in the reference tree;
# no reference test pins the bytes): ECPrivateKey ::= SEQUENCE { version INTEGER 1, privateKey OCTET STRING (32),
# publicKey [1] BIT STRING (uncompressed k*G) } with the curve parameters omitted:
#     30 6b  02 01 01  04 20 <k: 32 bytes>  a1 44 03 42 00  04 <x: 32> <y: 32>            (109 bytes)
_DER_HEAD = bytes.fromhex("306b0201010420")
_DER_MID = bytes.fromhex("a14403420004")


def scalars_to_sec1_der(scalars32, ctx=None):
    """u8[n,32] big-endian scalars in [1, n-1] -> list of 109-byte SEC1-DER EC private keys (public keys k*G from the GPU)."""
    a = _as_u8(scalars32)
    n = a.size // 32
    a = a.reshape(n, 32)
    for row in a:
        if not (1 <= int.from_bytes(bytes(row), "big") < ORDER):
            raise ValueError("scalar outside [1, n-1] has no SecretKey form")
    pub = (ctx or default_context()).fixed_base_mul_batch(a)
    return [_DER_HEAD + bytes(a[i]) + _DER_MID + bytes(pub[i]) for i in range(n)]


def scalar_from_sec1_der(der, ctx=None, check_public_key=True):
    """`SecretKey::from_sec1_der`: accepts the form above, and the same without the optional public key or with the
    optional secp256k1 parameters ([0] OID 1.3.132.0.10); when a public key is present it must equal k*G (k256 rejects a
    mismatch).  Returns the scalar as an int; raises ValueError on anything else."""
    der = bytes(der)

    def tlv(buf, pos):
        if pos + 2 > len(buf):
            raise ValueError("truncated DER")
        tag, ln = buf[pos], buf[pos + 1]
        pos += 2
        if ln & 0x80:
            k = ln & 0x7F
            if k == 0 or k > 2 or pos + k > len(buf):
                raise ValueError("bad DER length")
            ln = int.from_bytes(buf[pos:pos + k], "big")
            pos += k
        if pos + ln > len(buf):
            raise ValueError("truncated DER")
        return tag, buf[pos:pos + ln], pos + ln

    tag, body, end = tlv(der, 0)
    if tag != 0x30 or end != len(der):
        raise ValueError("not a DER SEQUENCE")
    tag, ver, pos = tlv(body, 0)
    if tag != 0x02 or ver != b"\x01":
        raise ValueError("ECPrivateKey version must be 1")
    tag, key, pos = tlv(body, pos)
    if tag != 0x04 or len(key) != 32:
        raise ValueError("privateKey must be a 32-byte OCTET STRING")
    k = int.from_bytes(key, "big")
    if not (1 <= k < ORDER):
        raise ValueError("scalar outside [1, n-1]")
    pub = None
    while pos < len(body):
        tag, val, pos = tlv(body, pos)
        if tag == 0xA0:
            if val != bytes.fromhex("06052b8104000a"):
                raise ValueError("parameters are not secp256k1")
        elif tag == 0xA1:
            t2, bits, e2 = tlv(val, 0)
            if t2 != 0x03 or e2 != len(val) or len(bits) != 66 or bits[0] != 0 or bits[1] != 4:
                raise ValueError("publicKey must be an uncompressed point")
            pub = bits[2:]
        else:
            raise ValueError("unexpected field in ECPrivateKey")
    if pub is not None and check_public_key:
        want = (ctx or default_context()).fixed_base_mul_batch(np.frombuffer(key, dtype=np.uint8))
        if bytes(want[0]) != pub:
            raise ValueError("public key does not match the private key")
    return k


# ---- the reference's types ----------------------------------------------------------------------------------
class PlumeSignatureV1Fields:
    """rust-k256/src/lib.rs:84-89"""

    def __init__(self, r_point, hashed_to_curve_r):
        self.r_point = r_point
        self.hashed_to_curve_r = hashed_to_curve_r


class SecretKey:
    """k256::SecretKey as the reference uses it: a scalar in [1, n-1] (`from_bytes` rejects the rest);
    `random(rng)` = 32 bytes from rng.fill_bytes, big-endian, rejection-sampled (pinned by the mock
    RNG of rust-k256/tests/signing.rs:23-44)."""

    def __init__(self, value):
        if not (1 <= value < ORDER):
            raise ValueError("secret key out of range")
        self.value = value

    @classmethod
    def from_bytes(cls, b):
        return cls(int.from_bytes(bytes(b), "big"))

    @classmethod
    def random(cls, rng):
        while True:
            buf = bytearray(32)
            rng.fill_bytes(buf)
            v = int.from_bytes(buf, "big")
            if 1 <= v < ORDER:
                return cls(v)

    def to_bytes(self):
        return self.value.to_bytes(32, "big")


class PlumeSigner:
    """rust-k256/src/randomizedsigner.rs:25-41: a borrowed secret key + the variant flag."""

    def __init__(self, secret_key, v1, ctx=None):
        self.secret_key = secret_key
        self.v1 = v1
        self._ctx = ctx

    def try_sign_with_rng(self, rng, msg):
        """randomizedsigner.rs:43-112 as a batch of one."""
        r = SecretKey.random(rng)                                           # :49
        ctx = self._ctx or default_context()
        o = ctx.sign_batch(1 if self.v1 else 2, [bytes(msg)], self.secret_key.to_bytes(), r.to_bytes())
        st = int(o["status"][0])
        if st != 0:
            raise PlumeError(STATUS_TEXT.get(st, "status %d" % st))         # the reference panics here
        v1f = None
        if self.v1:
            v1f = PlumeSignatureV1Fields(point_from_bytes(o["r_point"][0]), point_from_bytes(o["hashed_to_curve_r"][0]))
        return PlumeSignature(bytes(msg), point_from_bytes(o["pk"][0]), point_from_bytes(o["nullifier"][0]),
                              int.from_bytes(bytes(o["c"][0]), "big"), int.from_bytes(bytes(o["s"][0]), "big"), v1f, ctx=ctx)

    sign_with_rng = try_sign_with_rng


class PlumeSignature:
    """rust-k256/src/lib.rs:67-80"""

    def __init__(self, message, pk, nullifier, c, s, v1specific=None, ctx=None):
        self.message = bytes(message)
        self.pk = pk
        self.nullifier = nullifier
        self.c = c
        self.s = s
        self.v1specific = v1specific
        self._ctx = ctx

    def verify(self):
        """rust-k256/src/lib.rs:93-145 as a batch of one."""
        ctx = self._ctx or default_context()
        v1 = self.v1specific
        ok = ctx.verify_batch(1 if v1 else 2, [self.message], point_to_bytes(self.pk), point_to_bytes(self.nullifier),
                              self.c.to_bytes(32, "big"), self.s.to_bytes(32, "big"),
                              point_to_bytes(v1.r_point) if v1 else None,
                              point_to_bytes(v1.hashed_to_curve_r) if v1 else None)
        return bool(ok[0])

    # ---- wire form: the serde derive of rust-k256/src/lib.rs:66,83 through serde_json ---------------------------------
    # The struct derives Serialize/Deserialize behind the default feature `serde` (Cargo.toml:26-28 turns on k256/serde).
    # The field encodings come from k256 0.13 / elliptic-curve 0.13 (external crates; NO reference test pins them -- SURVEY.md
    # 8c "parity unpinned"): in a human-readable format AffinePoint is the upper-case hex of its SEC1 *compressed* encoding
    # and NonZeroScalar the upper-case hex of its 32 big-endian bytes (serdect), `message: Vec<u8>` is an array of numbers,
    # `v1specific` is an object or null.  Field order is the declaration order.
    def to_json(self):
        import json
        hexu = lambda b: bytes(b).hex().upper()
        d = {"message": list(self.message), "pk": hexu(encode_pt(self.pk)), "nullifier": hexu(encode_pt(self.nullifier)),
             "c": hexu(self.c.to_bytes(32, "big")), "s": hexu(self.s.to_bytes(32, "big")), "v1specific": None}
        if self.v1specific is not None:
            d["v1specific"] = {"r_point": hexu(encode_pt(self.v1specific.r_point)),
                               "hashed_to_curve_r": hexu(encode_pt(self.v1specific.hashed_to_curve_r))}
        return json.dumps(d, separators=(",", ":"))

    @classmethod
    def from_json(cls, text, ctx=None):
        """Inverse of to_json.  Points are decompressed on the GPU (plume_points_decompress_batch); a point that k256's
        AffinePoint decoding would reject, or a scalar outside [1, n-1] (NonZeroScalar), raises ValueError."""
        import json
        d = json.loads(text)
        ctx = ctx or default_context()
        names = ["pk", "nullifier"]
        enc = [bytes.fromhex(d["pk"]), bytes.fromhex(d["nullifier"])]
        v1 = d.get("v1specific")
        if v1 is not None:
            names += ["r_point", "hashed_to_curve_r"]
            enc += [bytes.fromhex(v1["r_point"]), bytes.fromhex(v1["hashed_to_curve_r"])]
        slots = np.zeros((len(enc), 33), dtype=np.uint8)
        for i, e in enumerate(enc):
            if not (len(e) == 33 or e == b"\x00"):
                raise ValueError("%s is not a SEC1 compressed point" % names[i])
            slots[i, :len(e)] = np.frombuffer(e, dtype=np.uint8)
        pts, ok = ctx.points_decompress(slots)
        if not ok.all():
            raise ValueError("%s does not decode to a curve point" % names[int(np.argmin(ok))])
        sc = {}
        for k in ("c", "s"):
            b = bytes.fromhex(d[k])
            if len(b) != 32 or not (1 <= int.from_bytes(b, "big") < ORDER):
                raise ValueError("%s is not a NonZeroScalar" % k)
            sc[k] = int.from_bytes(b, "big")
        p = [point_from_bytes(x) for x in pts]
        v1f = PlumeSignatureV1Fields(p[2], p[3]) if v1 is not None else None
        return cls(bytes(d["message"]), p[0], p[1], sc["c"], sc["s"], v1f, ctx=ctx)

    @staticmethod
    def sign_v1(secret_key, msg, rng, ctx=None):
        """rust-k256/src/lib.rs:149-151"""
        return PlumeSigner(secret_key, True, ctx).sign_with_rng(rng, msg)

    @staticmethod
    def sign_v2(secret_key, msg, rng, ctx=None):
        """rust-k256/src/lib.rs:154-156"""
        return PlumeSigner(secret_key, False, ctx).sign_with_rng(rng, msg)


def hash_to_curve(m, pk, ctx=None):
    """rust-k256/src/utils.rs:11-20"""
    ctx = ctx or default_context()
    slot = encode_pt(pk).ljust(33, b"\x00")
    return point_from_bytes(ctx.hash_to_curve_pk_batch([bytes(m)], slot)[0])

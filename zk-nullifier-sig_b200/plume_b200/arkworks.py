"""Host mirror of the twin crate `plume_arkworks` (rust-arkworks/src/lib.rs) over plume_ark_sign_batch /
plume_ark_verify_batch: same names, argument meaning and error behaviour, each call a batch of one; the batch
entry points are PlumeContext.ark_sign_batch / ark_verify_batch.  SURVEY.md 8f-3."""
from .api import ORDER, PlumeError, default_context, point_from_bytes, point_to_bytes


class PlumeVersion:
    """rust-arkworks/src/lib.rs:65-69"""
    V1 = 1
    V2 = 2


class HashToCurveError(PlumeError):
    """ark_ec::hashing::HashToCurveError as `hash_to_curve` raises it for the identity pk (lib.rs:97-100)."""


class PlumeSignaturePublic:
    """rust-arkworks/src/lib.rs:175-183"""

    def __init__(self, message, s, nullifier, variant=None):
        self.message, self.s, self.nullifier, self.variant = bytes(message), s, nullifier, variant


class PlumeSignaturePrivate:
    """rust-arkworks/src/lib.rs:185-194 (the witness: keep it secret)"""

    def __init__(self, hashed_to_curve_r, r_point, digest_private, variant):
        self.hashed_to_curve_r, self.r_point, self.digest_private, self.variant = hashed_to_curve_r, r_point, digest_private, variant


def _fr(x):
    return (x % ORDER).to_bytes(32, "big")


def sign_with_r(keypair, message, r_scalar, version, ctx=None):
    """rust-arkworks/src/lib.rs:229-278.  keypair = (pk point or None, sk int); scalars are taken mod n like Fr."""
    ctx = ctx or default_context()
    pk, sk = keypair
    o = ctx.ark_sign_batch(version, [bytes(message)], point_to_bytes(pk), _fr(sk), _fr(r_scalar))
    st = int(o["status"][0])
    if st == 6:
        raise HashToCurveError("`pk` shouldn't be the identity element")
    if st != 0:
        raise PlumeError("status %d" % st)
    pub = PlumeSignaturePublic(message, int.from_bytes(bytes(o["s"][0]), "big"), point_from_bytes(o["nullifier"][0]), version)
    priv = PlumeSignaturePrivate(point_from_bytes(o["hashed_to_curve_r"][0]), point_from_bytes(o["r_point"][0]),
                                 int.from_bytes(bytes(o["digest_private"][0]), "big"), version)
    return pub, priv


def sign(rng, keypair, message, version, ctx=None):
    """rust-arkworks/src/lib.rs:281-291: r = Fr::rand(rng) on the host, then sign_with_r.  `rng.fill_bytes(buf)`
    supplies the bytes; uniform in [0, n) by rejection."""
    while True:
        buf = bytearray(32)
        rng.fill_bytes(buf)
        r = int.from_bytes(buf, "big")
        if r < ORDER:
            break
    return sign_with_r(keypair, message, r, version, ctx)


def verify_non_zk(sig, pk, message, version, ctx=None):
    """rust-arkworks/src/tests.rs:28-78 (the generator parameter `pp` is the curve's G)."""
    ctx = ctx or default_context()
    pub, priv = sig
    if pk is None:
        raise HashToCurveError("`pk` shouldn't be the identity element")
    ok = ctx.ark_verify_batch(version, [bytes(message)], point_to_bytes(pk), point_to_bytes(pub.nullifier), _fr(priv.digest_private),
                              _fr(pub.s), point_to_bytes(priv.r_point), point_to_bytes(priv.hashed_to_curve_r))
    return bool(ok[0])

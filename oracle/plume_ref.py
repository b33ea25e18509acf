"""
oracle/plume_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Pure-Python (arbitrary precision int) restatement of the PLUME sign/verify hot path of
the reference crate `plume_rustcrypto` (rust-k256).  It exists to (1) pin the algorithm
against every golden vector the reference's own tests hold, and (2) cross-check the faster
C oracle (oracle/plume_oracle.c) on random inputs.  It is slow (~10 ms / signature) and is
only used on small cases.

The arithmetic the reference calls lives in third-party crates that are NOT vendored under
/root/reference: k256 ~0.13.3 (+ elliptic-curve 0.13.x, sha2 0.10.x), rust-k256/Cargo.toml:18.
What is restated here is their *published* algorithm: RFC 9380 suite
secp256k1_XMD:SHA-256_SSWU_RO_ and SEC1 point compression; parity is anchored on the
reference's call sites and golden vectors (tests/golden/reference_vectors.json).

Citations are relative to /root/reference/.
"""
import hashlib

# --- curve constants: rust-arkworks/src/secp256k1/fields/fq.rs:12, fields/fr.rs:19,
#     curves/mod.rs:39 (b=7), :50-58 (G)
P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
B = 7
GX = 55066263022277343669578718895168534326250603453777594175500187360389116729240
GY = 32670510020758816978083085130507043184471273380659243275938904335757337482424
G = (GX, GY)

# --- rust-k256/src/lib.rs:61
DST = b"QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_"

# --- isogenous curve E' and SSWU parameter: rust-arkworks/src/secp256k1/curves/mod.rs:71-73, :80
ISO_A = 0x3F8731ABDD661ADCA08A5558F0F5D272E953D363CB6F0E5D405447C01A444533
ISO_B = 1771
Z = (-11) % P

# --- 3-isogeny coefficients, ascending powers of x': curves/mod.rs:87-112
XNUM = [
    0x8E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38DAAAAA8C7,
    0x07D3D4C80BC321D5B9F315CEA7FD44C5D595D2FC0BF63B92DFFF1044F17C6581,
    0x534C328D23F234E6E2A413DECA25CAECE4506144037C40314ECBD0B53D9DD262,
    0x8E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38DAAAAA88C,
]
XDEN = [
    0xD35771193D94918A9CA34CCBB7B640DD86CD409542F8487D9FE6B745781EB49B,
    0xEDADC6F64383DC1DF7C4B2D51B54225406D36B641F5E41BBC52A56612A8C6D14,
    1,
    0,
]
YNUM = [
    0x4BDA12F684BDA12F684BDA12F684BDA12F684BDA12F684BDA12F684B8E38E23C,
    0xC75E0C32D5CB7C0FA9D0A54B12A0A6D5647AB046D686DA6FDFFC90FC201D71A3,
    0x29A6194691F91A73715209EF6512E576722830A201BE2018A765E85A9ECEE931,
    0x2F684BDA12F684BDA12F684BDA12F684BDA12F684BDA12F684BDA12F38E38D84,
]
YDEN = [
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFF93B,
    0x7A06534BB8BDB49FD5E9E6632722C2989467C1BFC8E8D978DFB425D2685C2573,
    0x6484AA716545CA2CF3A70C3FA8FE337E0A3D21162F0D6299A7BF8192BFD2A76F,
    1,
]

INF = None  # identity


def inv(a, m=P):
    return pow(a, -1, m)


def pt_add(p1, p2):
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * inv(2 * y1) % P
    else:
        lam = (y2 - y1) * inv(x2 - x1) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def pt_neg(p):
    return INF if p is INF else (p[0], (-p[1]) % P)


def pt_mul(p, k):
    k %= N
    acc = INF
    while k:
        if k & 1:
            acc = pt_add(acc, p)
        p = pt_add(p, p)
        k >>= 1
    return acc


def on_curve(p):
    return p is INF or (p[1] * p[1] - p[0] ** 3 - B) % P == 0


def encode_pt(p):
    """SEC1 compressed; identity -> single 0x00 byte.
    rust-k256/src/utils.rs:23-25 (`to_encoded_point(true)`); identity encoding pinned for the
    arkworks twin by rust-arkworks/src/tests/test_vectors.rs:3-7."""
    if p is INF:
        return b"\x00"
    return bytes([2 + (p[1] & 1)]) + p[0].to_bytes(32, "big")


def encode_pt_uncompressed(p):
    if p is INF:
        return b"\x00"
    return b"\x04" + p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def expand_message_xmd(msg, dst, n):
    """rust-arkworks/src/fixed_hasher/expander.rs:89-135 (SHA-256: b_len 32, block 64)."""
    ell = (n + 31) // 32
    assert ell <= 255 and n < 65536 and len(dst) <= 255
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + n.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(a ^ b for a, b in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:n]


def hash_to_field2(msg, dst=DST):
    """rust-arkworks/src/fixed_hasher/mod.rs:32-62: L = ceil((256+128)/8) = 48, N = 2."""
    ub = expand_message_xmd(msg, dst, 96)
    return [int.from_bytes(ub[0:48], "big") % P, int.from_bytes(ub[48:96], "big") % P]


def sgn0(x):
    return x & 1


def is_square(x):
    return x == 0 or pow(x, (P - 1) // 2, P) == 1


def sqrt(x):
    r = pow(x, (P + 1) // 4, P)
    assert r * r % P == x
    return r


def map_to_curve_sswu(u):
    """RFC 9380 6.6.2 simplified SWU on E' (A', B', Z=-11)."""
    A, Bp = ISO_A, ISO_B
    tv1 = (Z * Z * pow(u, 4, P) + Z * u * u) % P
    if tv1 == 0:
        x1 = Bp * inv(Z * A) % P
    else:
        x1 = (-Bp) * inv(A) % P * (1 + inv(tv1)) % P
    gx1 = (pow(x1, 3, P) + A * x1 + Bp) % P
    x2 = Z * u * u % P * x1 % P
    gx2 = (pow(x2, 3, P) + A * x2 + Bp) % P
    if is_square(gx1):
        x, y = x1, sqrt(gx1)
    else:
        x, y = x2, sqrt(gx2)
    if sgn0(u) != sgn0(y):
        y = (-y) % P
    return (x, y)


def iso_map(pt):
    """3-isogeny E' -> secp256k1, RFC 9380 appendix E.1; coefficients curves/mod.rs:87-112."""
    x, y = pt

    def ev(c):
        return (c[0] + c[1] * x + c[2] * x * x + c[3] * x * x * x) % P

    xn, xd, yn, yd = ev(XNUM), ev(XDEN), ev(YNUM), ev(YDEN)
    if xd == 0 or yd == 0:
        return INF
    return (xn * inv(xd) % P, y * yn % P * inv(yd) % P)


def hash_to_curve_bytes(msg, dst=DST):
    """`Secp256k1::hash_from_bytes::<ExpandMsgXmd<Sha256>>(&[msg], &[DST])`
    (rust-k256/src/utils.rs:15, randomizedsigner.rs:58)."""
    u0, u1 = hash_to_field2(msg, dst)
    q0 = iso_map(map_to_curve_sswu(u0))
    q1 = iso_map(map_to_curve_sswu(u1))
    return pt_add(q0, q1)  # cofactor 1


def hash_to_curve(m, pk):
    """rust-k256/src/utils.rs:11-20: h = H2C(m || enc33(pk))."""
    return hash_to_curve_bytes(m + encode_pt(pk))


def c_sha256_vec_signal(points):
    """rust-k256/src/lib.rs:159-168."""
    return hashlib.sha256(b"".join(encode_pt(p) for p in points)).digest()


# status codes of the batch ABI (include/plume_b200.h)
ST_OK, ST_BAD_R, ST_BAD_SK, ST_BAD_C, ST_ZERO_S, ST_H_INF = 0, 1, 2, 3, 4, 5


def sign(version, msg, sk, r):
    """rust-k256/src/randomizedsigner.rs:43-112 with r supplied (the reference draws it from
    the rng at :49; tests/signing.rs:23-44 mocks the rng with fixed bytes).
    Returns (status, dict).  The reference panics where status != 0."""
    if not (1 <= r < N):
        return ST_BAD_R, None
    if not (1 <= sk < N):
        return ST_BAD_SK, None
    r_point = pt_mul(G, r)                       # :51
    pk = pt_mul(G, sk)                           # :53
    pk_bytes = encode_pt(pk)                     # :54
    h = hash_to_curve_bytes(msg + pk_bytes)      # :57-61
    if h is INF:
        return ST_H_INF, None
    z = pt_mul(h, r)                             # :67
    nul = pt_mul(h, sk)                          # :70
    if version == 1:                             # :73-89
        c = c_sha256_vec_signal([G, pk, h, nul, r_point, z])
    else:
        c = c_sha256_vec_signal([nul, r_point, z])
    ci = int.from_bytes(c, "big")
    if not (1 <= ci < N):                        # :90-91 from_repr -> panic
        return ST_BAD_C, None
    s = (r + ci * sk) % N                        # :94
    if s == 0:                                   # :95
        return ST_ZERO_S, None
    return ST_OK, dict(pk=pk, nullifier=nul, c=ci, s=s, r_point=r_point, hashed_to_curve_r=z, h=h)


def verify(version, msg, pk, nul, c, s, r_point=None, hashed_to_curve_r=None):
    """rust-k256/src/lib.rs:93-145.  Points are affine tuples or INF; c, s ints in [1, n)."""
    rp = pt_add(pt_mul(G, s), pt_neg(pt_mul(pk, c)))            # :101
    h = hash_to_curve(msg, pk)                                  # :103
    zp = pt_add(pt_mul(h, s), pt_neg(pt_mul(nul, c)))           # :109
    if version == 1:
        if rp != r_point:                                       # :117
            return False
        if zp != hashed_to_curve_r:                             # :122
            return False
        d = c_sha256_vec_signal([G, pk, h, nul, rp, zp])        # :127-135
    else:
        d = c_sha256_vec_signal([nul, rp, zp])                  # :138-143
    return c == int.from_bytes(d, "big") % N                    # Scalar::reduce


# ---- the arkworks twin (rust-arkworks/src/lib.rs) ------------------------------------------------------
ST_BAD_PK = 6


def ark_sign_with_r(version, msg, pk, sk, r):
    """rust-arkworks/src/lib.rs:229-278.  pk: affine tuple (INF makes hash_to_curve fail, :97-100); sk, r: Fr, i.e.
    ints in [0, n).  Returns (status, dict): nothing else is rejected, c is reduced mod n (:257)."""
    if not (0 <= r < N):
        return ST_BAD_R, None
    if not (0 <= sk < N):
        return ST_BAD_SK, None
    if pk is INF:
        return ST_BAD_PK, None
    r_point = pt_mul(G, r) if r else INF                         # :235
    h = hash_to_curve(msg, pk)                                   # :238
    z = pt_mul(h, r) if r else INF                               # :241
    nul = pt_mul(h, sk) if sk else INF                           # :244
    if version == 1:                                             # :247-256 (identity -> the byte 00, :112-118)
        c = c_sha256_vec_signal([G, pk, h, nul, r_point, z])
    else:
        c = c_sha256_vec_signal([nul, r_point, z])
    ci = int.from_bytes(c, "big") % N                            # :257
    s = (r + sk * ci) % N                                        # :259-260
    return ST_OK, dict(nullifier=nul, digest_private=ci, s=s, r_point=r_point, hashed_to_curve_r=z, h=h)


def ark_verify_non_zk(version, msg, pk, nul, digest_private, s, r_point, hashed_to_curve_r):
    """rust-arkworks/src/tests.rs:28-78: Ok(bool); an identity pk is Err (returned as None here)."""
    if pk is INF:
        return None
    h = hash_to_curve(msg, pk)                                                       # :36
    if version == 1:                                                                 # :40-52
        c = c_sha256_vec_signal([G, pk, h, nul, r_point, hashed_to_curve_r])
    else:
        c = c_sha256_vec_signal([nul, r_point, hashed_to_curve_r])
    ci = int.from_bytes(c, "big") % N                                                # :53
    gs = pt_mul(G, s) if s else INF
    pkc = pt_mul(pk, digest_private) if digest_private else INF
    if r_point != pt_add(gs, pt_neg(pkc)):                                           # :56-62
        return False
    hs = pt_mul(h, s) if s and h is not INF else INF
    nc = pt_mul(nul, digest_private) if digest_private and nul is not INF else INF
    if hashed_to_curve_r != pt_add(hs, pt_neg(nc)):                                  # :65-71
        return False
    return ci == digest_private                                                      # :74


def h2c_witness(msg, dst=DST):
    """The intermediates of hash_to_curve a circuit-input generator needs: u0, u1; per u whether
    g(x1) is a square (RFC 9380 6.6.2 step 5: x = x1 if so, else x2); Q_k = iso_map(map_to_curve(u_k)); h = Q0 + Q1."""
    us = hash_to_field2(msg, dst)
    flags, qs = [], []
    for u in us:
        tv1 = (Z * Z * pow(u, 4, P) + Z * u * u) % P
        x1 = ISO_B * inv(Z * ISO_A) % P if tv1 == 0 else (-ISO_B) * inv(ISO_A) % P * (1 + inv(tv1)) % P
        flags.append(1 if is_square((pow(x1, 3, P) + ISO_A * x1 + ISO_B) % P) else 0)
        qs.append(iso_map(map_to_curve_sswu(u)))
    return us, flags, qs, pt_add(qs[0], qs[1])


def h2c_sqrt_hints(u):
    """Square-root hints of the circom hash_to_curve component for one u (circuits/circom/verify_nullifier.circom:21-31:
    q{0,1}_gx1_sqrt, q{0,1}_gx2_sqrt, q{0,1}_y_pos).  Their generator (npm secp256k1_hash_to_curve_circom) is not in the
    reference tree, so the convention is the one include/plume_b200.h DECLARES: exactly one of g(x1), g(x2) is a square
    (Z = -11 is not); that one's hint is its even square root (RFC 9380 sgn0 = 0), the other hint is 0; y_pos is the even
    square root of g(x) for the x the map takes.  Returns (gx1_sqrt, gx2_sqrt, y_pos) and the relations' inputs
    (x1, gx1, x2, gx2) so that tests can check what the circuit enforces."""
    A, Bp = ISO_A, ISO_B
    tv1 = (Z * Z * pow(u, 4, P) + Z * u * u) % P
    x1 = Bp * inv(Z * A) % P if tv1 == 0 else (-Bp) * inv(A) % P * (1 + inv(tv1)) % P
    gx1 = (pow(x1, 3, P) + A * x1 + Bp) % P
    x2 = Z * u * u % P * x1 % P
    gx2 = (pow(x2, 3, P) + A * x2 + Bp) % P

    def even_root(v):
        r = sqrt(v)
        return r if r % 2 == 0 else P - r

    if is_square(gx1):
        r = even_root(gx1)
        return (r, 0, r), (x1, gx1, x2, gx2)
    r = even_root(gx2)
    return (0, r, r), (x1, gx1, x2, gx2)


def registers(value, bits=64, count=4):
    """circuits/circom/utils.ts:32-51 bigIntToRegisters."""
    assert value < (1 << (bits * count))
    return [(value >> (bits * i)) & ((1 << bits) - 1) for i in range(count)]


def compress33(p):
    """33-byte SEC1 slot: 02/03 || x, identity = 00 followed by zeros."""
    return bytes(33) if p is INF else encode_pt(p)


def decompress33(b):
    """-> (point or INF, ok) with k256's acceptance rule: prefix 02/03, x < p, x^3 + 7 a square."""
    b = bytes(b)
    if b[0] == 0:
        return INF, all(v == 0 for v in b[1:])
    x = int.from_bytes(b[1:], "big")
    if b[0] not in (2, 3) or x >= P:
        return INF, False
    rhs = (x * x * x + 7) % P
    if not is_square(rhs):
        return INF, False
    y = sqrt(rhs)
    if (y & 1) != (b[0] & 1):
        y = P - y
    return (x, y), True

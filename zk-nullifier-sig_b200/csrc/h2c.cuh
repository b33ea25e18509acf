// h2c.cuh -- RFC 9380 hash_to_curve, suite secp256k1_XMD:SHA-256_SSWU_RO_, with the PLUME DST.
//
// Replaces `Secp256k1::hash_from_bytes::<ExpandMsgXmd<Sha256>>(&[m || enc33(pk)], &[DST])`
// (rust-k256/src/utils.rs:11-20, rust-k256/src/randomizedsigner.rs:57-61).  The in-repo spelled-out
// equivalents this follows: expand_message_xmd rust-arkworks/src/fixed_hasher/expander.rs:89-135,
// hash_to_field (L = 48, two elements) rust-arkworks/src/fixed_hasher/mod.rs:32-62, curve and map
// constants rust-arkworks/src/secp256k1/curves/mod.rs:71-112 (A', B', Z = -11, 3-isogeny).
//
// Everything stays projective: the SSWU output is (xn : xd, y), the isogeny is evaluated on the
// homogenised polynomials and lands directly in Jacobian coordinates, Q0 + Q1 is a Jacobian
// addition; the single inversion per point is left to the batched inversion kernel.
#pragma once
#include "ec.cuh"
#include "sha256.cuh"

// DST' = DST || len(DST) = 50 bytes (rust-k256/src/lib.rs:61)
PLUME_DEV uint8_t h2c_dst_prime(int i) {
    const uint8_t D[50] = {'Q', 'U', 'U', 'X', '-', 'V', '0', '1', '-', 'C', 'S', '0', '2', '-', 'w', 'i', 't', 'h', '-',
                           's', 'e', 'c', 'p', '2', '5', '6', 'k', '1', '_', 'X', 'M', 'D', ':', 'S', 'H', 'A', '-', '2',
                           '5', '6', '_', 'S', 'S', 'W', 'U', '_', 'R', 'O', '_', 49};
    return D[i];
}

// b_i = SHA-256(x(32 bytes) || idx || DST'), x given as 8 big-endian words: 83 bytes = 2 blocks
PLUME_DEV void h2c_hash_bi(uint32_t* out, const uint32_t* x, uint32_t idx) {
    uint32_t st[8], w[16];
    sha256_init(st);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = x[i];
    // bytes 32..63: idx, DST'[0..30]
    w[8] = (idx << 24) | ((uint32_t)h2c_dst_prime(0) << 16) | ((uint32_t)h2c_dst_prime(1) << 8) | h2c_dst_prime(2);
#pragma unroll
    for (int i = 9; i < 16; i++) {
        const int o = 3 + 4 * (i - 9);
        w[i] = ((uint32_t)h2c_dst_prime(o) << 24) | ((uint32_t)h2c_dst_prime(o + 1) << 16) | ((uint32_t)h2c_dst_prime(o + 2) << 8) | h2c_dst_prime(o + 3);
    }
    sha256_compress(st, w);
    // second block: DST'[31..49] (19 bytes), 0x80, zeros, bit length 83*8 = 664
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int o = 4 * i + b;  // byte offset inside the block
            uint32_t byte = (o < 19) ? h2c_dst_prime(31 + o) : (o == 19 ? 0x80u : 0u);
            v = (v << 8) | byte;
        }
        w[i] = v;
    }
    w[15] = 664;
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = st[i];
}

// 48 big-endian bytes (12 words, most significant first) mod p  -> fe
PLUME_DEV fe h2c_os2ip_mod_p(const uint32_t* w12) {
    // value = hi(128 bits) * 2^256 + lo(256 bits); 2^256 = C mod p
    uint32_t T[16];
#pragma unroll
    for (int i = 0; i < 8; i++) T[i] = w12[11 - i];
#pragma unroll
    for (int i = 0; i < 4; i++) T[8 + i] = w12[3 - i];
#pragma unroll
    for (int i = 12; i < 16; i++) T[i] = 0;
    return fe_norm(fe_reduce512(T));
}

// b_0 for the 65-byte preimage every PLUME call hashes (32-byte message || 33-byte SEC1 public key, or a 65-byte record
// handed to the hash_to_curve entry point): the stream after the Z_pad block is
//     preimage(65) || 00 60 00 || DST'(50)   = 118 bytes = two blocks with the padding,
// so the 32 message words are assembled in registers at constant offsets -- no byte buffer in local memory, no per-byte
// bookkeeping (the byte-stream hasher below spends ~8 000 instructions on these 118 bytes; this path ~150 plus the two
// compressions).  M: the preimage as 16 big-endian words, last: its 65th byte.
PLUME_DEV void h2c_b0_fixed65(uint32_t* b0, const uint32_t* M, uint32_t last) {
    uint32_t st[8], w[16];
    sha256_init_after_zero_block(st);
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = M[i];
    sha256_compress(st, w);
    // second block: byte 64 of the preimage, 00 60 00, DST' (50 bytes at offsets 4..53), 0x80 at 54, zeros, bit length
    w[0] = (last << 24) | 0x006000u;
#pragma unroll
    for (int i = 1; i < 16; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int o = 4 * i + b - 4;   // offset inside DST'
            uint32_t byte = (o < 50) ? h2c_dst_prime(o) : (o == 50 ? 0x80u : 0u);
            v = (v << 8) | byte;
        }
        w[i] = v;
    }
    w[15] = (64 + 118) * 8;
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 8; i++) b0[i] = st[i];
}

// 32 or 65 consecutive bytes -> big-endian words (word loads when the address allows it, byte loads otherwise)
template <int NW>
PLUME_DEV void h2c_load_be_words(uint32_t* W, const uint8_t* p) {
#ifndef PLUME_HOSTSIM
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int i = 0; i < NW; i++) W[i] = bswap32(__ldg(q + i));
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < NW; i++)
        W[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
}

// uniform_bytes = expand_message_xmd(msg || tail, DST, 96) then two field elements.
// The hashed stream is  Z_pad(64) || msg(len) || extra(nextra) || 0x00 0x60 || 0x00 || DST'.
// `extra` is the SEC1 encoding of pk that PLUME appends to the message (33 bytes, or 1 byte for
// the identity, or nothing for the plain hash_to_curve entry point).
PLUME_DEV void h2c_hash_to_field(fe& u0, fe& u1, const uint8_t* msg, uint32_t len, const uint8_t* extra, uint32_t nextra) {
    uint32_t b0[8];
    if (len == 32 && nextra == 33) {          // m || enc33(pk): the shape of every BASELINE config
        uint32_t M[16];
        h2c_load_be_words<8>(M, msg);
#pragma unroll
        for (int i = 0; i < 8; i++)
            M[8 + i] = ((uint32_t)extra[4 * i] << 24) | ((uint32_t)extra[4 * i + 1] << 16) | ((uint32_t)extra[4 * i + 2] << 8) | extra[4 * i + 3];
        h2c_b0_fixed65(b0, M, extra[32]);
    } else if (len == 65 && nextra == 0) {    // a caller-assembled 65-byte preimage (plume_hash_to_curve_batch)
        uint32_t M[16];
        h2c_load_be_words<16>(M, msg);
        h2c_b0_fixed65(b0, M, msg[64]);
    } else {
        sha256_stream s;
        sha256_init_after_zero_block(s.st);
        s.fill = 0;
        s.total = 64;
        sha256_stream_bytes(s, msg, len);
        sha256_stream_bytes(s, extra, nextra);
        sha256_stream_byte(s, 0x00);
        sha256_stream_byte(s, 0x60);  // len_in_bytes = 96
        sha256_stream_byte(s, 0x00);
#pragma unroll 1
        for (int i = 0; i < 50; i++) sha256_stream_byte(s, h2c_dst_prime(i));
        sha256_stream_final(s, b0);
    }
    uint32_t ub[24];
    h2c_hash_bi(ub, b0, 1);
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = b0[i] ^ ub[i];
    h2c_hash_bi(ub + 8, x, 2);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = b0[i] ^ ub[8 + i];
    h2c_hash_bi(ub + 16, x, 3);
    u0 = h2c_os2ip_mod_p(ub);
    u1 = h2c_os2ip_mod_p(ub + 12);
}

PLUME_DEV fe h2c_iso_a() { return fe_lit(0x3F8731ABu, 0xDD661ADCu, 0xA08A5558u, 0xF0F5D272u, 0xE953D363u, 0xCB6F0E5Du, 0x405447C0u, 0x1A444533u); }
#define H2C_ISO_B 1771u
// c2 = sqrt(-Z) = sqrt(11)
PLUME_DEV fe h2c_sqrt_neg_z() { return fe_lit(0x31FDF302u, 0x724013E5u, 0x7AD13FB3u, 0x8F842AFEu, 0xEC184F00u, 0xA74789DDu, 0x286729C8u, 0x303C4A59u); }

// RFC 9380 F.2.1.2 (q = 3 mod 4): returns is_square(u/v) and y = sqrt(u/v) or sqrt(Z*u/v)
PLUME_DEV bool h2c_sqrt_ratio(fe& y, const fe& u, const fe& v) {
    fe tv1 = fe_sqr(v);
    fe tv2 = fe_mul(u, v);
    tv1 = fe_mul(tv1, tv2);
    fe y1 = fe_pow_pm3d4(tv1);
    y1 = fe_mul(y1, tv2);
    fe y2 = fe_mul(y1, h2c_sqrt_neg_z());
    fe tv3 = fe_mul(fe_sqr(y1), v);
    bool is_qr = fe_eq(tv3, u);
    y = fe_cmov(y2, y1, is_qr);
    return is_qr;
}

// simplified SWU on E': y^2 = x^3 + A'x + B' (RFC 9380 F.2, straight line); x = xn / xd
// root (optional): the square root of g(x) the map found, before the sign is matched to sgn0(u) -- sqrt(g(x1)) when g(x1) is a
// square, u^3 Z sqrt(Z g(x1)) = sqrt(g(x2)) otherwise (the circuit-input hints of SURVEY.md 8f-4 are made from it)
PLUME_DEV bool h2c_map_sswu(fe& xn, fe& xd, fe& y, const fe& u, fe* root = nullptr) {   // returns is_square(g(x1)): which candidate x was taken
    const fe A = h2c_iso_a();
    fe tv1 = fe_neg(fe_mul_small(fe_sqr(u), 11));  // Z * u^2, Z = -11
    fe tv2 = fe_add(fe_sqr(tv1), tv1);
    fe tv3 = fe_mul_small(fe_add(tv2, fe_one()), H2C_ISO_B);
    // tv4 = A * (tv2 != 0 ? -tv2 : Z)
    fe zc = fe_neg(fe_set_u32(11));
    fe tv4 = fe_mul(A, fe_cmov(fe_neg(tv2), zc, fe_is_zero(tv2)));
    fe t2 = fe_sqr(tv3);
    fe tv6 = fe_sqr(tv4);
    fe tv5 = fe_mul(A, tv6);
    t2 = fe_mul(fe_add(t2, tv5), tv3);
    tv6 = fe_mul(tv6, tv4);
    tv5 = fe_mul_small(tv6, H2C_ISO_B);
    t2 = fe_add(t2, tv5);
    fe x = fe_mul(tv1, tv3);
    fe y1;
    bool is_gx1_square = h2c_sqrt_ratio(y1, t2, tv6);
    fe yy = fe_mul(fe_mul(tv1, u), y1);
    x = fe_cmov(x, tv3, is_gx1_square);
    yy = fe_cmov(yy, y1, is_gx1_square);
    if (root) *root = yy;
    bool e1 = fe_is_odd(u) == fe_is_odd(yy);
    y = fe_cmov(fe_neg(yy), yy, e1);
    xn = x;
    xd = tv4;
    return is_gx1_square;
}

// 3-isogeny E' -> secp256k1 on x' = xn/xd (RFC 9380 E.1), result in Jacobian coordinates
PLUME_DEV jac h2c_iso_map(const fe& xn, const fe& xd, const fe& yp) {
    const fe k10 = fe_lit(0x8E38E38Eu, 0x38E38E38u, 0xE38E38E3u, 0x8E38E38Eu, 0x38E38E38u, 0xE38E38E3u, 0x8E38E38Du, 0xAAAAA8C7u);
    const fe k11 = fe_lit(0x07D3D4C8u, 0x0BC321D5u, 0xB9F315CEu, 0xA7FD44C5u, 0xD595D2FCu, 0x0BF63B92u, 0xDFFF1044u, 0xF17C6581u);
    const fe k12 = fe_lit(0x534C328Du, 0x23F234E6u, 0xE2A413DEu, 0xCA25CAECu, 0xE4506144u, 0x037C4031u, 0x4ECBD0B5u, 0x3D9DD262u);
    const fe k13 = fe_lit(0x8E38E38Eu, 0x38E38E38u, 0xE38E38E3u, 0x8E38E38Eu, 0x38E38E38u, 0xE38E38E3u, 0x8E38E38Du, 0xAAAAA88Cu);
    const fe k20 = fe_lit(0xD3577119u, 0x3D94918Au, 0x9CA34CCBu, 0xB7B640DDu, 0x86CD4095u, 0x42F8487Du, 0x9FE6B745u, 0x781EB49Bu);
    const fe k21 = fe_lit(0xEDADC6F6u, 0x4383DC1Du, 0xF7C4B2D5u, 0x1B542254u, 0x06D36B64u, 0x1F5E41BBu, 0xC52A5661u, 0x2A8C6D14u);
    const fe k30 = fe_lit(0x4BDA12F6u, 0x84BDA12Fu, 0x684BDA12u, 0xF684BDA1u, 0x2F684BDAu, 0x12F684BDu, 0xA12F684Bu, 0x8E38E23Cu);
    const fe k31 = fe_lit(0xC75E0C32u, 0xD5CB7C0Fu, 0xA9D0A54Bu, 0x12A0A6D5u, 0x647AB046u, 0xD686DA6Fu, 0xDFFC90FCu, 0x201D71A3u);
    const fe k32 = fe_lit(0x29A61946u, 0x91F91A73u, 0x715209EFu, 0x6512E576u, 0x722830A2u, 0x01BE2018u, 0xA765E85Au, 0x9ECEE931u);
    const fe k33 = fe_lit(0x2F684BDAu, 0x12F684BDu, 0xA12F684Bu, 0xDA12F684u, 0xBDA12F68u, 0x4BDA12F6u, 0x84BDA12Fu, 0x38E38D84u);
    const fe k40 = fe_lit(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFEu, 0xFFFFF93Bu);
    const fe k41 = fe_lit(0x7A06534Bu, 0xB8BDB49Fu, 0xD5E9E663u, 0x2722C298u, 0x9467C1BFu, 0xC8E8D978u, 0xDFB425D2u, 0x685C2573u);
    const fe k42 = fe_lit(0x6484AA71u, 0x6545CA2Cu, 0xF3A70C3Fu, 0xA8FE337Eu, 0x0A3D2116u, 0x2F0D6299u, 0xA7BF8192u, 0xBFD2A76Fu);
    // homogenise with D = xd: X = xn
    fe X = xn, D = xd;
    fe X2 = fe_sqr(X), D2 = fe_sqr(D);
    fe X3 = fe_mul(X2, X), D3 = fe_mul(D2, D);
    fe XD = fe_mul(X, D);
    fe X2D = fe_mul(X2, D), XD2 = fe_mul(X, D2);
    // x_num / D^3, x_den / D^2, y_num / D^3, y_den / D^3
    fe nx = fe_add(fe_add(fe_mul(k13, X3), fe_mul(k12, X2D)), fe_add(fe_mul(k11, XD2), fe_mul(k10, D3)));
    fe dx = fe_add(fe_add(X2, fe_mul(k21, XD)), fe_mul(k20, D2));
    fe ny = fe_add(fe_add(fe_mul(k33, X3), fe_mul(k32, X2D)), fe_add(fe_mul(k31, XD2), fe_mul(k30, D3)));
    fe dy = fe_add(fe_add(X3, fe_mul(k42, X2D)), fe_add(fe_mul(k41, XD2), fe_mul(k40, D3)));
    // x = nx / (dx * D), y = y' * ny / dy
    fe Dx = fe_mul(dx, D);
    if (fe_is_zero(Dx) || fe_is_zero(dy)) return jac_infinity();
    // Jacobian with Z = Dx * dy: X = nx * Dx * dy^2, Y = y' * ny * Dx^3 * dy^2
    jac r;
    fe dy2 = fe_sqr(dy);
    fe t = fe_mul(Dx, dy2);            // Dx * dy^2
    r.x = fe_mul(nx, t);
    r.y = fe_mul(fe_mul(yp, ny), fe_mul(t, fe_sqr(Dx)));
    r.z = fe_mul(Dx, dy);
    r.inf = 0;
    return r;
}

// h = hash_to_curve(msg || extra), Jacobian (possibly the identity)
PLUME_DEV jac h2c_hash_to_curve(const uint8_t* msg, uint32_t len, const uint8_t* extra, uint32_t nextra) {
    fe u0, u1;
    h2c_hash_to_field(u0, u1, msg, len, extra, nextra);
    jac q[2];
#pragma unroll 1
    for (int i = 0; i < 2; i++) {
        fe xn, xd, y;
        h2c_map_sswu(xn, xd, y, i == 0 ? u0 : u1);
        q[i] = h2c_iso_map(xn, xd, y);
    }
    return jac_add(q[0], q[1]);
}

#ifndef PLUME_HOSTSIM
// Small batches: the two field elements of one item are mapped on two neighbouring lanes (j = 0, 1), which halves the serial
// work (two exponentiations of 254 squarings dominate hash_to_curve).  Both lanes hash (cheap) and both return Q0 + Q1.
PLUME_DEV jac h2c_hash_to_curve_team(uint32_t mask, uint32_t j, const uint8_t* msg, uint32_t len, const uint8_t* extra, uint32_t nextra) {
    fe u0, u1;
    h2c_hash_to_field(u0, u1, msg, len, extra, nextra);
    fe xn, xd, y;
    h2c_map_sswu(xn, xd, y, j == 0 ? u0 : u1);
    return jac_team_sum(mask, h2c_iso_map(xn, xd, y), j, 1);
}
#endif

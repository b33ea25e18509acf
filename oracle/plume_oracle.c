/*
 * oracle/plume_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the PLUME
 * sign / verify / hash_to_curve path of the reference crate `plume_rustcrypto` (rust-k256).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (libplume_b200.so) never links, imports or calls it.
 *
 * The arithmetic the reference calls is in third-party crates that are NOT vendored under
 * /root/reference: k256 ~0.13.3, elliptic-curve 0.13.x, sha2 0.10.x (rust-k256/Cargo.toml:18); no
 * Rust toolchain exists in the build image either, so the reference cannot be compiled here.
 * This file restates the published algorithms those crates implement (SEC1/SEC2 secp256k1,
 * FIPS 180-4 SHA-256, RFC 9380 suite secp256k1_XMD:SHA-256_SSWU_RO_) around the protocol exactly
 * as the reference's own files spell it.  PINNING: tests/test_oracle.py checks this oracle against
 * every golden vector the reference's tests hold for the path (tests/golden/reference_vectors.json,
 * extracted from rust-k256/tests/{signing,verification}.rs, rust-arkworks/src/tests.rs, .../secp256k1/tests.rs,
 * .../tests/test_vectors.rs, circuits/.../hashToCurve.test.ts) and against the independent
 * pure-Python oracle/plume_ref.py on random inputs.
 *
 * It is written independently of the CUDA code: 4 x 64-bit limbs with unsigned __int128, full
 * reduction after every field operation, width-5 wNAF scalar multiplication without endomorphism,
 * one field inversion per affine conversion -- roughly the work k256 does per signature, which
 * also makes it the "port" CPU baseline of bench.py.
 *
 * Citations are relative to /root/reference/.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;   /* element of Fp, always in [0, p) */
typedef struct { uint64_t v[4]; } sc;   /* integer mod n, always in [0, n) */
typedef struct { fe x, y; int inf; } aff;
typedef struct { fe x, y, z; int inf; } jac;

/* p = 2^256 - 2^32 - 977 (rust-arkworks/src/secp256k1/fields/fq.rs:12) */
static const uint64_t P[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
#define PC 0x1000003D1ULL /* 2^256 - p */
/* n (rust-arkworks/src/secp256k1/fields/fr.rs:19) */
static const uint64_t N[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
/* 2^256 - n */
static const uint64_t NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL};

/* ---------------------------------------------------------------- multiprecision helpers */
static int ge4(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t bo = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - bo;
        r[i] = (uint64_t)d;
        bo = (uint64_t)(d >> 64) & 1;
    }
    return bo;
}
static void mul4(uint64_t* t, const uint64_t* a, const uint64_t* b) {
    memset(t, 0, 8 * sizeof(uint64_t));
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a[i] * b[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + 4] = (uint64_t)c;
    }
}
static int is_zero4(const uint64_t* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static void from_be32(uint64_t* r, const uint8_t* b) {
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | b[8 * (3 - i) + k];
        r[i] = w;
    }
}
static void to_be32(uint8_t* b, const uint64_t* a) {
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) b[8 * (3 - i) + k] = (uint8_t)(a[i] >> (56 - 8 * k));
}

/* ---------------------------------------------------------------- Fp */
static void fe_set_u64(fe* r, uint64_t x) { r->v[0] = x; r->v[1] = r->v[2] = r->v[3] = 0; }
static int fe_is_zero(const fe* a) { return is_zero4(a->v); }
static int fe_eq(const fe* a, const fe* b) { return memcmp(a->v, b->v, 32) == 0; }
static int fe_is_odd(const fe* a) { return (int)(a->v[0] & 1); }
static void fe_add(fe* r, const fe* a, const fe* b) {
    uint64_t t[4];
    uint64_t c = add4(t, a->v, b->v);
    if (c || ge4(t, P)) sub4(t, t, P);
    memcpy(r->v, t, 32);
}
static void fe_sub(fe* r, const fe* a, const fe* b) {
    uint64_t t[4];
    if (sub4(t, a->v, b->v)) add4(t, t, P);
    memcpy(r->v, t, 32);
}
static void fe_neg(fe* r, const fe* a) {
    fe z;
    fe_set_u64(&z, 0);
    fe_sub(r, &z, a);
}
/* t (512 bits) mod p */
static void fe_reduce8(fe* r, const uint64_t* t) {
    /* lo + hi * PC, twice */
    uint64_t a[5];
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * PC + t[i]; a[i] = (uint64_t)c; c >>= 64; }
    a[4] = (uint64_t)c;
    uint64_t b[4];
    c = (u128)a[4] * PC;
    for (int i = 0; i < 4; i++) { c += a[i]; b[i] = (uint64_t)c; c >>= 64; }
    if ((uint64_t)c) {  /* one more wrap: value = b + 2^256 -> b + PC (cannot wrap again) */
        u128 d = PC;
        for (int i = 0; i < 4; i++) { d += b[i]; b[i] = (uint64_t)d; d >>= 64; }
    }
    if (ge4(b, P)) sub4(b, b, P);
    memcpy(r->v, b, 32);
}
static void fe_mul(fe* r, const fe* a, const fe* b) {
    uint64_t t[8];
    mul4(t, a->v, b->v);
    fe_reduce8(r, t);
}
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static void fe_sqrn(fe* r, const fe* a, int n) {
    *r = *a;
    for (int i = 0; i < n; i++) fe_sqr(r, r);
}
/* a^(2^223 - 1) plus the a^3 and a^(2^22 - 1) blocks, the common prefix of p - 2 and (p + 1)/4 whose
 * binary expansions are [223 ones][0][22 ones][0000] followed by 101101 resp. 1100 */
static void fe_pow_prefix(fe* x223, fe* x22, fe* x2, const fe* a) {
    fe x3, x6, x9, x11, x44, x88, x176, x220, t;
    fe_sqr(&t, a); fe_mul(x2, &t, a);
    fe_sqr(&t, x2); fe_mul(&x3, &t, a);
    fe_sqrn(&t, &x3, 3); fe_mul(&x6, &t, &x3);
    fe_sqrn(&t, &x6, 3); fe_mul(&x9, &t, &x3);
    fe_sqrn(&t, &x9, 2); fe_mul(&x11, &t, x2);
    fe_sqrn(&t, &x11, 11); fe_mul(x22, &t, &x11);
    fe_sqrn(&t, x22, 22); fe_mul(&x44, &t, x22);
    fe_sqrn(&t, &x44, 44); fe_mul(&x88, &t, &x44);
    fe_sqrn(&t, &x88, 88); fe_mul(&x176, &t, &x88);
    fe_sqrn(&t, &x176, 44); fe_mul(&x220, &t, &x44);
    fe_sqrn(&t, &x220, 3); fe_mul(x223, &t, &x3);
}
static void fe_inv(fe* r, const fe* a) {   /* a^(p-2) */
    fe x223, x22, x2, t;
    fe_pow_prefix(&x223, &x22, &x2, a);
    fe_sqrn(&t, &x223, 23); fe_mul(&t, &t, &x22);
    fe_sqrn(&t, &t, 5); fe_mul(&t, &t, a);
    fe_sqrn(&t, &t, 3); fe_mul(&t, &t, &x2);
    fe_sqrn(&t, &t, 2); fe_mul(r, &t, a);
}
static void fe_from_be(fe* r, const uint8_t* b) { from_be32(r->v, b); }
static void fe_to_be(uint8_t* b, const fe* a) { to_be32(b, a->v); }

/* ---------------------------------------------------------------- scalars mod n */
static void sc_reduce8(sc* r, const uint64_t* t) {
    /* fold the high half with 2^256 = NC (mod n) until it fits */
    uint64_t a[8];
    memcpy(a, t, sizeof(a));
    for (int round = 0; round < 3; round++) {
        uint64_t lo[8] = {a[0], a[1], a[2], a[3], 0, 0, 0, 0};
        uint64_t hi[4] = {a[4], a[5], a[6], a[7]};
        if (is_zero4(hi)) break;
        /* lo += hi * NC */
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 3; j++) {
                c += (u128)hi[i] * NC[j] + lo[i + j];
                lo[i + j] = (uint64_t)c;
                c >>= 64;
            }
            for (int k = i + 3; k < 8 && (uint64_t)c; k++) { c += lo[k]; lo[k] = (uint64_t)c; c >>= 64; }
        }
        memcpy(a, lo, sizeof(a));
    }
    /* now a < 2^256 + small; a[4] in {0,1} */
    while (a[4] || ge4(a, N)) {
        uint64_t bo = sub4(a, a, N);
        a[4] -= bo;
    }
    memcpy(r->v, a, 32);
}
static void sc_mul(sc* r, const sc* a, const sc* b) {
    uint64_t t[8];
    mul4(t, a->v, b->v);
    sc_reduce8(r, t);
}
static void sc_add(sc* r, const sc* a, const sc* b) {
    uint64_t t[8] = {0};
    t[4] = add4(t, a->v, b->v);
    sc_reduce8(r, t);
}
static void sc_neg(sc* r, const sc* a) {
    if (is_zero4(a->v)) { *r = *a; return; }
    sub4(r->v, N, a->v);
}
static int sc_valid_nonzero(const uint64_t* a) { return !is_zero4(a) && !ge4(a, N); }

/* ---------------------------------------------------------------- curve y^2 = x^3 + 7 */
/* generator: rust-arkworks/src/secp256k1/curves/mod.rs:50-58 */
static const uint8_t GX_BE[32] = {0x79, 0xBE, 0x66, 0x7E, 0xF9, 0xDC, 0xBB, 0xAC, 0x55, 0xA0, 0x62, 0x95, 0xCE, 0x87, 0x0B, 0x07,
                                  0x02, 0x9B, 0xFC, 0xDB, 0x2D, 0xCE, 0x28, 0xD9, 0x59, 0xF2, 0x81, 0x5B, 0x16, 0xF8, 0x17, 0x98};
static const uint8_t GY_BE[32] = {0x48, 0x3A, 0xDA, 0x77, 0x26, 0xA3, 0xC4, 0x65, 0x5D, 0xA4, 0xFB, 0xFC, 0x0E, 0x11, 0x08, 0xA8,
                                  0xFD, 0x17, 0xB4, 0x48, 0xA6, 0x85, 0x54, 0x19, 0x9C, 0x47, 0xD0, 0x8F, 0xFB, 0x10, 0xD4, 0xB8};
static void generator(aff* g) { fe_from_be(&g->x, GX_BE); fe_from_be(&g->y, GY_BE); g->inf = 0; }

static int on_curve(const aff* p) {
    if (p->inf) return 1;
    fe l, r, seven;
    fe_sqr(&l, &p->y);
    fe_sqr(&r, &p->x);
    fe_mul(&r, &r, &p->x);
    fe_set_u64(&seven, 7);
    fe_add(&r, &r, &seven);
    return fe_eq(&l, &r);
}
static void jac_set_inf(jac* r) { memset(r, 0, sizeof(*r)); r->inf = 1; }
static void jac_from_aff(jac* r, const aff* p) {
    r->x = p->x; r->y = p->y; fe_set_u64(&r->z, 1); r->inf = p->inf;
}
static void jac_dbl(jac* r, const jac* p) {
    if (p->inf || fe_is_zero(&p->y)) { jac_set_inf(r); return; }
    /* standard a = 0 doubling: S = 4XY^2, M = 3X^2, X' = M^2 - 2S, Y' = M(S - X') - 8Y^4, Z' = 2YZ */
    fe y2, s, m, t, x3, y3, z3;
    fe_sqr(&y2, &p->y);
    fe_mul(&s, &p->x, &y2); fe_add(&s, &s, &s); fe_add(&s, &s, &s);
    fe_sqr(&m, &p->x); fe_add(&t, &m, &m); fe_add(&m, &t, &m);
    fe_sqr(&x3, &m); fe_sub(&x3, &x3, &s); fe_sub(&x3, &x3, &s);
    fe_sqr(&t, &y2); fe_add(&t, &t, &t); fe_add(&t, &t, &t); fe_add(&t, &t, &t);
    fe_sub(&y3, &s, &x3); fe_mul(&y3, &y3, &m); fe_sub(&y3, &y3, &t);
    fe_mul(&z3, &p->y, &p->z); fe_add(&z3, &z3, &z3);
    r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}
static void jac_add(jac* r, const jac* p, const jac* q) {
    if (p->inf) { *r = *q; return; }
    if (q->inf) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, rr, h2, h3, v, t;
    fe_sqr(&z1z1, &p->z); fe_sqr(&z2z2, &q->z);
    fe_mul(&u1, &p->x, &z2z2); fe_mul(&u2, &q->x, &z1z1);
    fe_mul(&s1, &q->z, &z2z2); fe_mul(&s1, &s1, &p->y);
    fe_mul(&s2, &p->z, &z1z1); fe_mul(&s2, &s2, &q->y);
    fe_sub(&h, &u2, &u1); fe_sub(&rr, &s2, &s1);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&rr)) { jac_dbl(r, p); return; }
        jac_set_inf(r); return;
    }
    fe_sqr(&h2, &h); fe_mul(&h3, &h2, &h); fe_mul(&v, &u1, &h2);
    jac o;
    fe_sqr(&o.x, &rr); fe_sub(&o.x, &o.x, &h3); fe_sub(&o.x, &o.x, &v); fe_sub(&o.x, &o.x, &v);
    fe_sub(&t, &v, &o.x); fe_mul(&t, &t, &rr); fe_mul(&h3, &h3, &s1); fe_sub(&o.y, &t, &h3);
    fe_mul(&o.z, &p->z, &q->z); fe_mul(&o.z, &o.z, &h);
    o.inf = 0;
    *r = o;
}
static void jac_neg(jac* r, const jac* p) { *r = *p; fe_neg(&r->y, &p->y); }
static void jac_to_aff(aff* r, const jac* p) {
    if (p->inf) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
    fe zi, zi2, zi3;
    fe_inv(&zi, &p->z);  /* one inversion per conversion, like k256's to_affine */
    fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
    fe_mul(&r->x, &p->x, &zi2); fe_mul(&r->y, &p->y, &zi3);
    r->inf = 0;
}
/* k * P: width-5 wNAF, odd multiples 1P..15P in Jacobian form (`ProjectivePoint * Scalar`) */
static void jac_mul(jac* r, const aff* p, const sc* k) {
    if (p->inf || is_zero4(k->v)) { jac_set_inf(r); return; }
    jac tab[8], p2, base;
    jac_from_aff(&base, p);
    tab[0] = base;
    jac_dbl(&p2, &base);
    for (int i = 1; i < 8; i++) jac_add(&tab[i], &tab[i - 1], &p2);
    int8_t naf[260];
    int len = 0;
    uint64_t d[5] = {k->v[0], k->v[1], k->v[2], k->v[3], 0};
    while (d[0] | d[1] | d[2] | d[3] | d[4]) {
        int8_t z = 0;
        if (d[0] & 1) {
            int w = (int)(d[0] & 31);
            if (w > 16) w -= 32;
            z = (int8_t)w;
            /* d -= w */
            if (w > 0) {
                uint64_t bo = (uint64_t)w;
                for (int i = 0; i < 5 && bo; i++) { uint64_t o = d[i]; d[i] = o - bo; bo = o < bo; }
            } else {
                uint64_t c = (uint64_t)(-w);
                for (int i = 0; i < 5 && c; i++) { uint64_t o = d[i]; d[i] = o + c; c = d[i] < o; }
            }
        }
        naf[len++] = z;
        for (int i = 0; i < 4; i++) d[i] = (d[i] >> 1) | (d[i + 1] << 63);
        d[4] >>= 1;
    }
    jac acc;
    jac_set_inf(&acc);
    for (int i = len - 1; i >= 0; i--) {
        jac_dbl(&acc, &acc);
        if (naf[i] > 0) jac_add(&acc, &acc, &tab[(naf[i] - 1) / 2]);
        else if (naf[i] < 0) { jac t; jac_neg(&t, &tab[(-naf[i] - 1) / 2]); jac_add(&acc, &acc, &t); }
    }
    *r = acc;
}

/* wire helpers: 64-byte x||y, zeros = identity */
static int point_from_wire(aff* p, const uint8_t* b) {
    static const uint8_t zero[64] = {0};
    if (memcmp(b, zero, 64) == 0) { memset(p, 0, sizeof(*p)); p->inf = 1; return 1; }
    uint64_t x[4], y[4];
    from_be32(x, b); from_be32(y, b + 32);
    if (ge4(x, P) || ge4(y, P)) return 0;
    memcpy(p->x.v, x, 32); memcpy(p->y.v, y, 32); p->inf = 0;
    return on_curve(p);
}
static void point_to_wire(uint8_t* b, const aff* p) {
    if (p->inf) { memset(b, 0, 64); return; }
    fe_to_be(b, &p->x); fe_to_be(b + 32, &p->y);
}
/* SEC1 compressed; identity -> single 0x00 (rust-k256/src/utils.rs:23-25) */
static size_t encode_pt(uint8_t* out, const aff* p) {
    if (p->inf) { out[0] = 0; return 1; }
    out[0] = (uint8_t)(2 + fe_is_odd(&p->y));
    fe_to_be(out + 1, &p->x);
    return 33;
}

/* ---------------------------------------------------------------- SHA-256 (FIPS 180-4) */
typedef struct { uint32_t h[8]; uint8_t buf[64]; size_t fill; uint64_t total; } sha256_t;
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block(sha256_t* s, const uint8_t* b) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = s->h[0], bb = s->h[1], c = s->h[2], d = s->h[3], e = s->h[4], f = s->h[5], g = s->h[6], h = s->h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    s->h[0] += a; s->h[1] += bb; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}
static void sha256_init(sha256_t* s) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(s->h, iv, sizeof(iv));
    s->fill = 0; s->total = 0;
}
static void sha256_update(sha256_t* s, const uint8_t* p, size_t n) {
    s->total += n;
    while (n) {
        size_t k = 64 - s->fill;
        if (k > n) k = n;
        memcpy(s->buf + s->fill, p, k);
        s->fill += k; p += k; n -= k;
        if (s->fill == 64) { sha256_block(s, s->buf); s->fill = 0; }
    }
}
static void sha256_final(sha256_t* s, uint8_t* out) {
    uint64_t bits = s->total * 8;
    uint8_t pad[72] = {0x80};
    size_t padlen = (s->fill < 56) ? 56 - s->fill : 120 - s->fill;
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_update(s, pad, padlen);
    sha256_update(s, len, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(s->h[i] >> 24); out[4 * i + 1] = (uint8_t)(s->h[i] >> 16); out[4 * i + 2] = (uint8_t)(s->h[i] >> 8); out[4 * i + 3] = (uint8_t)s->h[i]; }
}

/* ---------------------------------------------------------------- hash to curve */
/* rust-k256/src/lib.rs:61 */
static const char DST[] = "QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_";
#define DST_LEN 49

/* rust-arkworks/src/fixed_hasher/expander.rs:89-135; msg given as two pieces (PLUME hashes m then
 * enc(pk): rust-k256/src/randomizedsigner.rs:58 passes them as two slices) */
static void expand_message_xmd(uint8_t* out, size_t n, const uint8_t* m1, size_t l1, const uint8_t* m2, size_t l2) {
    uint8_t dst_prime[DST_LEN + 1];
    memcpy(dst_prime, DST, DST_LEN);
    dst_prime[DST_LEN] = DST_LEN;
    size_t ell = (n + 31) / 32;
    uint8_t zpad[64] = {0}, lib[3] = {(uint8_t)(n >> 8), (uint8_t)n, 0};
    uint8_t b0[32], bi[32];
    sha256_t s;
    sha256_init(&s);
    sha256_update(&s, zpad, 64);
    sha256_update(&s, m1, l1);
    sha256_update(&s, m2, l2);
    sha256_update(&s, lib, 3);
    sha256_update(&s, dst_prime, sizeof(dst_prime));
    sha256_final(&s, b0);
    uint8_t one = 1;
    sha256_init(&s);
    sha256_update(&s, b0, 32);
    sha256_update(&s, &one, 1);
    sha256_update(&s, dst_prime, sizeof(dst_prime));
    sha256_final(&s, bi);
    size_t off = 0;
    for (size_t i = 1; i <= ell; i++) {
        size_t k = (n - off < 32) ? n - off : 32;
        memcpy(out + off, bi, k);
        off += k;
        if (i == ell) break;
        uint8_t x[32], idx = (uint8_t)(i + 1);
        for (int j = 0; j < 32; j++) x[j] = b0[j] ^ bi[j];
        sha256_init(&s);
        sha256_update(&s, x, 32);
        sha256_update(&s, &idx, 1);
        sha256_update(&s, dst_prime, sizeof(dst_prime));
        sha256_final(&s, bi);
    }
}
/* 48 big-endian bytes mod p (rust-arkworks/src/fixed_hasher/mod.rs:32-62, L = 48) */
static void fe_from_be48(fe* r, const uint8_t* b) {
    uint8_t pad[64] = {0};
    memcpy(pad + 16, b, 48);
    uint64_t t[8];
    from_be32(t + 4, pad);
    from_be32(t, pad + 32);
    fe_reduce8(r, t);
}
static void fe_from_hex(fe* r, const char* hex) {
    uint8_t b[32];
    for (int i = 0; i < 32; i++) {
        unsigned hi = (unsigned char)hex[2 * i], lo = (unsigned char)hex[2 * i + 1];
        hi = hi <= '9' ? hi - '0' : (hi | 32) - 'a' + 10;
        lo = lo <= '9' ? lo - '0' : (lo | 32) - 'a' + 10;
        b[i] = (uint8_t)(hi * 16 + lo);
    }
    fe_from_be(r, b);
}
/* is a a quadratic residue; if so *r = sqrt(a) */
static int fe_sqrt(fe* r, const fe* a) {
    /* a^((p + 1) / 4) */
    fe x223, x22, x2, c, c2;
    fe_pow_prefix(&x223, &x22, &x2, a);
    fe_sqrn(&c, &x223, 23); fe_mul(&c, &c, &x22);
    fe_sqrn(&c, &c, 6); fe_mul(&c, &c, &x2);
    fe_sqrn(&c, &c, 2);
    fe_sqr(&c2, &c);
    if (!fe_eq(&c2, a)) return 0;
    *r = c;
    return 1;
}
/* RFC 9380 6.6.2 simplified SWU on E': y^2 = x^3 + A'x + B', Z = -11
 * (rust-arkworks/src/secp256k1/curves/mod.rs:71-80) */
static void map_to_curve_sswu(aff* out, const fe* u) {
    fe A, B, Z, one, t1, t2, x1, x2, gx1, gx2, y;
    fe_from_hex(&A, "3f8731abdd661adca08a5558f0f5d272e953d363cb6f0e5d405447c01a444533");
    fe_set_u64(&B, 1771);
    fe_set_u64(&Z, 11); fe_neg(&Z, &Z);
    fe_set_u64(&one, 1);
    fe u2, zu2;
    fe_sqr(&u2, u); fe_mul(&zu2, &Z, &u2);         /* Z u^2 */
    fe_sqr(&t1, &zu2); fe_add(&t1, &t1, &zu2);      /* Z^2 u^4 + Z u^2 */
    if (fe_is_zero(&t1)) {
        fe_mul(&t2, &Z, &A); fe_inv(&t2, &t2); fe_mul(&x1, &B, &t2);           /* B / (Z A) */
    } else {
        fe_inv(&t2, &t1); fe_add(&t2, &t2, &one);                               /* 1 + 1/tv1 */
        fe nb, ai;
        fe_neg(&nb, &B); fe_inv(&ai, &A); fe_mul(&x1, &nb, &ai); fe_mul(&x1, &x1, &t2);  /* (-B/A)(1 + 1/tv1) */
    }
    fe_sqr(&gx1, &x1); fe_add(&gx1, &gx1, &A); fe_mul(&gx1, &gx1, &x1); fe_add(&gx1, &gx1, &B);
    fe_mul(&x2, &zu2, &x1);
    fe_sqr(&gx2, &x2); fe_add(&gx2, &gx2, &A); fe_mul(&gx2, &gx2, &x2); fe_add(&gx2, &gx2, &B);
    if (fe_sqrt(&y, &gx1)) out->x = x1;
    else { fe_sqrt(&y, &gx2); out->x = x2; }
    if (fe_is_odd(u) != fe_is_odd(&y)) fe_neg(&y, &y);
    out->y = y; out->inf = 0;
}
/* 3-isogeny E' -> E (RFC 9380 appendix E.1; rust-arkworks/src/secp256k1/curves/mod.rs:87-112) */
static void iso_map(aff* out, const aff* in) {
    static const char* K[4][4] = {
        {"8e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38daaaaa8c7", "07d3d4c80bc321d5b9f315cea7fd44c5d595d2fc0bf63b92dfff1044f17c6581",
         "534c328d23f234e6e2a413deca25caece4506144037c40314ecbd0b53d9dd262", "8e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38daaaaa88c"},
        {"d35771193d94918a9ca34ccbb7b640dd86cd409542f8487d9fe6b745781eb49b", "edadc6f64383dc1df7c4b2d51b54225406d36b641f5e41bbc52a56612a8c6d14",
         "0000000000000000000000000000000000000000000000000000000000000001", "0000000000000000000000000000000000000000000000000000000000000000"},
        {"4bda12f684bda12f684bda12f684bda12f684bda12f684bda12f684b8e38e23c", "c75e0c32d5cb7c0fa9d0a54b12a0a6d5647ab046d686da6fdffc90fc201d71a3",
         "29a6194691f91a73715209ef6512e576722830a201be2018a765e85a9ecee931", "2f684bda12f684bda12f684bda12f684bda12f684bda12f684bda12f38e38d84"},
        {"fffffffffffffffffffffffffffffffffffffffffffffffffffffffefffff93b", "7a06534bb8bdb49fd5e9e6632722c2989467c1bfc8e8d978dfb425d2685c2573",
         "6484aa716545ca2cf3a70c3fa8fe337e0a3d21162f0d6299a7bf8192bfd2a76f", "0000000000000000000000000000000000000000000000000000000000000001"}};
    fe v[4];
    for (int k = 0; k < 4; k++) {
        fe acc, c;
        fe_from_hex(&acc, K[k][3]);
        for (int i = 2; i >= 0; i--) { fe_mul(&acc, &acc, &in->x); fe_from_hex(&c, K[k][i]); fe_add(&acc, &acc, &c); }
        v[k] = acc;
    }
    if (fe_is_zero(&v[1]) || fe_is_zero(&v[3])) { memset(out, 0, sizeof(*out)); out->inf = 1; return; }
    fe i1, i3;
    fe_inv(&i1, &v[1]); fe_inv(&i3, &v[3]);
    fe_mul(&out->x, &v[0], &i1);
    fe_mul(&out->y, &v[2], &i3); fe_mul(&out->y, &out->y, &in->y);
    out->inf = 0;
}
/* `Secp256k1::hash_from_bytes::<ExpandMsgXmd<Sha256>>(&[m1, m2], &[DST])`
 * (rust-k256/src/randomizedsigner.rs:57-61, rust-k256/src/utils.rs:11-20) */
static void hash_to_curve2(aff* out, const uint8_t* m1, size_t l1, const uint8_t* m2, size_t l2) {
    uint8_t ub[96];
    expand_message_xmd(ub, 96, m1, l1, m2, l2);
    fe u0, u1;
    fe_from_be48(&u0, ub);
    fe_from_be48(&u1, ub + 48);
    aff q0, q1, t;
    map_to_curve_sswu(&t, &u0); iso_map(&q0, &t);
    map_to_curve_sswu(&t, &u1); iso_map(&q1, &t);
    jac j0, j1, s;
    jac_from_aff(&j0, &q0); jac_from_aff(&j1, &q1);
    jac_add(&s, &j0, &j1);
    jac_to_aff(out, &s);   /* cofactor 1 */
}

/* c = SHA-256(concat(enc(p_i)))   (rust-k256/src/lib.rs:159-168, randomizedsigner.rs:73-89) */
static void challenge(uint8_t* c, int version, const aff* pk, const aff* h, const aff* nul, const aff* rp, const aff* z) {
    sha256_t s;
    uint8_t e[33];
    sha256_init(&s);
    if (version == 1) {
        aff g;
        generator(&g);
        sha256_update(&s, e, encode_pt(e, &g));
        sha256_update(&s, e, encode_pt(e, pk));
        sha256_update(&s, e, encode_pt(e, h));
    }
    sha256_update(&s, e, encode_pt(e, nul));
    sha256_update(&s, e, encode_pt(e, rp));
    sha256_update(&s, e, encode_pt(e, z));
    sha256_final(&s, c);
}

/* ================================================================ public entry points */

/* rust-k256/src/randomizedsigner.rs:43-112, r supplied instead of drawn from the rng (:49).
 * Returns the status code of include/plume_b200.h; on non-zero status all outputs are zero. */
int plume_oracle_sign(int version, const uint8_t* msg, size_t len, const uint8_t* sk32, const uint8_t* r32, uint8_t* pk64,
                      uint8_t* nul64, uint8_t* c32, uint8_t* s32, uint8_t* rpoint64, uint8_t* hr64) {
    int st = 0;
    sc sk, r;
    from_be32(sk.v, sk32); from_be32(r.v, r32);
    aff g, R, K, h, z, nul;
    jac t;
    uint8_t cb[32];
    sc c, s;
    memset(&R, 0, sizeof(R)); memset(&K, 0, sizeof(K)); memset(&z, 0, sizeof(z)); memset(&nul, 0, sizeof(nul));
    if (!sc_valid_nonzero(r.v)) st = 1;
    else if (!sc_valid_nonzero(sk.v)) st = 2;
    if (!st) {
        generator(&g);
        jac_mul(&t, &g, &r); jac_to_aff(&R, &t);                 /* :51 */
        jac_mul(&t, &g, &sk); jac_to_aff(&K, &t);                /* :53 */
        uint8_t pkb[33];
        size_t npk = encode_pt(pkb, &K);                         /* :54 */
        hash_to_curve2(&h, msg, len, pkb, npk);                  /* :57-61 */
        if (h.inf) st = 5;
    }
    if (!st) {
        jac_mul(&t, &h, &r); jac_to_aff(&z, &t);                 /* :67 */
        jac_mul(&t, &h, &sk); jac_to_aff(&nul, &t);              /* :70 */
        challenge(cb, version, &K, &h, &nul, &R, &z);            /* :73-89 */
        from_be32(c.v, cb);
        if (!sc_valid_nonzero(c.v)) st = 3;                      /* :90-91 */
    }
    if (!st) {
        sc_mul(&s, &c, &sk); sc_add(&s, &s, &r);                 /* :94 */
        if (is_zero4(s.v)) st = 4;                               /* :95 */
    }
    if (st) {
        memset(pk64, 0, 64); memset(nul64, 0, 64); memset(c32, 0, 32); memset(s32, 0, 32);
        if (rpoint64) memset(rpoint64, 0, 64);
        if (hr64) memset(hr64, 0, 64);
        return st;
    }
    point_to_wire(pk64, &K); point_to_wire(nul64, &nul);
    memcpy(c32, cb, 32); to_be32(s32, s.v);
    if (rpoint64) point_to_wire(rpoint64, &R);
    if (hr64) point_to_wire(hr64, &z);
    return 0;
}

/* rust-k256/src/lib.rs:93-145.  Returns 1 iff verify() would return true. */
int plume_oracle_verify(int version, const uint8_t* msg, size_t len, const uint8_t* pk64, const uint8_t* nul64,
                        const uint8_t* c32, const uint8_t* s32, const uint8_t* rpoint64, const uint8_t* hr64) {
    aff pk, nul, rs, zs, g, h, A, B;
    sc c, s, mc;
    memset(&rs, 0, sizeof(rs)); memset(&zs, 0, sizeof(zs));
    if (!point_from_wire(&pk, pk64) || !point_from_wire(&nul, nul64)) return 0;  /* not representable as AffinePoint */
    from_be32(c.v, c32); from_be32(s.v, s32);
    if (!sc_valid_nonzero(c.v) || !sc_valid_nonzero(s.v)) return 0;             /* not NonZeroScalar */
    if (version == 1 && (!point_from_wire(&rs, rpoint64) || !point_from_wire(&zs, hr64))) return 0;
    generator(&g);
    sc_neg(&mc, &c);
    jac t1, t2, t;
    jac_mul(&t1, &g, &s); jac_mul(&t2, &pk, &mc); jac_add(&t, &t1, &t2); jac_to_aff(&A, &t);      /* :101 */
    uint8_t pkb[33];
    size_t npk = encode_pt(pkb, &pk);
    hash_to_curve2(&h, msg, len, pkb, npk);                                                        /* :103 */
    jac_mul(&t1, &h, &s); jac_mul(&t2, &nul, &mc); jac_add(&t, &t1, &t2); jac_to_aff(&B, &t);     /* :109 */
    if (version == 1) {
        if (A.inf != rs.inf || (!A.inf && (!fe_eq(&A.x, &rs.x) || !fe_eq(&A.y, &rs.y)))) return 0;  /* :117 */
        if (B.inf != zs.inf || (!B.inf && (!fe_eq(&B.x, &zs.x) || !fe_eq(&B.y, &zs.y)))) return 0;  /* :122 */
    }
    uint8_t d[32];
    challenge(d, version, &pk, &h, &nul, &A, &B);                                                   /* :127-143 */
    uint64_t dv[8] = {0};
    from_be32(dv, d);
    sc dr;
    sc_reduce8(&dr, dv);                                                                           /* Scalar::reduce */
    return memcmp(dr.v, c.v, 32) == 0;
}

/* out64 = hash_to_curve(msg) with the PLUME DST (rust-k256/tests/verification.rs `hash_to_secp`) */
void plume_oracle_hash_to_curve(const uint8_t* msg, size_t len, uint8_t* out64) {
    aff h;
    hash_to_curve2(&h, msg, len, msg, 0);
    point_to_wire(out64, &h);
}
/* k * G -> wire point, and SEC1 encodings (for the k*G table of rust-arkworks/src/tests/test_vectors.rs) */
void plume_oracle_mul_g(const uint8_t* k32, uint8_t* out64) {
    aff g, a;
    sc k;
    jac t;
    from_be32(k.v, k32);
    generator(&g);
    jac_mul(&t, &g, &k); jac_to_aff(&a, &t);
    point_to_wire(out64, &a);
}
void plume_oracle_mul(const uint8_t* p64, const uint8_t* k32, uint8_t* out64) {
    aff p, a;
    sc k;
    jac t;
    point_from_wire(&p, p64);
    from_be32(k.v, k32);
    jac_mul(&t, &p, &k); jac_to_aff(&a, &t);
    point_to_wire(out64, &a);
}
size_t plume_oracle_encode_pt(const uint8_t* p64, uint8_t* out33) {
    aff p;
    point_from_wire(&p, p64);
    return encode_pt(out33, &p);
}
/* SEC1-compressed 33-byte slot (identity: 00 + 32 zero bytes) <-> 64-byte wire point; decoding follows k256's
 * AffinePoint::from_encoded_point: prefix 02/03, x < p, x^3 + 7 must be a square.  Returns 1 when accepted. */
void plume_oracle_compress33(const uint8_t* p64, uint8_t* out33) {
    aff p;
    memset(out33, 0, 33);
    static const uint8_t zero[64] = {0};
    if (memcmp(p64, zero, 64) == 0) return;
    from_be32(p.x.v, p64); from_be32(p.y.v, p64 + 32); p.inf = 0;
    encode_pt(out33, &p);
}
int plume_oracle_decompress33(const uint8_t* in33, uint8_t* out64) {
    memset(out64, 0, 64);
    if (in33[0] == 0) {
        for (int i = 1; i < 33; i++) if (in33[i]) return 0;
        return 1;
    }
    if (in33[0] != 2 && in33[0] != 3) return 0;
    uint64_t x[4];
    from_be32(x, in33 + 1);
    if (ge4(x, P)) return 0;
    fe fx, rhs, y, seven;
    memcpy(fx.v, x, 32);
    fe_sqr(&rhs, &fx); fe_mul(&rhs, &rhs, &fx); fe_set_u64(&seven, 7); fe_add(&rhs, &rhs, &seven);
    if (!fe_sqrt(&y, &rhs)) return 0;
    if (fe_is_odd(&y) != (in33[0] & 1)) fe_neg(&y, &y);
    fe_to_be(out64, &fx); fe_to_be(out64 + 32, &y);
    return 1;
}
void plume_oracle_expand_message_xmd(const uint8_t* msg, size_t len, size_t n, uint8_t* out) {
    expand_message_xmd(out, n, msg, len, msg, 0);
}
void plume_oracle_sha256(const uint8_t* msg, size_t len, uint8_t* out32) {
    sha256_t s;
    sha256_init(&s); sha256_update(&s, msg, len); sha256_final(&s, out32);
}

/* ---------------------------------------------------------------- threaded batch drivers */
typedef struct {
    int kind, version;
    size_t i0, i1;
    const uint8_t* msgs; const uint64_t* offs; size_t msg_len;
    const uint8_t *a0, *a1, *a2, *a3, *a4, *a5;   /* inputs */
    uint8_t *o0, *o1, *o2, *o3, *o4, *o5, *o6;    /* outputs */
} job_t;

static const uint8_t* msg_at(const job_t* j, size_t i, size_t* len) {
    if (j->offs) { *len = (size_t)(j->offs[i + 1] - j->offs[i]); return j->msgs + j->offs[i]; }
    *len = j->msg_len;
    return j->msgs + i * j->msg_len;
}
static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (size_t i = j->i0; i < j->i1; i++) {
        size_t len;
        const uint8_t* m = msg_at(j, i, &len);
        if (j->kind == 0) {
            j->o6[i] = (uint8_t)plume_oracle_sign(j->version, m, len, j->a0 + 32 * i, j->a1 + 32 * i, j->o0 + 64 * i, j->o1 + 64 * i,
                                                  j->o2 + 32 * i, j->o3 + 32 * i, j->o4 ? j->o4 + 64 * i : NULL, j->o5 ? j->o5 + 64 * i : NULL);
        } else if (j->kind == 1) {
            j->o0[i] = (uint8_t)plume_oracle_verify(j->version, m, len, j->a0 + 64 * i, j->a1 + 64 * i, j->a2 + 32 * i, j->a3 + 32 * i,
                                                    j->a4 ? j->a4 + 64 * i : NULL, j->a5 ? j->a5 + 64 * i : NULL);
        } else {
            plume_oracle_hash_to_curve(m, len, j->o0 + 64 * i);
        }
    }
    return NULL;
}
static int run_jobs(job_t* proto, size_t n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *proto;
        jobs[t].i0 = n * (size_t)t / (size_t)nthreads;
        jobs[t].i1 = n * (size_t)(t + 1) / (size_t)nthreads;
        if (pthread_create(&th[t], NULL, worker, &jobs[t]) != 0) { worker(&jobs[t]); th[t] = 0; }
    }
    for (int t = 0; t < nthreads; t++) if (th[t]) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
int plume_oracle_sign_batch(int version, size_t n, const uint8_t* msgs, const uint64_t* offs, size_t msg_len, const uint8_t* sk,
                            const uint8_t* r, uint8_t* pk, uint8_t* nul, uint8_t* c, uint8_t* s, uint8_t* rpoint, uint8_t* hr,
                            uint8_t* status, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 0; j.version = version; j.msgs = msgs; j.offs = offs; j.msg_len = msg_len;
    j.a0 = sk; j.a1 = r; j.o0 = pk; j.o1 = nul; j.o2 = c; j.o3 = s; j.o4 = rpoint; j.o5 = hr; j.o6 = status;
    return run_jobs(&j, n, nthreads);
}
int plume_oracle_verify_batch(int version, size_t n, const uint8_t* msgs, const uint64_t* offs, size_t msg_len, const uint8_t* pk,
                              const uint8_t* nul, const uint8_t* c, const uint8_t* s, const uint8_t* rpoint, const uint8_t* hr,
                              uint8_t* ok, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 1; j.version = version; j.msgs = msgs; j.offs = offs; j.msg_len = msg_len;
    j.a0 = pk; j.a1 = nul; j.a2 = c; j.a3 = s; j.a4 = rpoint; j.a5 = hr; j.o0 = ok;
    return run_jobs(&j, n, nthreads);
}
int plume_oracle_h2c_batch(size_t n, const uint8_t* msgs, const uint64_t* offs, size_t msg_len, uint8_t* out, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 2; j.msgs = msgs; j.offs = offs; j.msg_len = msg_len; j.o0 = out;
    return run_jobs(&j, n, nthreads);
}

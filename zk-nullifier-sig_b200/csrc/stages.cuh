// stages.cuh -- per-item bodies of the pipeline stages.  Each kernel in k_sign.cu / k_verify.cu / k_misc.cu is
// `i = global thread id; if (i < n) stage_body(i, args)`; the host-sim test build loops the same
// bodies on the CPU.  Between stages the per-item state lives in HBM as structure-of-arrays
// "slots" of 32-byte field elements (8 little-endian limbs), so every access is a pair of
// coalesced 128-bit loads/stores, and the only field inversions are done by the batched
// inversion stage over a contiguous Z array.
//
// Sign   (rust-k256/src/randomizedsigner.rs:43-112):
//   S1 fixed-base  g^r, g^sk            -> Jacobian                     [sign_stage_fixed]
//   BI batched inversion of the 2 Z's
//   S2 affine R, pk; pk33; h = H2C(m || pk33) -> Jacobian               [sign_stage_h2c]
//   BI batched inversion of Z_h
//   S3 affine h; signed-comb table of h; h^r, h^sk -> Jacobian           [sign_stage_varbase_comb]
//      (windowed ladder per scalar: sign_stage_varbase, -DPLUME_SIGN_WINDOWED builds)
//   BI batched inversion of the 2 Z's
//   S4 affine z, nul; c = SHA-256(...); s = r + c*sk; outputs + status   [sign_stage_final]
// Verify (rust-k256/src/lib.rs:93-145):
//   V1 input checks; h = H2C(m || enc(pk)) -> Jacobian                  [verify_stage_h2c]
//   BI
//   V2 window tables of h, nul [verify_stage_mul_b1]; B = s*h - c*nul [verify_stage_mul_b2]; A = s*G - c*pk
//      [verify_stage_mul_a] -> Jacobian   (three kernels, each with its own register budget)
//   BI
//   V3 affine A, B; (V1: compare with r_point, hashed_to_curve_r); c == SHA-256(...) mod n  [verify_stage_final]
#pragma once
#include "h2c.cuh"
#include "inv.cuh"
#include "mul.cuh"

// ---- status codes (include/plume_b200.h) ---------------------------------------------------------
#define PLUME_ST_OK 0
#define PLUME_ST_BAD_R 1
#define PLUME_ST_BAD_SK 2
#define PLUME_ST_BAD_C 3
#define PLUME_ST_ZERO_S 4
#define PLUME_ST_H_INF 5
#define PLUME_ST_BAD_PK 6

// which reference crate's semantics a batch follows
#define PLUME_FLAVOUR_K256 0      // rust-k256/: pk derived from sk, c and s are NonZeroScalar (sign rejects c = 0 / c >= n / s = 0)
#define PLUME_FLAVOUR_ARKWORKS 1  // rust-arkworks/: pk is an input, every scalar is an Fr (zero allowed), c is reduced mod n

// ---- 32-byte element I/O ---------------------------------------------------------------------------
PLUME_DEV fe ld_fe(const uint32_t* p) {
    fe r;
#ifdef PLUME_HOSTSIM
    for (int i = 0; i < 8; i++) r.v[i] = p[i];
#else
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#endif
    return r;
}
PLUME_DEV void st_fe(uint32_t* p, const fe& a) {
#ifdef PLUME_HOSTSIM
    for (int i = 0; i < 8; i++) p[i] = a.v[i];
#else
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
#endif
}
// 32 big-endian bytes (16-byte aligned) -> limbs, and back
PLUME_DEV fe ld_fe_be(const uint8_t* p) {
    uint32_t w[8];
#ifdef PLUME_HOSTSIM
    for (int i = 0; i < 8; i++) w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
#else
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
#endif
    return fe_from_be_words(w);  // bswap per word + word reversal
}
PLUME_DEV void st_fe_be(uint8_t* p, const fe& a) {
    uint32_t w[8];
    fe_to_be_words(w, a);
#ifdef PLUME_HOSTSIM
    for (int i = 0; i < 8; i++) { p[4 * i] = (uint8_t)w[i]; p[4 * i + 1] = (uint8_t)(w[i] >> 8); p[4 * i + 2] = (uint8_t)(w[i] >> 16); p[4 * i + 3] = (uint8_t)(w[i] >> 24); }
#else
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
#endif
}
PLUME_DEV sc ld_sc_be(const uint8_t* p) {
    fe t = ld_fe_be(p);
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t.v[i];
    return r;
}
PLUME_DEV void st_sc_be(uint8_t* p, const sc& a) {
    fe t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.v[i] = a.v[i];
    st_fe_be(p, t);
}
PLUME_DEV void st_zero32(uint8_t* p) { st_fe_be(p, fe_zero()); }

// a >= p ?  (wire coordinates must be canonical, as k256's AffinePoint decoding requires)
PLUME_DEV bool fe_ge_p(const fe& a) {
    fe n = fe_norm(a);
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= n.v[i] ^ a.v[i];
    return d != 0;
}

// wire point (x || y, 64 bytes; 64 zero bytes = identity) -> aff; returns false if malformed
PLUME_DEV bool ld_point_be(aff& p, const uint8_t* src) {
    p.x = ld_fe_be(src);
    p.y = ld_fe_be(src + 32);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= p.x.v[i] | p.y.v[i];
    p.inf = (o == 0);
    if (p.inf) return true;
    if (fe_ge_p(p.x) || fe_ge_p(p.y)) return false;
    return aff_on_curve(p.x, p.y);
}
PLUME_DEV void st_point_be(uint8_t* dst, const aff& p) {
    if (p.inf) { st_zero32(dst); st_zero32(dst + 32); return; }
    st_fe_be(dst, p.x);
    st_fe_be(dst + 32, p.y);
}

// ---- messages ----------------------------------------------------------------------------------------
struct msg_view {
    const uint8_t* base;
    const uint64_t* offs;  // n+1 offsets, or null for fixed-length records
    uint32_t fixed_len;
};
PLUME_DEV const uint8_t* msg_ptr(const msg_view& m, uint32_t i, uint32_t& len) {
    if (m.offs) {
        uint64_t a = m.offs[i], b = m.offs[i + 1];
        len = (uint32_t)(b - a);
        return m.base + a;
    }
    len = m.fixed_len;
    return m.base + (size_t)i * m.fixed_len;
}

// SEC1 compressed encoding into a byte buffer; returns the length (33, or 1 for the identity)
// (rust-k256/src/utils.rs:23-25)
PLUME_DEV uint32_t enc_point33(uint8_t* out, const aff& p) {
    if (p.inf) { out[0] = 0; return 1; }
    out[0] = (uint8_t)(2 + (p.y.v[0] & 1));
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t w = p.x.v[7 - i];
        out[1 + 4 * i] = (uint8_t)(w >> 24);
        out[2 + 4 * i] = (uint8_t)(w >> 16);
        out[3 + 4 * i] = (uint8_t)(w >> 8);
        out[4 + 4 * i] = (uint8_t)w;
    }
    return 33;
}
PLUME_DEV void sha_put_point(sha256_stream& s, const aff& p) {
    if (p.inf) { sha256_stream_byte(s, 0); return; }
    sha256_stream_byte(s, (uint8_t)(2 + (p.y.v[0] & 1)));
#pragma unroll 1
    for (int i = 7; i >= 0; i--) sha256_stream_word(s, p.x.v[i]);
}
PLUME_DEV aff aff_generator() { aff g; g.x = ec_gx(); g.y = ec_gy(); g.inf = 0; return g; }

// c = SHA-256(enc(G) || enc(pk) || enc(h) || enc(nul) || enc(R) || enc(z))  (V1)
//   = SHA-256(enc(nul) || enc(R) || enc(z))                                  (V2)
// rust-k256/src/lib.rs:159-168, rust-k256/src/randomizedsigner.rs:73-89
// Fixed-layout fast path: when every point is finite the preimage is NP x 33 bytes at known offsets, so the message
// words are assembled in registers with constant shifts (the byte-stream hasher below goes through a local-memory byte
// buffer: ~8 000 instructions for the 198 bytes of V1 against ~150 here).
// W: 16 * ceil((33 NP + 9) / 64) words, zero-initialised.  Point j occupies bytes [33 j, 33 j + 33).
template <int J>
PLUME_DEV void challenge_put_point(uint32_t* W, const aff& p) {
    // the 33-byte string as nine left-aligned big-endian words: P0 = prefix | x[0..2], ..., P8 = last byte of x
    uint32_t X[8];
#pragma unroll
    for (int i = 0; i < 8; i++) X[i] = p.x.v[7 - i];
    uint32_t P[9];
    P[0] = ((2u + (p.y.v[0] & 1u)) << 24) | (X[0] >> 8);
#pragma unroll
    for (int i = 1; i < 8; i++) P[i] = (X[i - 1] << 24) | (X[i] >> 8);
    P[8] = X[7] << 24;
    constexpr int off = 33 * J, w0 = off / 4, r = off % 4;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        if (r == 0) {
            W[w0 + k] |= P[k];
        } else {
            W[w0 + k] |= P[k] >> (8 * r);
            W[w0 + k + 1] |= P[k] << (32 - 8 * r);
        }
    }
}
template <int NP>
PLUME_DEV sc challenge_finish(uint32_t* W) {
    constexpr int len = 33 * NP, nblk = (len + 9 + 63) / 64;
    W[len / 4] |= 0x80u << (24 - 8 * (len % 4));
    W[16 * nblk - 1] = len * 8;
    uint32_t st[8];
    sha256_init(st);
#pragma unroll
    for (int b = 0; b < nblk; b++) sha256_compress(st, W + 16 * b);
    sc c;
#pragma unroll
    for (int i = 0; i < 8; i++) c.v[i] = st[7 - i];
    return c;
}

PLUME_DEV sc plume_challenge(int version, const aff& pk, const aff& h, const aff& nul, const aff& R, const aff& z) {
    const bool finite3 = !(nul.inf | R.inf | z.inf);
    if (version == 1 && finite3 && !(pk.inf | h.inf)) {
        uint32_t W[64];
#pragma unroll
        for (int i = 0; i < 64; i++) W[i] = 0;
        challenge_put_point<0>(W, aff_generator());
        challenge_put_point<1>(W, pk);
        challenge_put_point<2>(W, h);
        challenge_put_point<3>(W, nul);
        challenge_put_point<4>(W, R);
        challenge_put_point<5>(W, z);
        return challenge_finish<6>(W);
    }
    if (version != 1 && finite3) {
        uint32_t W[32];
#pragma unroll
        for (int i = 0; i < 32; i++) W[i] = 0;
        challenge_put_point<0>(W, nul);
        challenge_put_point<1>(W, R);
        challenge_put_point<2>(W, z);
        return challenge_finish<3>(W);
    }
    // an identity among the points (1-byte encoding): the general byte-stream path
    sha256_stream s;
    sha256_init(s.st);
    s.fill = 0;
    s.total = 0;
    if (version == 1) {
        sha_put_point(s, aff_generator());
        sha_put_point(s, pk);
        sha_put_point(s, h);
    }
    sha_put_point(s, nul);
    sha_put_point(s, R);
    sha_put_point(s, z);
    uint32_t d[8];
    sha256_stream_final(s, d);
    sc c;
#pragma unroll
    for (int i = 0; i < 8; i++) c.v[i] = d[7 - i];
    return c;
}

// ---- workspace ---------------------------------------------------------------------------------------
// slot s, item i  ->  ws + (s * n + i) * 8 words
PLUME_DEV uint32_t* ws_at(uint32_t* ws, uint32_t n, int slot, uint32_t i) { return ws + ((size_t)slot * n + i) * 8; }

// slots shared by sign / verify / h2c pipelines
enum {
    WS_Z0 = 0, WS_Z1 = 1,        // contiguous Z array fed to the batched inversion (2n elements)
    WS_P0 = 2, WS_P1 = 3,        // prefix-product scratch of the batched inversion
    WS_AX = 4, WS_AY = 5,        // Jacobian X, Y of point A (sign: R = g^r, later z = h^r; verify: A)
    WS_BX = 6, WS_BY = 7,        // Jacobian X, Y of point B (sign: pk, later nul; verify: B)
    WS_HX = 8, WS_HY = 9,        // h: Jacobian X, Y, then affine x, y
    WS_RX = 10, WS_RY = 11,      // sign: affine R
    WS_KX = 12, WS_KY = 13,      // sign: affine pk
    WS_SLOTS = 14
};

struct sign_args {
    int version;
    int flavour;          // PLUME_FLAVOUR_*
    uint32_t n;
    msg_view msgs;
    const uint8_t* sk;    // n x 32 BE
    const uint8_t* r;     // n x 32 BE
    uint8_t* pk;          // n x 64   out (k256 flavour; may be null in the arkworks flavour)
    const uint8_t* pk_in; // n x 64   in  (arkworks flavour: the keypair's public half, rust-arkworks/src/lib.rs:229-235)
    uint8_t* nullifier;   // n x 64
    uint8_t* c;           // n x 32
    uint8_t* s;           // n x 32
    uint8_t* r_point;     // n x 64 or null
    uint8_t* hashed_to_curve_r;  // n x 64 or null
    uint8_t* status;      // n
    uint32_t* ws;
    const uint32_t* gtab;
    int gw;
    uint32_t* vbtab;      // n x COMB_AREA_WORDS words of table scratch
};

PLUME_DEV sc sc_one() { sc r; for (int i = 0; i < 8; i++) r.v[i] = (i == 0); return r; }

PLUME_DEV aff ws_load_affine(uint32_t* ws, uint32_t n, int sx, int sy, int sz, uint32_t i) {
    jac p;
    p.x = ld_fe(ws_at(ws, n, sx, i));
    p.y = ld_fe(ws_at(ws, n, sy, i));
    fe zinv = ld_fe(ws_at(ws, n, sz, i));
    p.z = zinv;
    p.inf = fe_is_zero(zinv);  // the inversion stage maps Z = 0 (identity) to 0
    return aff_from_jac_zinv(p, zinv);
}
PLUME_DEV void ws_store_jac(uint32_t* ws, uint32_t n, int sx, int sy, int sz, uint32_t i, const jac& p) {
    st_fe(ws_at(ws, n, sx, i), p.x);
    st_fe(ws_at(ws, n, sy, i), p.y);
    st_fe(ws_at(ws, n, sz, i), p.inf ? fe_zero() : p.z);
}
PLUME_DEV void ws_store_aff(uint32_t* ws, uint32_t n, int sx, int sy, uint32_t i, const aff& p) {
    st_fe(ws_at(ws, n, sx, i), p.x);
    st_fe(ws_at(ws, n, sy, i), p.y);
}

PLUME_DEV void sign_stage_fixed(uint32_t i, const sign_args& a) {
    sc r = ld_sc_be(a.r + (size_t)i * 32);
    sc sk = ld_sc_be(a.sk + (size_t)i * 32);
    uint8_t st = PLUME_ST_OK;
    if (a.flavour == PLUME_FLAVOUR_ARKWORKS) {
        // sign_with_r (rust-arkworks/src/lib.rs:229-278): r and sk are any Fr (zero included), pk comes with the keypair;
        // hash_to_curve fails on the identity pk (:97-100)
        if (sc_ge_n(sk)) { st = PLUME_ST_BAD_SK; sk = sc_one(); }
        if (sc_ge_n(r)) { st = PLUME_ST_BAD_R; r = sc_one(); }
        aff P;
        if (!ld_point_be(P, a.pk_in + (size_t)i * 64) || P.inf) { st = PLUME_ST_BAD_PK; P = aff_generator(); }
        a.status[i] = st;
        jac R = fb_mul(r, a.gtab, a.gw);
        ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, R);
        ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, jac_from_aff(P));
        return;
    }
    if (!sc_is_valid_nonzero(sk)) { st = PLUME_ST_BAD_SK; sk = sc_one(); }
    if (!sc_is_valid_nonzero(r)) { st = PLUME_ST_BAD_R; r = sc_one(); }
    a.status[i] = st;
    jac R = fb_mul(r, a.gtab, a.gw);
    st_fe(ws_at(a.ws, a.n, WS_AX, i), R.x);
    st_fe(ws_at(a.ws, a.n, WS_AY, i), R.y);
    st_fe(ws_at(a.ws, a.n, WS_Z0, i), R.z);
    jac K = fb_mul(sk, a.gtab, a.gw);
    st_fe(ws_at(a.ws, a.n, WS_BX, i), K.x);
    st_fe(ws_at(a.ws, a.n, WS_BY, i), K.y);
    st_fe(ws_at(a.ws, a.n, WS_Z1, i), K.z);
}

PLUME_DEV void sign_stage_h2c(uint32_t i, const sign_args& a) {
    aff R = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i);
    aff K = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i);
    ws_store_aff(a.ws, a.n, WS_RX, WS_RY, i, R);
    ws_store_aff(a.ws, a.n, WS_KX, WS_KY, i, K);
    uint8_t pk33[33];
    uint32_t npk = enc_point33(pk33, K);
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    jac h = h2c_hash_to_curve(m, len, pk33, npk);
    ws_store_jac(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i, h);
}

// tab: VB_TAB_WORDS words of this thread's global scratch
PLUME_DEV void sign_stage_varbase(uint32_t i, const sign_args& a, uint32_t* tab) {
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i);
    ws_store_aff(a.ws, a.n, WS_HX, WS_HY, i, h);
    if (h.inf) {
        // the reference panics here (randomizedsigner.rs:61); flag it and leave identity results
        a.status[i] = PLUME_ST_H_INF;
        jac o = jac_infinity();
        ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, o);
        ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, o);
        return;
    }
    fe zg = vb_build_table(h.x, h.y, tab, true);
    sc r = ld_sc_be(a.r + (size_t)i * 32);
    sc sk = ld_sc_be(a.sk + (size_t)i * 32);
    const bool zero_ok = a.flavour == PLUME_FLAVOUR_ARKWORKS;   // an Fr may be zero: h^0 is the identity
    if (sc_ge_n(sk) || (!zero_ok && sc_is_zero(sk))) sk = sc_one();
    if (sc_ge_n(r) || (!zero_ok && sc_is_zero(r))) r = sc_one();
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        jac o = vb_mul_tab(which == 0 ? r : sk, tab, zg);
        if (which == 0) ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, o);
        else ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, o);
    }
}

// The same stage with the signed comb (mul.cuh): the 99 doublings that build the teeth are shared by h^r and h^sk.
// `area`: COMB_AREA_WORDS words of this thread's global scratch.
PLUME_DEV void sign_stage_varbase_comb(uint32_t i, const sign_args& a, uint32_t* area) {
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i);
    ws_store_aff(a.ws, a.n, WS_HX, WS_HY, i, h);
    if (h.inf) {
        a.status[i] = PLUME_ST_H_INF;   // the reference panics here (randomizedsigner.rs:61)
        jac o = jac_infinity();
        ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, o);
        ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, o);
        return;
    }
    fe zg = comb_build_table(h.x, h.y, area);
    fe zg2 = fe_sqr(zg);
    fe hxs = fe_mul(h.x, zg2), hys = fe_mul(h.y, fe_mul(zg2, zg));   // h on the table's isomorphic curve
    sc r = ld_sc_be(a.r + (size_t)i * 32);
    sc sk = ld_sc_be(a.sk + (size_t)i * 32);
    const bool zero_ok = a.flavour == PLUME_FLAVOUR_ARKWORKS;
    if (sc_ge_n(sk) || (!zero_ok && sc_is_zero(sk))) sk = sc_one();
    if (sc_ge_n(r) || (!zero_ok && sc_is_zero(r))) r = sc_one();
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        jac o = comb_mul_tab(which == 0 ? r : sk, area, zg, hxs, hys);
        if (which == 0) ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, o);
        else ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, o);
    }
}

// The same stage as two kernels: the table (one thread per item) and the ladders (one thread per item AND scalar: 2n
// threads, thread idx < n takes r, idx >= n takes sk of item idx - n).  Each kernel's code is one small loop nest, so the
// instruction cache is not shared between warps in the table builder and warps in a ladder, and each gets its own register
// budget.  Zg travels through WS_P0 (scratch of the batched inversion, idle here); Zg = 0 marks an item without ladders.
PLUME_DEV void sign_stage_varbase_tab(uint32_t i, const sign_args& a, uint32_t* area) {
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i);
    ws_store_aff(a.ws, a.n, WS_HX, WS_HY, i, h);
    if (h.inf) {
        a.status[i] = PLUME_ST_H_INF;   // the reference panics here (randomizedsigner.rs:61)
        jac o = jac_infinity();
        ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, o);
        ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, o);
        st_fe(ws_at(a.ws, a.n, WS_P0, i), fe_zero());
        return;
    }
    st_fe(ws_at(a.ws, a.n, WS_P0, i), comb_build_table(h.x, h.y, area));
}
PLUME_DEV void sign_stage_varbase_lad(uint32_t idx, const sign_args& a, const uint32_t* vbtab) {
    const uint32_t which = idx >= a.n ? 1u : 0u, i = idx - which * a.n;
    fe zg = ld_fe(ws_at(a.ws, a.n, WS_P0, i));
    if (fe_is_zero(zg)) return;   // h was the identity: the table stage stored the results
    fe hx = ld_fe(ws_at(a.ws, a.n, WS_HX, i)), hy = ld_fe(ws_at(a.ws, a.n, WS_HY, i));
    fe zg2 = fe_sqr(zg);
    fe hxs = fe_mul(hx, zg2), hys = fe_mul(hy, fe_mul(zg2, zg));   // h on the table's isomorphic curve
    sc k = ld_sc_be((which ? a.sk : a.r) + (size_t)i * 32);
    const bool zero_ok = a.flavour == PLUME_FLAVOUR_ARKWORKS;
    if (sc_ge_n(k) || (!zero_ok && sc_is_zero(k))) k = sc_one();
    jac o = comb_mul_tab(k, vbtab + (size_t)i * VB_ITEM_WORDS, zg, hxs, hys);
    if (which == 0) ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, o);
    else ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, o);
}

// affine point read back from two workspace slots: (0, 0) is how the identity was stored (no curve point has y = 0)
PLUME_DEV aff ws_load_aff_xy(uint32_t* ws, uint32_t n, int sx, int sy, uint32_t i) {
    aff p;
    p.x = ld_fe(ws_at(ws, n, sx, i));
    p.y = ld_fe(ws_at(ws, n, sy, i));
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) o |= p.x.v[k] | p.y.v[k];
    p.inf = (o == 0);
    return p;
}

// challenge, s = r + c sk, status and outputs from the affine points of one item (K = pk, z = h^r)
PLUME_DEV void sign_final_finish(uint32_t i, const sign_args& a, const aff& K, const aff& h, const aff& nul, const aff& R, const aff& z) {
    uint8_t st = a.status[i];
    sc c = plume_challenge(a.version, K, h, nul, R, z);
    sc s = sc_one();
    if (a.flavour == PLUME_FLAVOUR_ARKWORKS) {
        // c = Fr::from_be_bytes_mod_order(digest); s = r + sk * c, nothing is rejected (rust-arkworks/src/lib.rs:257-262)
        c = sc_reduce256(c);
        if (st == PLUME_ST_OK) {
            sc r = ld_sc_be(a.r + (size_t)i * 32);
            sc sk = ld_sc_be(a.sk + (size_t)i * 32);
            s = sc_add(r, sc_mul(c, sk));
        }
    } else if (st == PLUME_ST_OK) {
        if (!sc_is_valid_nonzero(c)) {
            st = PLUME_ST_BAD_C;  // NonZeroScalar::from_repr(c).expect (randomizedsigner.rs:90-91)
        } else {
            sc r = ld_sc_be(a.r + (size_t)i * 32);
            sc sk = ld_sc_be(a.sk + (size_t)i * 32);
            s = sc_add(r, sc_mul(c, sk));  // randomizedsigner.rs:94
            if (sc_is_zero(s)) st = PLUME_ST_ZERO_S;
        }
    }
    a.status[i] = st;
    const bool ok = (st == PLUME_ST_OK);
    aff none = aff_infinity();
    if (a.pk) st_point_be(a.pk + (size_t)i * 64, ok ? K : none);
    st_point_be(a.nullifier + (size_t)i * 64, ok ? nul : none);
    if (ok) { st_sc_be(a.c + (size_t)i * 32, c); st_sc_be(a.s + (size_t)i * 32, s); }
    else { st_zero32(a.c + (size_t)i * 32); st_zero32(a.s + (size_t)i * 32); }
    if (a.r_point) st_point_be(a.r_point + (size_t)i * 64, ok ? R : none);
    if (a.hashed_to_curve_r) st_point_be(a.hashed_to_curve_r + (size_t)i * 64, ok ? z : none);
}
PLUME_DEV void sign_stage_final(uint32_t i, const sign_args& a) {
    aff z = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i);
    aff nul = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i);
    aff h = ws_load_aff_xy(a.ws, a.n, WS_HX, WS_HY, i);
    aff R = ws_load_aff_xy(a.ws, a.n, WS_RX, WS_RY, i);   // the identity only in the arkworks flavour (r = 0)
    aff K = ws_load_aff_xy(a.ws, a.n, WS_KX, WS_KY, i);
    sign_final_finish(i, a, K, h, nul, R, z);
}

// ---- batched inversion ---------------------------------------------------------------------------------
// Z[0..m) in place -> 1/Z (0 stays 0).  Thread t of T owns elements t, t+T, t+2T, ... (coalesced),
// multiplies them into a running product, inverts once, and walks back (Montgomery's trick):
// 3 multiplications per element + one ~270-multiplication inversion per K elements.  VAR: the inversion by division steps
// (inv.cuh) instead of Fermat's -- a quarter of the latency, which is what a small batch waits for.
template <bool VAR = false>
PLUME_DEV void binv_body(uint32_t t, uint32_t T, uint32_t* Z, uint32_t* scratch, uint32_t m) {
    fe acc = fe_one();
    uint32_t cnt = 0;
#pragma unroll 1
    for (uint32_t idx = t; idx < m; idx += T, cnt++) {
        fe z = ld_fe(Z + (size_t)idx * 8);
        st_fe(scratch + (size_t)idx * 8, acc);
        if (!fe_is_zero(z)) acc = fe_mul(acc, z);
    }
    if (cnt == 0) return;
    fe inv = VAR ? fe_inv_var(acc) : fe_inv(acc);
#pragma unroll 1
    for (uint32_t j = cnt; j-- > 0;) {
        uint32_t idx = t + j * T;
        fe z = ld_fe(Z + (size_t)idx * 8);
        if (fe_is_zero(z)) { st_fe(Z + (size_t)idx * 8, fe_zero()); continue; }
        fe pre = ld_fe(scratch + (size_t)idx * 8);
        st_fe(Z + (size_t)idx * 8, fe_mul(inv, pre));
        inv = fe_mul(inv, z);
    }
}

// ---- verify -----------------------------------------------------------------------------------------------
struct verify_args {
    int version;
    int flavour;               // PLUME_FLAVOUR_*; arkworks = verify_non_zk (rust-arkworks/src/tests.rs:28-78)
    uint32_t n;
    msg_view msgs;
    const uint8_t* pk;         // n x 64
    const uint8_t* nullifier;  // n x 64
    const uint8_t* c;          // n x 32
    const uint8_t* s;          // n x 32
    const uint8_t* r_point;    // n x 64 (V1)
    const uint8_t* hashed_to_curve_r;  // n x 64 (V1)
    uint8_t* ok;               // n
    uint32_t* ws;
    const uint32_t* gtab;
    int gw;
    uint32_t* vbtab;
};

// ok[i] is used as scratch between stages: 1 = inputs well-formed so far, 0 = reject
// the input checks of verify_stage_h2c; pk33 / the return value: the SEC1 form of pk (of G for a rejected item)
PLUME_DEV uint32_t verify_h2c_check(uint32_t i, const verify_args& a, uint8_t* pk33, bool& good) {
    aff pk, nul;
    good = ld_point_be(pk, a.pk + (size_t)i * 64);
    good = ld_point_be(nul, a.nullifier + (size_t)i * 64) && good;
    sc c = ld_sc_be(a.c + (size_t)i * 32), s = ld_sc_be(a.s + (size_t)i * 32);
    const bool ark = a.flavour == PLUME_FLAVOUR_ARKWORKS;
    if (ark) good = good && !sc_ge_n(c) && !sc_ge_n(s) && !pk.inf;   // Fr fields (zero allowed); hash_to_curve fails on pk = identity
    else good = good && sc_is_valid_nonzero(c) && sc_is_valid_nonzero(s);  // NonZeroScalar fields
    if (a.version == 1 || ark) {   // the arkworks signature always carries r_point and hashed_to_curve_r
        aff t;
        good = ld_point_be(t, a.r_point + (size_t)i * 64) && good;
        good = ld_point_be(t, a.hashed_to_curve_r + (size_t)i * 64) && good;
    }
    if (!good) pk = aff_generator();  // keep the lane on the common path; result is discarded
    return enc_point33(pk33, pk);
}
PLUME_DEV void verify_stage_h2c(uint32_t i, const verify_args& a) {
    uint8_t pk33[33];
    bool good;
    uint32_t npk = verify_h2c_check(i, a, pk33, good);
    a.ok[i] = good ? 1 : 0;
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    jac h = h2c_hash_to_curve(m, len, pk33, npk);  // lib.rs:103
    // Z of h goes to WS_Z1 (inverted there, consumed by the table stage, later overwritten by B's Z): WS_Z0 belongs to
    // A = G*s - pk*c alone, so that stage may run concurrently with the two stages of B (small batches: two streams)
    ws_store_jac(a.ws, a.n, WS_HX, WS_HY, WS_Z1, i, h);
}

// k * P for an affine P that may be the identity
PLUME_DEV jac vb_mul_point(const aff& p, const sc& k, uint32_t* tab) {
    if (p.inf) return jac_infinity();
    fe zg = vb_build_table(p.x, p.y, tab, true);
    return vb_mul_tab(k, tab, zg);
}

// h*s - nul*c as two kernels: the table pair (b1) and the double-base ladder (b2), so that the ladder, which is
// where the time goes, is compiled for 4 blocks per SM (b1 consumes the inverted Z of h in WS_Z1, which b2 then
// overwrites with the Z of the result).  tab1, tab2: two table areas of this thread (global scratch).
// zg travels through WS_KX, the "ladder to do" flag through WS_RY;
// the rare case of an identity among h, nul (adversarial inputs only) is finished inside b1.
PLUME_DEV void verify_stage_mul_b1(uint32_t i, const verify_args& a, uint32_t* tab1, uint32_t* tab2) {
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z1, i);
    ws_store_aff(a.ws, a.n, WS_HX, WS_HY, i, h);
    st_fe(ws_at(a.ws, a.n, WS_RX, i), fe_set_u32(h.inf));
    const bool good = a.ok[i] != 0;
    aff nul;
    if (good) ld_point_be(nul, a.nullifier + (size_t)i * 64);
    else nul = aff_generator();
    const bool both = !h.inf && !nul.inf;
    st_fe(ws_at(a.ws, a.n, WS_RY, i), fe_set_u32(both ? 1u : 0u));
    if (both) {
        st_fe(ws_at(a.ws, a.n, WS_KX, i), vb_build_table_pair(h.x, h.y, tab1, nul.x, nul.y, tab2));
        return;
    }
    sc c = sc_one(), s = sc_one();
    if (good) {
        c = ld_sc_be(a.c + (size_t)i * 32);
        s = ld_sc_be(a.s + (size_t)i * 32);
    }
    jac B = jac_add(vb_mul_point(h, s, tab1), vb_mul_point(nul, sc_neg(c), tab1));
    ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, B);
}
PLUME_DEV void verify_stage_mul_b2(uint32_t i, const verify_args& a, const uint32_t* tab1, const uint32_t* tab2) {
    if (ld_fe(ws_at(a.ws, a.n, WS_RY, i)).v[0] == 0) return;   // finished in b1
    sc c = sc_one(), s = sc_one();
    if (a.ok[i] != 0) {
        c = ld_sc_be(a.c + (size_t)i * 32);
        s = ld_sc_be(a.s + (size_t)i * 32);
    }
    jac B = vb_mul2_tab(s, tab1, sc_neg(c), tab2, ld_fe(ws_at(a.ws, a.n, WS_KX, i)));
    ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, B);
}
PLUME_DEV void verify_stage_mul_a(uint32_t i, const verify_args& a, uint32_t* tab1) {
    aff pk;
    bool good = a.ok[i] != 0;
    sc c = sc_one(), s = sc_one();
    if (good) {
        ld_point_be(pk, a.pk + (size_t)i * 64);
        c = ld_sc_be(a.c + (size_t)i * 32);
        s = ld_sc_be(a.s + (size_t)i * 32);
    } else {
        pk = aff_generator();
    }
    sc mc = sc_neg(c);
    jac A = fb_mul(s, a.gtab, a.gw);
    A = jac_add(A, vb_mul_point(pk, mc, tab1));
    ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, A);
}

// the checks of the last stage on the affine A = G*s - pk*c, B = h*s - nul*c and h
PLUME_DEV void verify_final_check(uint32_t i, const verify_args& a, const aff& A, const aff& B, const aff& h) {
    aff pk, nul;
    ld_point_be(pk, a.pk + (size_t)i * 64);
    ld_point_be(nul, a.nullifier + (size_t)i * 64);
    if (a.version == 1 || a.flavour == PLUME_FLAVOUR_ARKWORKS) {   // verify_non_zk compares both points in V2 as well
        aff rs, zs;
        ld_point_be(rs, a.r_point + (size_t)i * 64);
        ld_point_be(zs, a.hashed_to_curve_r + (size_t)i * 64);
        if (!aff_eq(A, rs)) { a.ok[i] = 0; return; }   // lib.rs:117
        if (!aff_eq(B, zs)) { a.ok[i] = 0; return; }   // lib.rs:122
    }
    sc d = sc_reduce256(plume_challenge(a.version, pk, h, nul, A, B));  // lib.rs:127-143
    sc c = ld_sc_be(a.c + (size_t)i * 32);
    a.ok[i] = sc_eq(c, d) ? 1 : 0;
}
PLUME_DEV void verify_stage_final(uint32_t i, const verify_args& a) {
    bool good = a.ok[i] != 0;
    if (!good) { a.ok[i] = 0; return; }
    aff A = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i);
    aff B = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i);
    aff h;
    h.x = ld_fe(ws_at(a.ws, a.n, WS_HX, i));
    h.y = ld_fe(ws_at(a.ws, a.n, WS_HY, i));
    h.inf = ld_fe(ws_at(a.ws, a.n, WS_RX, i)).v[0];
    verify_final_check(i, a, A, B, h);
}

// ---- hash_to_curve only (rust-k256/src/utils.rs:11-20 with the preimage supplied by the caller) ----------
struct h2c_args {
    uint32_t n;
    msg_view msgs;            // the full preimage (PLUME: m || enc33(pk)), or just m when pk33 is given
    const uint8_t* pk33;      // null, or n x 33 SEC1 slots appended to the messages (00 + zeros = the identity's one byte)
    uint8_t* out;             // n x 64 affine (zeros = identity)
    uint32_t* ws;
};
PLUME_DEV void h2c_stage_map(uint32_t i, const h2c_args& a) {
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    jac h;
    if (a.pk33) {   // utils::hash_to_curve(m, pk): the preimage is m || encode_pt(pk)  (rust-k256/src/utils.rs:11-20)
        uint8_t e[33];
#pragma unroll 1
        for (int k = 0; k < 33; k++) e[k] = a.pk33[(size_t)i * 33 + k];
        h = h2c_hash_to_curve(m, len, e, e[0] == 0 ? 1u : 33u);
    } else {
        h = h2c_hash_to_curve(m, len, m, 0);
    }
    ws_store_jac(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i, h);
}
PLUME_DEV void h2c_stage_out(uint32_t i, const h2c_args& a) {
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i);
    st_point_be(a.out + (size_t)i * 64, h);
}

// ---- k * G for a batch of scalars (public keys from secret keys; the public-key field of the JS wire form's SEC1-DER
// scalars, javascript/src/lib.rs:97-117) -------------------------------------------------------------------------------
struct fbmul_args {
    uint32_t n;
    const uint8_t* k;         // n x 32 big-endian, taken mod n
    uint8_t* out;             // n x 64 affine (zeros = identity)
    uint32_t* ws;
    const uint32_t* gtab;
    int gw;
};
PLUME_DEV void fbmul_stage_map(uint32_t i, const fbmul_args& a) {
    sc k = sc_reduce256(ld_sc_be(a.k + (size_t)i * 32));
    ws_store_jac(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i, fb_mul(k, a.gtab, a.gw));
}
PLUME_DEV void fbmul_stage_out(uint32_t i, const fbmul_args& a) {
    aff p = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i);
    st_point_be(a.out + (size_t)i * 64, p);
}

// ---- hash_to_curve with its intermediates (what a circuit-input generator starts from) ----------------
// The verify_nullifier circuit takes per-u witness hints (circuits/circom/verify_nullifier.circom:21-31) produced today by
// the external generate_inputs_from_array (circuits/circom/test/v1.test.ts:5,38-40).  The device computes the RFC 9380
// values those hints are derived from: u0, u1 = hash_to_field, which SSWU candidate was taken (g(x1) square or not),
// the mapped points Q0, Q1 = iso_map(map_to_curve(u_k)) ("x_mapped", "y_mapped") and h = Q0 + Q1.
struct h2cw_args {
    uint32_t n;
    msg_view msgs;
    uint8_t* u;          // n x 2 x 32  big-endian canonical u0, u1
    uint8_t* q;          // n x 2 x 64  Q0, Q1 affine
    uint8_t* gx1_square; // n x 2       1 when g(x1) is a square (x = x1), 0 when x = x2 = Z u^2 x1
    uint8_t* h;          // n x 64      Q0 + Q1
    uint8_t* hints;      // n x 2 x 3 x 32 or null: per u_k the circuit's gx1_sqrt, gx2_sqrt, y_pos (see h2cw_stage_map)
    uint32_t* ws;
};
PLUME_DEV void h2cw_stage_map(uint32_t i, const h2cw_args& a) {
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    fe u0, u1;
    h2c_hash_to_field(u0, u1, m, len, m, 0);
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        fe u = fe_norm(k == 0 ? u0 : u1);
        st_fe_be(a.u + ((size_t)i * 2 + k) * 32, u);
        fe xn, xd, y, root;
        bool sq = h2c_map_sswu(xn, xd, y, u, &root);
        a.gx1_square[(size_t)i * 2 + k] = sq ? 1 : 0;
        if (a.hints) {
            // Square-root hints of the circom hash_to_curve component (verify_nullifier.circom:21-31).  The generator that
            // defines them is not in the reference tree, so the convention is DECLARED here: exactly one of g(x1), g(x2)
            // is a square (Z is not); that one's hint is its EVEN square root (sgn0 = 0), the other hint is 0; y_pos is
            // the even square root of g(x) for the x that was taken -- the map's y is y_pos or p - y_pos by sgn0(u).
            root = fe_norm(root);
            if (root.v[0] & 1) root = fe_norm(fe_neg(root));
            uint8_t* hp = a.hints + ((size_t)i * 2 + k) * 96;
            st_fe_be(hp, sq ? root : fe_zero());
            st_fe_be(hp + 32, sq ? fe_zero() : root);
            st_fe_be(hp + 64, root);
        }
        jac qk = h2c_iso_map(xn, xd, y);
        if (k == 0) ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, qk);
        else ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, qk);
    }
}
PLUME_DEV void h2cw_stage_sum(uint32_t i, const h2cw_args& a) {
    aff q0 = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i);
    aff q1 = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i);
    st_point_be(a.q + (size_t)i * 128, q0);
    st_point_be(a.q + (size_t)i * 128 + 64, q1);
    jac h = jac_add_aff(jac_from_aff(q0), q1.x, q1.y, q1.inf);
    ws_store_jac(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i, h);
}
PLUME_DEV void h2cw_stage_out(uint32_t i, const h2cw_args& a) {
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i);
    st_point_be(a.h + (size_t)i * 64, h);
}
// 32 big-endian bytes -> the circuit's four 64-bit registers, least significant first (circuits/circom/utils.ts:11-17,32-51:
// bigIntToRegisters(value, 64, 4)); as bytes that is the reversal of the 32-byte string
PLUME_DEV void registers_body(uint32_t i, const uint8_t* in32, uint64_t* out4) {
    const uint8_t* s = in32 + (size_t)i * 32;
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        uint64_t v = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) v = (v << 8) | s[(3 - r) * 8 + b];
        out4[(size_t)i * 4 + r] = v;
    }
}

// ---- SEC1-compressed wire form (the form the JS binding exchanges, javascript/src/lib.rs:97-117) ----
// 33-byte slots: 02/03 || x for a finite point, 00 followed by 32 zero bytes for the identity.
PLUME_DEV void sec1_compress_body(uint32_t i, const uint8_t* in64, uint8_t* out33) {
    aff p;
    p.x = ld_fe_be(in64 + (size_t)i * 64);
    p.y = ld_fe_be(in64 + (size_t)i * 64 + 32);
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) o |= p.x.v[k] | p.y.v[k];
    p.inf = (o == 0);
    uint8_t e[33];
    uint32_t len = enc_point33(e, p);
    uint8_t* dst = out33 + (size_t)i * 33;
#pragma unroll 1
    for (uint32_t k = 0; k < 33; k++) dst[k] = k < len ? e[k] : 0;
}
// ok = 1 and out = x || y when the slot decodes the way k256's AffinePoint::from_encoded_point accepts it
// (prefix 02/03, x < p, x^3 + 7 a square); ok = 1 and out = 0 for the identity slot; ok = 0, out = 0 otherwise.
PLUME_DEV void sec1_decompress_body(uint32_t i, const uint8_t* in33, uint8_t* out64, uint8_t* ok) {
    const uint8_t* src = in33 + (size_t)i * 33;
    uint8_t prefix = src[0];
    uint32_t w[8];
    uint32_t any = 0;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        w[k] = ((uint32_t)src[1 + 4 * k] << 24) | ((uint32_t)src[2 + 4 * k] << 16) | ((uint32_t)src[3 + 4 * k] << 8) | src[4 + 4 * k];
        any |= w[k];
    }
    fe x;
#pragma unroll
    for (int k = 0; k < 8; k++) x.v[k] = w[7 - k];
    aff p = aff_infinity();
    bool good = false;
    if (prefix == 0) {
        good = (any == 0);
    } else if ((prefix == 2 || prefix == 3) && !fe_ge_p(x)) {
        fe rhs = fe_add(fe_mul(fe_sqr(x), x), fe_set_u32(7));
        fe y = fe_norm(fe_sqrt_cand(rhs));
        if (fe_eq(fe_sqr(y), rhs)) {
            if ((y.v[0] & 1) != (uint32_t)(prefix & 1)) y = fe_norm(fe_neg(y));
            p.x = x; p.y = y; p.inf = 0;
            good = true;
        }
    }
    st_point_be(out64 + (size_t)i * 64, p);
    ok[i] = good ? 1 : 0;
}
// ok[i] &= f0[i] & f1[i] & f2[i] & f3[i]   (null pointers are skipped)
PLUME_DEV void and_flags_body(uint32_t i, uint8_t* ok, const uint8_t* f0, const uint8_t* f1, const uint8_t* f2, const uint8_t* f3) {
    uint8_t v = ok[i];
    if (f0) v &= f0[i];
    if (f1) v &= f1[i];
    if (f2) v &= f2[i];
    if (f3) v &= f3[i];
    ok[i] = v;
}

// ---- generator table construction (once per context) -------------------------------------------------------
// bases[j] = 2^(w*j) * G, affine, 16 words each
PLUME_DEV void gtab_bases_body(uint32_t* bases, int w) {
    const int nwin = (256 + w - 1) / w;
    jac cur;
    cur.x = ec_gx(); cur.y = ec_gy(); cur.z = fe_one(); cur.inf = 0;
#pragma unroll 1
    for (int j = 0; j < nwin; j++) {
        fe zi = fe_inv(cur.z);
        aff p = aff_from_jac_zinv(cur, zi);
        st_fe(bases + j * 16, p.x);
        st_fe(bases + j * 16 + 8, p.y);
#pragma unroll 1
        for (int k = 0; k < w; k++) cur = jac_dbl(cur);
    }
}
// entry e = (j << w) + d: Jacobian d * bases[j] -> tab (X, Y) and Zs[e]
PLUME_DEV void gtab_entry_body(uint32_t e, uint32_t* tab, uint32_t* zs, const uint32_t* bases, int w) {
    uint32_t j = e >> w, d = e & ((1u << w) - 1);
    fe bx = ld_fe(bases + j * 16), by = ld_fe(bases + j * 16 + 8);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int b = w - 1; b >= 0; b--) {
        acc = jac_dbl(acc);
        if ((d >> b) & 1) acc = jac_add_aff(acc, bx, by, 0);
    }
    st_fe(tab + (size_t)e * 16, acc.x);
    st_fe(tab + (size_t)e * 16 + 8, acc.y);
    st_fe(zs + (size_t)e * 8, acc.inf ? fe_zero() : acc.z);
}
PLUME_DEV void gtab_norm_body(uint32_t e, uint32_t* tab, const uint32_t* zs) {
    jac p;
    p.x = ld_fe(tab + (size_t)e * 16);
    p.y = ld_fe(tab + (size_t)e * 16 + 8);
    fe zi = ld_fe(zs + (size_t)e * 8);
    p.z = zi;
    p.inf = fe_is_zero(zi);
    aff q = aff_from_jac_zinv(p, zi);
    st_fe(tab + (size_t)e * 16, q.x);
    st_fe(tab + (size_t)e * 16 + 8, q.y);
}

// sc.cuh -- scalars mod n (group order of secp256k1), GLV split and window recoding.
//
// Not throughput critical (a few hundred limb products per signature against ~3*10^5 in the
// curve arithmetic), so this is plain 32x32->64 C++ that compiles for the device and for the
// host-sim build alike.
//
// Replaces k256::Scalar / NonZeroScalar as used at rust-k256/src/randomizedsigner.rs:90-95
// (from_repr: reject c = 0 or c >= n; s = r + c*sk) and rust-k256/src/lib.rs:128,139
// (Scalar::reduce of the SHA-256 output).  n per rust-arkworks/src/secp256k1/fields/fr.rs:19.
#pragma once
#include "ptx.cuh"

struct sc { uint32_t v[8]; };

#define SC_N_LIMBS {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}

PLUME_DEV uint32_t sc_n_limb(int i) {
    switch (i) {
        case 0: return 0xD0364141u;
        case 1: return 0xBFD25E8Cu;
        case 2: return 0xAF48A03Bu;
        case 3: return 0xBAAEDCE6u;
        case 4: return 0xFFFFFFFEu;
        default: return 0xFFFFFFFFu;
    }
}
// 2^256 - n (129 bits)
PLUME_DEV uint32_t sc_nc_limb(int i) {
    switch (i) {
        case 0: return 0x2FC9BEBFu;
        case 1: return 0x402DA173u;
        case 2: return 0x50B75FC4u;
        case 3: return 0x45512319u;
        case 4: return 1u;
        default: return 0u;
    }
}

PLUME_DEV sc sc_from_be_words(const uint32_t* w) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = bswap32(w[7 - i]);
    return r;
}
PLUME_DEV void sc_to_be_words(uint32_t* w, const sc& a) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[7 - i] = bswap32(a.v[i]);
}
PLUME_DEV bool sc_is_zero(const sc& a) {
    return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}
// a >= n ?
PLUME_DEV bool sc_ge_n(const sc& a) {
    uint32_t bo = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a.v[i] - sc_n_limb(i) - bo;
        bo = (uint32_t)(t >> 63);
    }
    return bo == 0;
}
// 1 <= a < n  (k256 NonZeroScalar::from_repr accepts exactly these)
PLUME_DEV bool sc_is_valid_nonzero(const sc& a) { return !sc_is_zero(a) && !sc_ge_n(a); }
PLUME_DEV bool sc_eq(const sc& a, const sc& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
PLUME_DEV sc sc_sub_n_if_ge(const sc& a) {
    sc t;
    uint32_t bo = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - sc_n_limb(i) - bo;
        t.v[i] = (uint32_t)d;
        bo = (uint32_t)(d >> 63);
    }
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = bo ? a.v[i] : t.v[i];
    return r;
}
// n - a for a in [1, n); 0 -> 0
PLUME_DEV sc sc_neg(const sc& a) {
    sc r;
    uint32_t bo = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)sc_n_limb(i) - a.v[i] - bo;
        r.v[i] = (uint32_t)d;
        bo = (uint32_t)(d >> 63);
    }
    bool z = sc_is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = z ? 0 : r.v[i];
    return r;
}

// x (16 limbs, < 2^512) mod n, canonical
PLUME_DEV sc sc_reduce512(const uint32_t* x) {
    // fold 1: y (13 limbs) = x_lo + x_hi * NC
    uint32_t y[14];
#pragma unroll
    for (int i = 0; i < 14; i++) y[i] = (i < 8) ? x[i] : 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        uint64_t carry = 0;
        const uint32_t m = sc_nc_limb(j);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)x[8 + i] * m + y[i + j] + carry;
            y[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
#pragma unroll
        for (int i = 8 + j; i < 14; i++) {
            uint64_t t = (uint64_t)y[i] + carry;
            y[i] = (uint32_t)t;
            carry = t >> 32;
        }
    }
    // fold 2: z (10 limbs) = y_lo + y_hi(6 limbs, value < 2^130) * NC
    uint32_t z[12];
#pragma unroll
    for (int i = 0; i < 12; i++) z[i] = (i < 8) ? y[i] : 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        uint64_t carry = 0;
        const uint32_t m = sc_nc_limb(j);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            uint64_t t = (uint64_t)y[8 + i] * m + z[i + j] + carry;
            z[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
#pragma unroll
        for (int i = 6 + j; i < 12; i++) {
            uint64_t t = (uint64_t)z[i] + carry;
            z[i] = (uint32_t)t;
            carry = t >> 32;
        }
    }
    // fold 3: z_hi (limbs 8..11) is < 2^4; w = z_lo + z_hi * NC < 2^256 + 2^134
    uint32_t w[9];
    {
        uint64_t carry = 0;
        const uint32_t hi = z[8];  // z[9..11] are zero: total < 2^256 + 2^130 * 2^129
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)hi * sc_nc_limb(i) + z[i] + carry;
            w[i] = (uint32_t)t;
            carry = t >> 32;
        }
        w[8] = (uint32_t)carry;
    }
    // fold 4: w8 in {0,1}
    sc r;
    {
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)w[8] * sc_nc_limb(i) + w[i] + carry;
            r.v[i] = (uint32_t)t;
            carry = t >> 32;
        }
        // carry == 0 here: w8 = 1 implies w_lo < 2^134
    }
    return sc_sub_n_if_ge(r);
}

PLUME_DEV void sc_mul_wide(uint32_t* t, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t p = (uint64_t)a[i] * b[j] + t[i + j] + carry;
            t[i + j] = (uint32_t)p;
            carry = p >> 32;
        }
        t[j + 8] = (uint32_t)carry;
    }
}
PLUME_DEV sc sc_mul(const sc& a, const sc& b) {
    uint32_t t[16];
    sc_mul_wide(t, a.v, b.v);
    return sc_reduce512(t);
}
// (a + b) mod n for canonical a, b
PLUME_DEV sc sc_add(const sc& a, const sc& b) {
    uint32_t t[16];
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t s = (uint64_t)a.v[i] + b.v[i] + carry;
        t[i] = (uint32_t)s;
        carry = s >> 32;
    }
    t[8] = (uint32_t)carry;
#pragma unroll
    for (int i = 9; i < 16; i++) t[i] = 0;
    return sc_reduce512(t);
}
// reduce an arbitrary 256-bit value mod n (k256 `Scalar::reduce`, rust-k256/src/lib.rs:128)
PLUME_DEV sc sc_reduce256(const sc& a) { return sc_sub_n_if_ge(a); }

// ---- GLV: k = k1 + k2*lambda (mod n), |k1|, |k2| < 2^128 ---------------------------------------
// lambda^3 = 1 mod n, lambda*(x, y) = (beta*x, y).  Lattice constants are the well-known
// secp256k1 ones; tests/test_hostsim.py re-derives them (a1 - mb1*lambda = 0, a2 + b2*lambda = 0,
// g_i = round(2^384 * b_i / n)) and checks the 128-bit bound on random and extreme scalars.
struct glv_half { uint32_t mag[5]; uint32_t neg; };  // |k_i| (<= 129 bits) and its sign

PLUME_DEV uint32_t glv_const(int which, int i) {
    // which: 0 g1, 1 g2, 2 minus_b1 (128 bits), 3 minus_b2 = n - b2, 4 lambda
    const uint32_t G1[8] = {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
    const uint32_t G2[8] = {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
    const uint32_t MB1[8] = {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u, 0, 0, 0, 0};
    const uint32_t MB2[8] = {0x3DB1562Cu, 0xD765CDA8u, 0x0774346Du, 0x8A280AC5u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    const uint32_t LAM[8] = {0x1B23BD72u, 0xDF02967Cu, 0x20816678u, 0x122E22EAu, 0x8812645Au, 0xA5261C02u, 0xC05C30E0u, 0x5363AD4Cu};
    switch (which) {
        case 0: return G1[i];
        case 1: return G2[i];
        case 2: return MB1[i];
        case 3: return MB2[i];
        default: return LAM[i];
    }
}
PLUME_DEV sc glv_sc(int which) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = glv_const(which, i);
    return r;
}
// round(k * g / 2^384) as a scalar (fits 128 bits + 1)
PLUME_DEV sc glv_mul_shift384(const sc& k, int which) {
    sc g = glv_sc(which);
    uint32_t t[16];
    sc_mul_wide(t, k.v, g.v);
    sc r;
    uint64_t carry = (t[11] >> 31) & 1;  // rounding bit 383
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t s = (uint64_t)(i < 4 ? t[12 + i] : 0) + carry;
        r.v[i] = (uint32_t)s;
        carry = s >> 32;
    }
    return r;
}
PLUME_DEV glv_half glv_abs(const sc& k) {
    // k canonical; "high" when k > n/2 -> use n - k.  Any k whose top 96 bits are nonzero is high
    // because the halves are < 2^129.
    bool high = (k.v[7] | k.v[6] | k.v[5]) != 0 || (k.v[4] > 1);
    sc m = high ? sc_neg(k) : k;
    glv_half h;
#pragma unroll
    for (int i = 0; i < 5; i++) h.mag[i] = m.v[i];
    h.neg = high ? 1u : 0u;
    return h;
}
PLUME_DEV void glv_split(const sc& k, glv_half& h1, glv_half& h2) {
    sc c1 = glv_mul_shift384(k, 0);
    sc c2 = glv_mul_shift384(k, 1);
    sc k2 = sc_add(sc_mul(c1, glv_sc(2)), sc_mul(c2, glv_sc(3)));
    sc k1 = sc_add(k, sc_neg(sc_mul(k2, glv_sc(4))));
    h1 = glv_abs(k1);
    h2 = glv_abs(k2);
}

// ---- signed radix-16 (Booth) digits -------------------------------------------------------------
// m = sum_{i=0}^{32} d_i 16^i with d_i = w_i + b_{i-1} - 16 b_i in [-8, 8], where w_i is the i-th
// nibble and b_i its top bit: digit i only needs bits 4i-1 .. 4i+3, so no carry chain.
// The 33 digits are consumed most-significant first from a left-aligned 160-bit shift register:
// 5 words (w[4] = most significant), window = top 5 bits, shift left by 4 after each digit.
struct booth_reg { uint32_t w[5]; };

PLUME_DEV booth_reg booth_init(const glv_half& h) {
    // place m * 2 (one guard bit b_{-1} = 0 at the bottom) so that digit 32's window
    // (bits 127..131 of m) sits in the top 5 bits of 160: shift left by 160 - 132 - ... :
    // window for digit i covers bits [4i-1, 4i+3] of m = bits [4i, 4i+4] of 2m.  For i = 32 that is
    // bits [128, 132] of 2m -> want them at [155, 159]: shift 2m left by 27, i.e. m left by 28.
    booth_reg r;
    uint32_t m[6] = {h.mag[0], h.mag[1], h.mag[2], h.mag[3], h.mag[4], 0};
    // (m << 28) over 160 bits: word j = (m[j] << 28) | (m[j-1] >> 4); m < 2^129 so nothing is lost
    r.w[0] = m[0] << 28;
#pragma unroll
    for (int j = 1; j < 5; j++) r.w[j] = (m[j] << 28) | (m[j - 1] >> 4);
    return r;
}
// returns the next digit (most significant first) in [-8, 8] and advances
PLUME_DEV int booth_next(booth_reg& r) {
    uint32_t win = r.w[4] >> 27;  // 5 bits: [b_i w_i(3 bits below) ... b_{i-1}]
    int d = (int)((win >> 1) & 15) + (int)(win & 1) - (int)((win >> 4) << 4);
#pragma unroll
    for (int j = 4; j > 0; j--) r.w[j] = (r.w[j] << 4) | (r.w[j - 1] >> 28);
    r.w[0] <<= 4;
    return d;
}

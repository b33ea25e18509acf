// mul.cuh -- scalar multiplication: fixed-base (generator, precomputed table in HBM/L2) and
// variable-base (GLV + signed radix-16 windows over a per-thread co-Z table in shared memory).
//
// One thread owns one scalar multiplication; all threads of a warp add at the same loop step
// (fixed windows instead of NAF), so the warp never serialises on data-dependent add/skip
// decisions.  Replaces `ProjectivePoint * Scalar` / `NonIdentity * NonZeroScalar` of k256
// (rust-k256/src/randomizedsigner.rs:51,53,67,70; rust-k256/src/lib.rs:101,109).
#pragma once
#include "ec.cuh"
#include "sc.cuh"

// ---------------------------------------------------------------------------------------------
// Fixed base.  gtab holds, for window j (w bits each) and digit d in [1, 2^w), the affine point
// d * 2^(w*j) * G as 16 words (x limbs then y limbs) at index (j << w) + d.
// ---------------------------------------------------------------------------------------------
PLUME_DEV jac fb_mul(const sc& k, const uint32_t* gtab, int w) {
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = k.v[i];
    const int nwin = (256 + w - 1) / w;
    const uint32_t mask = (1u << w) - 1;
    jac acc = jac_infinity();
#pragma unroll 1
    for (int j = 0; j < nwin; j++) {
        uint32_t d = s[0] & mask;
        // s >>= w
#pragma unroll
        for (int i = 0; i < 7; i++) s[i] = (s[i] >> w) | (s[i + 1] << (32 - w));
        s[7] >>= w;
        if (d != 0) {
            const uint32_t* e = gtab + (((size_t)j << w) + d) * 16;
            fe qx, qy;
#ifdef PLUME_HOSTSIM
            for (int i = 0; i < 8; i++) { qx.v[i] = e[i]; qy.v[i] = e[8 + i]; }
#else
            const uint4* e4 = reinterpret_cast<const uint4*>(e);
            uint4 a = __ldg(e4), b = __ldg(e4 + 1), c = __ldg(e4 + 2), dd = __ldg(e4 + 3);
            qx.v[0] = a.x; qx.v[1] = a.y; qx.v[2] = a.z; qx.v[3] = a.w; qx.v[4] = b.x; qx.v[5] = b.y; qx.v[6] = b.z; qx.v[7] = b.w;
            qy.v[0] = c.x; qy.v[1] = c.y; qy.v[2] = c.z; qy.v[3] = c.w; qy.v[4] = dd.x; qy.v[5] = dd.y; qy.v[6] = dd.z; qy.v[7] = dd.w;
#endif
            acc = jac_add_aff(acc, qx, qy, 0);
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Variable base.  Per-thread table of 1P .. 8P, all brought to the common denominator Zg = Z(8P)
// ("co-Z": the entries are affine points of the isomorphic curve y^2 = x^3 + 7*Zg^6, on which the
// a = 0 doubling/addition formulas are unchanged), so the main loop uses the cheap mixed addition
// without any inversion; the result's Z is multiplied by Zg at the end.
// Table word (entry e in 0..7, word i in 0..15) lives at tab[(e*16 + i) * stride].
// ---------------------------------------------------------------------------------------------
#define VB_TAB_WORDS 128

// Where a thread's table lives.  vb_tab_strided: word (e, i) at p[(e*16 + i) * stride] -- shared memory,
// p already offset by the thread index and stride = threads per block, so every access is
// bank-conflict free whatever entries the lanes pick.  vb_tab_linear: 128 contiguous words per
// thread in global memory (L1/L2 resident scratch), entry = 64 contiguous bytes = four 128-bit loads.
struct vb_tab_strided {
    uint32_t* p;
    int stride;
    PLUME_DEV_MEMBER void store(int e, const fe& x, const fe& y) const {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            p[(e * 16 + i) * stride] = x.v[i];
            p[(e * 16 + 8 + i) * stride] = y.v[i];
        }
    }
    PLUME_DEV_MEMBER void load(int e, fe& x, fe& y) const {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x.v[i] = p[(e * 16 + i) * stride];
            y.v[i] = p[(e * 16 + 8 + i) * stride];
        }
    }
};
struct vb_tab_linear {
    uint32_t* p;
    PLUME_DEV_MEMBER void store(int e, const fe& x, const fe& y) const {
#ifdef PLUME_HOSTSIM
        for (int i = 0; i < 8; i++) { p[e * 16 + i] = x.v[i]; p[e * 16 + 8 + i] = y.v[i]; }
#else
        uint4* q = reinterpret_cast<uint4*>(p + e * 16);
        q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
        q[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
        q[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
#endif
    }
    PLUME_DEV_MEMBER void load(int e, fe& x, fe& y) const {
#ifdef PLUME_HOSTSIM
        for (int i = 0; i < 8; i++) { x.v[i] = p[e * 16 + i]; y.v[i] = p[e * 16 + 8 + i]; }
#else
        const uint4* q = reinterpret_cast<const uint4*>(p + e * 16);
        uint4 a = q[0], b = q[1], c = q[2], d = q[3];
        x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w; x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
        y.v[0] = c.x; y.v[1] = c.y; y.v[2] = c.z; y.v[3] = c.w; y.v[4] = d.x; y.v[5] = d.y; y.v[6] = d.z; y.v[7] = d.w;
#endif
    }
};

// builds the table for the affine, on-curve, non-identity point (px, py); returns Zg
template <class Tab>
PLUME_DEV fe vb_build_table(const fe& px, const fe& py, const Tab& tab) {
    fe hs[8];  // hs[k] = Z_{k+1} / Z_k, k = 1..7 (local memory; touched 14 times per table)
    jac cur;
    cur.x = px; cur.y = py; cur.z = fe_one(); cur.inf = 0;
    tab.store(0, px, py);
    cur = jac_dbl(cur);
    hs[1] = cur.z;
    tab.store(1, cur.x, cur.y);
#pragma unroll 1
    for (int k = 2; k < 8; k++) {
        // cur = k*P (Jacobian) ; next = cur + P ; Z_next = Z_cur * H with H = px*Z^2 - X
        fe z2 = fe_sqr(cur.z);
        fe H = fe_sub(fe_mul(px, z2), cur.x);
        cur = jac_add_aff(cur, px, py, 0);
        hs[k] = H;
        tab.store(k, cur.x, cur.y);
    }
    fe zg = cur.z;
    fe ratio = fe_one();
#pragma unroll 1
    for (int k = 7; k >= 1; k--) {
        ratio = fe_mul(ratio, hs[k]);  // Z_8 / Z_k
        fe x, y;
        tab.load(k - 1, x, y);
        fe r2 = fe_sqr(ratio);
        x = fe_mul(x, r2);
        y = fe_mul(y, fe_mul(r2, ratio));
        tab.store(k - 1, x, y);
    }
    return zg;
}

// acc += d * T  (d in [-8, 8], `flip` negates, `endo` applies (x, y) -> (beta*x, y))
template <class Tab>
PLUME_DEV jac vb_add_digit(const jac& acc, int d, uint32_t flip, bool endo, const Tab& tab) {
    if (d == 0) return acc;
    uint32_t neg = (d < 0 ? 1u : 0u) ^ flip;
    int e = (d < 0 ? -d : d) - 1;
    fe x, y;
    tab.load(e, x, y);
    if (endo) x = fe_mul(x, ec_beta());
    fe ny = fe_neg(y);
    y = fe_cmov(y, ny, neg != 0);
    return jac_add_aff(acc, x, y, 0);
}

// k * P from the prepared table; k canonical in [0, n)
template <class Tab>
PLUME_DEV jac vb_mul_tab(const sc& k, const Tab& tab, const fe& zg) {
    glv_half h1, h2;
    glv_split(k, h1, h2);
    booth_reg b1 = booth_init(h1), b2 = booth_init(h2);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int i = 32; i >= 0; i--) {
#pragma unroll 1
        for (int j = 0; j < 4; j++) acc = jac_dbl(acc);
        int d1 = booth_next(b1);
        int d2 = booth_next(b2);
        // one addition body for both halves (the loop is kept rolled on purpose: code size)
#pragma unroll 1
        for (int h = 0; h < 2; h++) acc = vb_add_digit(acc, h ? d2 : d1, h ? h2.neg : h1.neg, h != 0, tab);
    }
    if (!acc.inf) acc.z = fe_mul(acc.z, zg);
    return acc;
}

// k1 * P1 + k2 * P2 with shared doublings (Straus): both tables must be expressed on the same
// isomorphic curve (common denominator zg), see vb_build_table_pair.
template <class Tab>
PLUME_DEV jac vb_mul2_tab(const sc& k1, const Tab& tab1, const sc& k2, const Tab& tab2, const fe& zg) {
    glv_half h[4];
    glv_split(k1, h[0], h[1]);
    glv_split(k2, h[2], h[3]);
    booth_reg b0 = booth_init(h[0]), b1 = booth_init(h[1]), b2 = booth_init(h[2]), b3 = booth_init(h[3]);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int i = 32; i >= 0; i--) {
#pragma unroll 1
        for (int j = 0; j < 4; j++) acc = jac_dbl(acc);
        int d0 = booth_next(b0), d1 = booth_next(b1), d2 = booth_next(b2), d3 = booth_next(b3);
        // one addition body for the four half-scalars (rolled on purpose: code size)
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            int d = q == 0 ? d0 : q == 1 ? d1 : q == 2 ? d2 : d3;
            uint32_t flip = q == 0 ? h[0].neg : q == 1 ? h[1].neg : q == 2 ? h[2].neg : h[3].neg;
            if (q < 2) acc = vb_add_digit(acc, d, flip, (q & 1) != 0, tab1);
            else acc = vb_add_digit(acc, d, flip, (q & 1) != 0, tab2);
        }
    }
    if (!acc.inf) acc.z = fe_mul(acc.z, zg);
    return acc;
}

// Tables of two affine, on-curve, non-identity points over one common denominator.  P1's table is
// built first (denominator z1); P2 is moved onto P1's isomorphic curve (x*z1^2, y*z1^3) and its table
// built there (relative denominator z2); P1's entries are then rescaled by z2.  Returns z1*z2.
template <class Tab>
PLUME_DEV fe vb_build_table_pair(const fe& p1x, const fe& p1y, const Tab& tab1, const fe& p2x, const fe& p2y, const Tab& tab2) {
    fe z1 = vb_build_table(p1x, p1y, tab1);
    fe z1_2 = fe_sqr(z1);
    fe qx = fe_mul(p2x, z1_2);
    fe qy = fe_mul(p2y, fe_mul(z1_2, z1));
    fe z2 = vb_build_table(qx, qy, tab2);
    fe z2_2 = fe_sqr(z2), z2_3 = fe_mul(z2_2, z2);
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        fe x, y;
        tab1.load(e, x, y);
        tab1.store(e, fe_mul(x, z2_2), fe_mul(y, z2_3));
    }
    return fe_mul(z1, z2);
}

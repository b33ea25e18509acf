// mul.cuh -- scalar multiplication: fixed-base (generator, precomputed window table in HBM), variable-base (GLV +
// signed radix-16 windows over a per-thread co-Z table in global scratch or shared memory; Straus for two bases) and
// a signed comb for several scalars on one variable base (end of file).
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

// ---------------------------------------------------------------------------------------------
// Signed comb for SEVERAL scalars on ONE variable base (the signer's h^r and h^sk).
//
// With teeth at 2^(33 j), j = 0..3, a 132-bit odd magnitude m = sum_i b_i 2^i with every b_i in {-1, +1}
// (b_i = +1 iff bit i+1 of m is set, b_131 = +1) is  sum_{c=0..32} 2^c * (b_c T0 + b_{c+33} T1 + b_{c+66} T2 + b_{c+99} T3)
// with T_j = 2^(33 j) P: 33 doublings and 33 additions of a table entry  +-(T3 +- T2 +- T1 +- T0)  per half-scalar,
// the 99 doublings that produce T1..T3 being paid ONCE for all scalars on this base.  Two scalars with GLV halves:
// 99 + 2*33 doublings and 4*33 additions instead of 2*132 doublings and 4*33 additions of the windowed ladder.
// The eight entries (T3 positive) are brought to a common denominator Zg like the window table above, so the main
// loop uses mixed additions and the result's Z is multiplied by Zg at the end.  An even magnitude is made odd by
// adding one, and the base is subtracted again at the end.
//
// Storage (global scratch, 256 words per item): words 0..127 the table (entry e: x at 16e, y at 16e+8), words
// 128..191 beta * x of every entry (the endomorphism of the second GLV half, so no multiplication per addition),
// words 192..255 scratch for T1, T2 while the chain runs.
// ---------------------------------------------------------------------------------------------
#define COMB_TEETH_BITS 33
#define COMB_AREA_WORDS 256

PLUME_DEV void comb_st_fe(uint32_t* p, const fe& a) {
#ifdef PLUME_HOSTSIM
    for (int i = 0; i < 8; i++) p[i] = a.v[i];
#else
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
#endif
}
PLUME_DEV fe comb_ld_fe(const uint32_t* p) {
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
PLUME_DEV void comb_st_jac(uint32_t* p, const jac& a) { comb_st_fe(p, a.x); comb_st_fe(p + 8, a.y); comb_st_fe(p + 16, a.z); }
PLUME_DEV jac comb_ld_jac(const uint32_t* p) { jac r; r.x = comb_ld_fe(p); r.y = comb_ld_fe(p + 8); r.z = comb_ld_fe(p + 16); r.inf = 0; return r; }

// Builds the table for the affine, on-curve, non-identity point (px, py) of prime order; returns Zg.
PLUME_DEV fe comb_build_table(const fe& px, const fe& py, uint32_t* area) {
    uint32_t* tmp = area + 192;
    jac cur;
    cur.x = px; cur.y = py; cur.z = fe_one(); cur.inf = 0;
#pragma unroll 1
    for (int j = 1; j <= 3; j++) {
#pragma unroll 1
        for (int i = 0; i < COMB_TEETH_BITS; i++) cur = jac_dbl(cur);
        if (j < 3) comb_st_jac(tmp + (j - 1) * 24, cur);   // T1, T2
    }
    // cur = T3.  lo[q]: q = 0: -(T1 + T0), 1: -(T1 - T0) , 2: T1 - T0, 3: T1 + T0   (index = t1 t0, t = "same sign as T3")
    // entries: e = t2 t1 t0:  T3 + (t2 ? T2 : -T2) + lo[t1 t0]
    jac t2p = comb_ld_jac(tmp + 24);
    jac hi1 = jac_add(cur, t2p);            // T3 + T2
    jac hi0 = jac_add(cur, jac_neg(t2p));   // T3 - T2
    jac t1p = comb_ld_jac(tmp);
    jac s = jac_add_aff(t1p, px, py, 0);                    // T1 + T0
    jac d = jac_add_aff(t1p, px, fe_neg(py), 0);            // T1 - T0
    // the four low combinations live in the scratch area (T1, T2 are dead now): 4 x 24 words = words 192..255 + 32 words
    // of the beta area, which is written last
    uint32_t* lo = area + 160;
    comb_st_jac(lo + 0 * 24, jac_neg(s));
    comb_st_jac(lo + 1 * 24, jac_neg(d));
    comb_st_jac(lo + 2 * 24, d);
    comb_st_jac(lo + 3 * 24, s);
    // entries as Jacobian points: x, y into the table, z kept in local memory for the common-denominator pass
    fe zs[8];
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        jac l = comb_ld_jac(lo + (e & 3) * 24);
        jac r = jac_add((e & 4) ? hi1 : hi0, l);
        comb_st_fe(area + e * 16, r.x);
        comb_st_fe(area + e * 16 + 8, r.y);
        zs[e] = r.z;
    }
    // common denominator Zg = z0 z1 ... z7: entry e is scaled by f = Zg / z_e (x f^2, y f^3)
    fe pre[8];                              // pre[e] = z0 ... z_{e-1}
    fe acc = fe_one();
#pragma unroll 1
    for (int e = 0; e < 8; e++) { pre[e] = acc; acc = fe_mul(acc, zs[e]); }
    fe zg = acc;
    fe suf = fe_one();                      // z_{e+1} ... z7
    const fe beta = ec_beta();
#pragma unroll 1
    for (int e = 7; e >= 0; e--) {
        fe f = fe_mul(pre[e], suf);
        suf = fe_mul(suf, zs[e]);
        fe f2 = fe_sqr(f);
        fe x = fe_mul(comb_ld_fe(area + e * 16), f2);
        fe y = fe_mul(comb_ld_fe(area + e * 16 + 8), fe_mul(f2, f));
        comb_st_fe(area + e * 16, x);
        comb_st_fe(area + e * 16 + 8, y);
    }
#pragma unroll 1
    for (int e = 0; e < 8; e++) comb_st_fe(area + 128 + e * 8, fe_mul(comb_ld_fe(area + e * 16), beta));
    return zg;
}

// rows of the signed-digit matrix of one half-scalar: bit c of row[j] = 1 iff b_{c + 33 j} = +1
struct comb_rows { uint64_t row[4]; uint32_t neg; uint32_t even; };
PLUME_DEV comb_rows comb_recode(const glv_half& h) {
    // m' = m | 1 (m + 1 when m is even); digits from m'' = (m' >> 1) | 2^131
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = h.mag[i];
    comb_rows r;
    r.even = (w[0] & 1) ^ 1;
    r.neg = h.neg;
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
    w[4] = (w[4] >> 1) | (1u << 3);        // bit 131 = bit 3 of word 4
    // row j = bits [33 j, 33 j + 33)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int lo = COMB_TEETH_BITS * j, wi = lo >> 5, sh = lo & 31;
        uint64_t v = ((uint64_t)w[wi] >> sh);
        if (wi + 1 < 5) v |= (uint64_t)w[wi + 1] << (32 - sh);
        if (wi + 2 < 5 && sh != 0) v |= (uint64_t)w[wi + 2] << (64 - sh);
        r.row[j] = v & ((1ull << COMB_TEETH_BITS) - 1);
    }
    return r;
}

// acc += (digit column c of rows) * table, on the second GLV half with the beta-twisted x
PLUME_DEV jac comb_add_column(const jac& acc, const comb_rows& r, int c, bool endo, const uint32_t* area) {
    uint32_t s0 = (uint32_t)(r.row[0] >> c) & 1, s1 = (uint32_t)(r.row[1] >> c) & 1, s2 = (uint32_t)(r.row[2] >> c) & 1,
             s3 = (uint32_t)(r.row[3] >> c) & 1;
    uint32_t e = ((s2 == s3) << 2) | ((s1 == s3) << 1) | (s0 == s3);
    uint32_t neg = (s3 ^ 1) ^ r.neg;
    fe x = comb_ld_fe(endo ? area + 128 + e * 8 : area + e * 16);
    fe y = comb_ld_fe(area + e * 16 + 8);
    fe ny = fe_neg(y);
    y = fe_cmov(y, ny, neg != 0);
    return jac_add_aff(acc, x, y, 0);
}

// k * P from the prepared comb table; k canonical in [0, n); (px, py) = P on the isomorphic curve is entry-free: the
// parity correction needs P itself, expressed with the table's denominator: P' = (px zg^2, py zg^3)
PLUME_DEV jac comb_mul_tab(const sc& k, const uint32_t* area, const fe& zg, const fe& pxs, const fe& pys) {
    glv_half h1, h2;
    glv_split(k, h1, h2);
    comb_rows r1 = comb_recode(h1), r2 = comb_recode(h2);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int c = COMB_TEETH_BITS - 1; c >= 0; c--) {
        acc = jac_dbl(acc);
#pragma unroll 1
        for (int h = 0; h < 2; h++) acc = comb_add_column(acc, h ? r2 : r1, c, h != 0, area);
    }
    // undo the "+ 1" of the even magnitudes: subtract sign * P (first half) / sign * beta(P) (second half)
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        const comb_rows& r = h ? r2 : r1;
        if (r.even) {
            fe x = h ? fe_mul(pxs, ec_beta()) : pxs;
            fe y = r.neg ? pys : fe_neg(pys);     // subtracting (+P) when the half is positive
            acc = jac_add_aff(acc, x, y, 0);
        }
    }
    if (!acc.inf) acc.z = fe_mul(acc.z, zg);
    return acc;
}

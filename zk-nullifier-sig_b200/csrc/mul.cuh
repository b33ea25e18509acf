// mul.cuh -- scalar multiplication: fixed-base (generator, precomputed window table in HBM), variable-base (GLV +
// signed radix-16 windows over a per-thread co-Z table in global scratch; Straus for two bases) and a signed comb
// for several scalars on one variable base (end of file).
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
// Per-thread tables in global scratch (L1/L2 resident).  One entry is ONE 128-byte line:
//     words 0..7   x        words 8..15   y        words 16..23  beta * x        words 24..31  -y
// so that an addition of +-entry, with or without the endomorphism (x, y) -> (beta x, y) of the second GLV half, is
// four 128-bit loads and nothing else: no multiplication by beta and no negate-and-select per addition (round 1 paid
// one fe_mul on every second addition and a subtraction plus eight selects on every addition).  All entries of a table
// share one denominator Zg ("co-Z": they are affine points of the isomorphic curve y^2 = x^3 + 7 Zg^6, on which the a = 0
// formulas are unchanged), so the main loops use the mixed addition without any inversion and the result's Z is
// multiplied by Zg at the end.
// ---------------------------------------------------------------------------------------------
#define VB_ENT_WORDS 32
#define VB_TAB_WORDS (8 * VB_ENT_WORDS)   // window table: 1P .. 8P

PLUME_DEV void ent_st_fe(uint32_t* p, const fe& a) {
#ifdef PLUME_HOSTSIM
    for (int i = 0; i < 8; i++) p[i] = a.v[i];
#else
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
#endif
}
PLUME_DEV fe ent_ld_fe(const uint32_t* p) {
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
// raw coordinates of an entry (before the common-denominator pass) / the finished entry
PLUME_DEV void ent_store_xy(uint32_t* e, const fe& x, const fe& y) { ent_st_fe(e, x); ent_st_fe(e + 8, y); }
PLUME_DEV void ent_finish(uint32_t* e, const fe& x, const fe& y) {
    ent_st_fe(e, x);
    ent_st_fe(e + 8, y);
    ent_st_fe(e + 16, fe_mul(x, ec_beta()));
    ent_st_fe(e + 24, fe_neg(y));
}
// +-entry, optionally through the endomorphism
PLUME_DEV void ent_load(const uint32_t* e, bool endo, bool neg, fe& x, fe& y) {
    x = ent_ld_fe(e + (endo ? 16 : 0));
    y = ent_ld_fe(e + (neg ? 24 : 8));
}

// P + Q for the table builders: Q affine, P != +-Q and both finite by construction (small multiples of a point of
// prime order); also hands back H = Z3 / Z1, which the common-denominator pass needs.
PLUME_DEV jac jac_add_aff_h(const jac& p, const fe& qx, const fe& qy, fe& H) {
    fe z2 = fe_sqr(p.z);
    H = fe_sub(fe_mul(qx, z2), p.x);
    fe R = fe_sub(fe_mul(qy, fe_mul(p.z, z2)), p.y);
    jac r;
    r.z = fe_mul(p.z, H);
    fe H2 = fe_sqr(H);
    fe H3 = fe_mul(H, H2);
    fe V = fe_mul(p.x, H2);
    r.x = fe_sub(fe_sub(fe_sqr(R), H3), fe_dbl(V));
    r.y = fe_sub(fe_mul(R, fe_sub(V, r.x)), fe_mul(p.y, H3));
    r.inf = 0;
    return r;
}

// Window table 1P .. 8P of the affine, on-curve, non-identity point (px, py), entries brought to the denominator
// Zg = Z(8P), which is returned.  finish = false leaves the raw (x, y) only (the pair builder rescales them once more).
PLUME_DEV fe vb_build_table(const fe& px, const fe& py, uint32_t* tab, bool finish) {
    fe hs[8];  // hs[k] = Z_{k+1} / Z_k, k = 1..7 (local memory; touched 14 times per table)
    jac cur;
    cur.x = px; cur.y = py; cur.z = fe_one(); cur.inf = 0;
    ent_store_xy(tab, px, py);
    cur = jac_dbl_fast(cur);
    hs[1] = cur.z;
    ent_store_xy(tab + VB_ENT_WORDS, cur.x, cur.y);
#pragma unroll 1
    for (int k = 2; k < 8; k++) {
        fe H;
        cur = jac_add_aff_h(cur, px, py, H);   // (k+1) P, Z_next = Z_cur * H
        hs[k] = H;
        if (k < 7 || !finish) ent_store_xy(tab + k * VB_ENT_WORDS, cur.x, cur.y);
    }
    if (finish) ent_finish(tab + 7 * VB_ENT_WORDS, cur.x, cur.y);
    fe zg = cur.z;
    fe ratio = fe_one();
#pragma unroll 1
    for (int k = 7; k >= 1; k--) {
        uint32_t* e = tab + (k - 1) * VB_ENT_WORDS;
        // the raw entry was stored a while ago and comes back from the L2: ask for it before the three field operations
        // that do not need it (40 % of this loop's stall samples were waits on these loads when they sat next to their use)
        const fe x0 = ent_ld_fe(e), y0 = ent_ld_fe(e + 8);
        ratio = (k == 7) ? hs[7] : fe_mul(ratio, hs[k]);  // Z_8 / Z_k
        fe r2 = fe_sqr(ratio);
        fe r3 = fe_mul(r2, ratio);
        fe x = fe_mul(x0, r2);
        fe y = fe_mul(y0, r3);
        if (finish) ent_finish(e, x, y);
        else ent_store_xy(e, x, y);
    }
    return zg;
}

// acc += d * T  (d in [-8, 8], `flip` negates, `endo` applies (x, y) -> (beta*x, y))
PLUME_DEV jac vb_add_digit(const jac& acc, int d, uint32_t flip, bool endo, const uint32_t* tab) {
    if (d == 0) return acc;
    const bool neg = ((d < 0 ? 1u : 0u) ^ flip) != 0;
    const int e = (d < 0 ? -d : d) - 1;
    fe x, y;
    ent_load(tab + e * VB_ENT_WORDS, endo, neg, x, y);
    return jac_add_aff_fast(acc, x, y);
}

// k * P from the prepared table; k canonical in [0, n)
PLUME_DEV jac vb_mul_tab(const sc& k, const uint32_t* tab, const fe& zg) {
    glv_half h1, h2;
    glv_split(k, h1, h2);
    booth_reg b1 = booth_init(h1), b2 = booth_init(h2);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int i = 32; i >= 0; i--) {
        if (i < 32) {   // the accumulator is still the identity in the top window: nothing to double
#pragma unroll 1
            for (int j = 0; j < 4; j++) acc = jac_dbl_fast(acc);
        }
        int d1 = booth_next(b1);
        int d2 = booth_next(b2);
        // one addition body for both halves (the loop is kept rolled on purpose: code size)
#pragma unroll 1
        for (int h = 0; h < 2; h++) acc = vb_add_digit(acc, h ? d2 : d1, h ? h2.neg : h1.neg, h != 0, tab);
    }
    if (!acc.inf) acc.z = fe_mul(acc.z, zg);
    return acc;
}

// Windows [j0, j1) of fb_mul (the small-batch kernels split one scalar over two lanes); j0 = 0, j1 = fb_windows(w): all.
PLUME_DEV int fb_windows(int w) { return (256 + w - 1) / w; }
PLUME_DEV jac fb_mul_windows(const sc& k, const uint32_t* gtab, int w, int j0, int j1) {
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = k.v[i];
    const uint32_t mask = (1u << w) - 1;
    jac acc = jac_infinity();
#pragma unroll 1
    for (int j = 0; j < j1; j++) {
        uint32_t d = s[0] & mask;
#pragma unroll
        for (int i = 0; i < 7; i++) s[i] = (s[i] >> w) | (s[i + 1] << (32 - w));
        s[7] >>= w;
        if (d != 0 && j >= j0) {
            const uint32_t* e = gtab + (((size_t)j << w) + d) * 16;
            acc = jac_add_aff(acc, ent_ld_fe(e), ent_ld_fe(e + 8), 0);
        }
    }
    return acc;
}

// The share of ONE half-scalar in the ladders above and below, for the small-batch kernels: the half-scalars of an item run
// on neighbouring lanes, each with all the doublings and a quarter (Straus) or half of the additions, and the lanes' points
// are added at the end.  The result lives on the table's isomorphic curve: the caller multiplies Z by Zg after the sum.
PLUME_DEV jac vb_ladder_half(const glv_half& h, bool endo, const uint32_t* tab) {
    booth_reg b = booth_init(h);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int i = 32; i >= 0; i--) {
        if (i < 32) {
#pragma unroll 1
            for (int j = 0; j < 4; j++) acc = jac_dbl_fast(acc);
        }
        acc = vb_add_digit(acc, booth_next(b), h.neg, endo, tab);
    }
    return acc;
}

// k1 * P1 + k2 * P2 with shared doublings (Straus): both tables must be expressed on the same
// isomorphic curve (common denominator zg), see vb_build_table_pair.
PLUME_DEV jac vb_mul2_tab(const sc& k1, const uint32_t* tab1, const sc& k2, const uint32_t* tab2, const fe& zg) {
    glv_half h[4];
    glv_split(k1, h[0], h[1]);
    glv_split(k2, h[2], h[3]);
    booth_reg b0 = booth_init(h[0]), b1 = booth_init(h[1]), b2 = booth_init(h[2]), b3 = booth_init(h[3]);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int i = 32; i >= 0; i--) {
        if (i < 32) {   // the accumulator is still the identity in the top window: nothing to double
#pragma unroll 1
            for (int j = 0; j < 4; j++) acc = jac_dbl_fast(acc);
        }
        int d0 = booth_next(b0), d1 = booth_next(b1), d2 = booth_next(b2), d3 = booth_next(b3);
        // one addition body for the four half-scalars (rolled on purpose: code size)
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            int d = q == 0 ? d0 : q == 1 ? d1 : q == 2 ? d2 : d3;
            uint32_t flip = q == 0 ? h[0].neg : q == 1 ? h[1].neg : q == 2 ? h[2].neg : h[3].neg;
            acc = vb_add_digit(acc, d, flip, (q & 1) != 0, q < 2 ? tab1 : tab2);
        }
    }
    if (!acc.inf) acc.z = fe_mul(acc.z, zg);
    return acc;
}

// Tables of two affine, on-curve, non-identity points over one common denominator.  P1's table is
// built first (denominator z1); P2 is moved onto P1's isomorphic curve (x*z1^2, y*z1^3) and its table
// built there (relative denominator z2); P1's entries are then rescaled by z2.  Returns z1*z2.
PLUME_DEV fe vb_build_table_pair(const fe& p1x, const fe& p1y, uint32_t* tab1, const fe& p2x, const fe& p2y, uint32_t* tab2) {
    fe z1 = vb_build_table(p1x, p1y, tab1, false);
    fe z1_2 = fe_sqr(z1);
    fe qx = fe_mul(p2x, z1_2);
    fe qy = fe_mul(p2y, fe_mul(z1_2, z1));
    fe z2 = vb_build_table(qx, qy, tab2, true);
    fe x0 = ent_ld_fe(tab1), y0 = ent_ld_fe(tab1 + 8);
    fe z2_2 = fe_sqr(z2), z2_3 = fe_mul(z2_2, z2);
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        uint32_t* t = tab1 + e * VB_ENT_WORDS;
        const fe xc = x0, yc = y0;
        if (e < 7) { x0 = ent_ld_fe(t + VB_ENT_WORDS); y0 = ent_ld_fe(t + VB_ENT_WORDS + 8); }   // next entry: in flight during this one
        ent_finish(t, fe_mul(xc, z2_2), fe_mul(yc, z2_3));
    }
    return fe_mul(z1, z2);
}

// ---------------------------------------------------------------------------------------------
// Signed comb for SEVERAL scalars on ONE variable base (the signer's h^r and h^sk).
//
// T teeth at 2^(D j), j = 0..T-1 (T D >= 130 bits).  An odd magnitude m = sum_i b_i 2^i with every b_i in {-1, +1}
// (b_i = +1 iff bit i+1 of m is set, the top one +1) is  sum_{c<D} 2^c * (b_c T_0 + b_{c+D} T_1 + ... + b_{c+(T-1)D} T_{T-1})
// with T_j = 2^(D j) P: D doublings and D additions of a table entry  +-(T_{T-1} +- ... +- T_0)  per half-scalar, the
// (T-1) D doublings that produce the teeth being paid ONCE for all scalars on this base.  T = 5 (D = 27), two scalars with
// GLV halves: 108 + 2*27 doublings and 4*27 additions, against 2*132 doublings and 4*33 additions of the windowed ladder
// (and 99 + 66 doublings, 132 additions of the T = 4 comb of round 1).  An even magnitude is made odd by adding one, and
// the base is subtracted again at the end.
//
// The 2^(T-1) entries (top tooth positive) are built as a binary tree of CONJUGATE additions: the teeth are first brought
// to one denominator (affine points of an isomorphic curve), then level L turns every node N into N - T_j and N + T_j
// with one jac_conj_add_aff (13 field operations for the two, and the two share their Z): 1 + 2 + ... + 2^(T-2) conjugate
// additions in all, against one full Jacobian addition (16) per entry before.  The 2^(T-2) distinct Z's of the leaves are
// then multiplied into the common denominator Zg like the window table above.
//
// Storage per item (global scratch): the table (COMB_ENTRIES lines of 128 bytes) followed by COMB_SCRATCH_WORDS of
// working storage (the teeth and two level buffers of the tree).
// ---------------------------------------------------------------------------------------------
#ifndef PLUME_COMB_T
#define PLUME_COMB_T 5
#endif
#define COMB_T PLUME_COMB_T
#define COMB_D ((130 + COMB_T - 1) / COMB_T)
#define COMB_BITS (COMB_T * COMB_D)
#define COMB_ENTRIES (1 << (COMB_T - 1))
#define COMB_NZ (1 << (COMB_T - 2))                     // distinct Z's among the leaves
#define COMB_TAB_WORDS (COMB_ENTRIES * VB_ENT_WORDS)
#define COMB_LEVEL_WORDS (40 << (COMB_T - 3))           // largest inner level: 2^(T-2) points (16 words) + 2^(T-3) z (8 words)
#define COMB_SCRATCH_WORDS ((16 * COMB_T + 24 * (COMB_T - 1) + 2 * COMB_LEVEL_WORDS + 31) / 32 * 32)
#define COMB_AREA_WORDS (COMB_TAB_WORDS + COMB_SCRATCH_WORDS)
static_assert(COMB_T >= 3 && COMB_T <= 6, "comb teeth");
static_assert(COMB_BITS >= 130 && COMB_BITS <= 160 && COMB_D <= 63, "comb geometry");

PLUME_DEV void comb_st_jac(uint32_t* p, const jac& a) { ent_st_fe(p, a.x); ent_st_fe(p + 8, a.y); ent_st_fe(p + 16, a.z); }

// Builds the table for the affine, on-curve, non-identity point (px, py) of prime order; returns Zg (the product of
// every denominator introduced on the way: the entries are (x Zg^2, y Zg^3) of true affine points).
PLUME_DEV fe comb_build_table(const fe& px, const fe& py, uint32_t* area) {
    uint32_t* tab = area;
    uint32_t* teeth = area + COMB_TAB_WORDS;               // affine teeth on the common curve: tooth j at 16 j
    uint32_t* J = teeth + 16 * COMB_T;                     // Jacobian teeth T_1 .. T_{T-1}, 24 words each
    uint32_t* buf[2] = {J + 24 * (COMB_T - 1), J + 24 * (COMB_T - 1) + COMB_LEVEL_WORDS};
    // 1. the doubling chain
    {
        jac cur;
        cur.x = px; cur.y = py; cur.z = fe_one(); cur.inf = 0;
#pragma unroll 1
        for (int j = 1; j < COMB_T; j++) {
#pragma unroll 1
            for (int i = 0; i < COMB_D; i++) cur = jac_dbl_fast(cur);
            comb_st_jac(J + (j - 1) * 24, cur);
        }
    }
    // 2. teeth onto one curve: Zc = Z_1 ... Z_{T-1}, tooth j scaled by f_j = Zc / Z_j (tooth 0 by Zc)
    fe zc;
    {
        fe pre[COMB_T - 1];
        fe acc = ent_ld_fe(J + 16);
        pre[0] = fe_one();
#pragma unroll 1
        for (int j = 2; j < COMB_T; j++) { pre[j - 1] = acc; acc = fe_mul(acc, ent_ld_fe(J + (j - 1) * 24 + 16)); }
        zc = acc;
        fe suf = fe_one();
#pragma unroll 1
        for (int j = COMB_T - 1; j >= 0; j--) {
            fe f, X, Y;
            if (j == 0) {
                f = zc; X = px; Y = py;
            } else {
                const uint32_t* t = J + (j - 1) * 24;
                fe zj = ent_ld_fe(t + 16);
                f = (j == COMB_T - 1) ? pre[j - 1] : (j == 1 ? suf : fe_mul(pre[j - 1], suf));
                suf = (j == COMB_T - 1) ? zj : fe_mul(suf, zj);
                X = ent_ld_fe(t); Y = ent_ld_fe(t + 8);
            }
            fe f2 = fe_sqr(f);
            ent_st_fe(teeth + 16 * j, fe_mul(X, f2));
            ent_st_fe(teeth + 16 * j + 8, fe_mul(Y, fe_mul(f2, f)));
        }
    }
    // 3. the tree.  Level L (1 .. T-1) adds -+ tooth T-1-L to every node of level L-1; node k's children are 2k (minus)
    //    and 2k+1 (plus), so a leaf's index is its sign pattern t_{T-2} .. t_0 (1 = same sign as the top tooth).
    fe zs[COMB_NZ];
#pragma unroll 1
    for (int L = 1; L < COMB_T; L++) {
        const uint32_t* q = teeth + 16 * (COMB_T - 1 - L);
        const fe qx = ent_ld_fe(q), qy = ent_ld_fe(q + 8);
        const uint32_t* src = buf[(L - 1) & 1];            // level L-1: 2^(L-1) points, then 2^(L-2) z's (level 0: the top tooth)
        uint32_t* dst = buf[L & 1];
        const int parents = 1 << (L - 1);
#pragma unroll 1
        for (int k = 0; k < parents; k++) {
            fe X, Y, Z;
            if (L == 1) {
                X = ent_ld_fe(teeth + 16 * (COMB_T - 1)); Y = ent_ld_fe(teeth + 16 * (COMB_T - 1) + 8); Z = fe_one();
            } else {
                X = ent_ld_fe(src + 16 * k); Y = ent_ld_fe(src + 16 * k + 8); Z = ent_ld_fe(src + 16 * parents + 8 * (k >> 1));
            }
            jac_pair r = jac_conj_add_aff(X, Y, Z, L == 1, qx, qy);
            if (L < COMB_T - 1) {
                ent_store_xy(dst + 16 * (2 * k), r.xd, r.yd);
                ent_store_xy(dst + 16 * (2 * k + 1), r.xs, r.ys);
                ent_st_fe(dst + 16 * (2 * parents) + 8 * k, r.z);
            } else {
                ent_store_xy(tab + (2 * k) * VB_ENT_WORDS, r.xd, r.yd);
                ent_store_xy(tab + (2 * k + 1) * VB_ENT_WORDS, r.xs, r.ys);
                zs[k] = r.z;
            }
        }
    }
    // 4. common denominator Zg = z_0 ... z_{NZ-1}: the two entries of pair k are scaled by f = Zg / z_k (x f^2, y f^3)
    fe pre[COMB_NZ];                              // pre[k] = z_0 ... z_{k-1}
    fe acc = zs[0];
    pre[0] = fe_one();
#pragma unroll 1
    for (int k = 1; k < COMB_NZ; k++) { pre[k] = acc; acc = fe_mul(acc, zs[k]); }
    const fe zg = acc;
    fe suf = fe_one();                            // z_{k+1} ... z_{NZ-1}
#pragma unroll 1
    for (int k = COMB_NZ - 1; k >= 0; k--) {
        uint32_t* e0 = tab + (2 * k) * VB_ENT_WORDS;
        const fe xa = ent_ld_fe(e0), ya = ent_ld_fe(e0 + 8);   // raw leaves, back from the L2: requested before the products below
        const fe xb = ent_ld_fe(e0 + VB_ENT_WORDS), yb = ent_ld_fe(e0 + VB_ENT_WORDS + 8);
        fe f = (k == COMB_NZ - 1) ? pre[k] : (k == 0 ? suf : fe_mul(pre[k], suf));
        suf = (k == COMB_NZ - 1) ? zs[k] : fe_mul(suf, zs[k]);
        fe f2 = fe_sqr(f), f3 = fe_mul(f2, f);
        ent_finish(e0, fe_mul(xa, f2), fe_mul(ya, f3));
        ent_finish(e0 + VB_ENT_WORDS, fe_mul(xb, f2), fe_mul(yb, f3));
    }
    return fe_mul(zc, zg);
}

// rows of the signed-digit matrix of one half-scalar: bit c of row[j] = 1 iff b_{c + D j} = +1
struct comb_rows { uint64_t row[COMB_T]; uint32_t neg; uint32_t even; };
PLUME_DEV comb_rows comb_recode(const glv_half& h) {
    // m' = m | 1 (m + 1 when m is even); digits from m'' = (m' >> 1) | 2^(T D - 1)
    uint32_t w[6];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = h.mag[i];
    w[5] = 0;
    comb_rows r;
    r.even = (w[0] & 1) ^ 1;
    r.neg = h.neg;
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
    w[(COMB_BITS - 1) >> 5] |= 1u << ((COMB_BITS - 1) & 31);
    // row j = bits [D j, D j + D)
#pragma unroll
    for (int j = 0; j < COMB_T; j++) {
        const int lo = COMB_D * j, wi = lo >> 5, sh = lo & 31;
        uint64_t v = ((uint64_t)w[wi] >> sh);
        if (wi + 1 < 6) v |= (uint64_t)w[wi + 1] << (32 - sh);
        if (wi + 2 < 6 && sh != 0) v |= (uint64_t)w[wi + 2] << (64 - sh);
        r.row[j] = v & ((1ull << COMB_D) - 1);
    }
    return r;
}

// acc += (digit column c of rows) * table, on the second GLV half through the endomorphism
PLUME_DEV jac comb_add_column(const jac& acc, const comb_rows& r, int c, bool endo, const uint32_t* tab) {
    const uint32_t top = (uint32_t)(r.row[COMB_T - 1] >> c) & 1;
    uint32_t e = 0;
#pragma unroll
    for (int j = 0; j < COMB_T - 1; j++) e |= ((((uint32_t)(r.row[j] >> c) & 1) == top) ? 1u : 0u) << j;
    const bool neg = ((top ^ 1) ^ r.neg) != 0;
    fe x, y;
    ent_load(tab + e * VB_ENT_WORDS, endo, neg, x, y);
    return jac_add_aff_fast(acc, x, y);
}

// k * P from the prepared comb table; k canonical in [0, n).  The parity correction needs P itself, expressed with the
// table's denominator: (pxs, pys) = (px zg^2, py zg^3).
PLUME_DEV jac comb_mul_tab(const sc& k, const uint32_t* tab, const fe& zg, const fe& pxs, const fe& pys) {
    glv_half h1, h2;
    glv_split(k, h1, h2);
    comb_rows r1 = comb_recode(h1), r2 = comb_recode(h2);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int c = COMB_D - 1; c >= 0; c--) {
        if (c < COMB_D - 1) acc = jac_dbl_fast(acc);   // (the identity in the top column: nothing to double)
#pragma unroll 1
        for (int h = 0; h < 2; h++) acc = comb_add_column(acc, h ? r2 : r1, c, h != 0, tab);
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

// One half-scalar's share of comb_mul_tab (small-batch kernels; see vb_ladder_half), parity correction included, on the
// table's isomorphic curve.
PLUME_DEV jac comb_ladder_half(const glv_half& h, bool endo, const uint32_t* tab, const fe& pxs, const fe& pys) {
    comb_rows r = comb_recode(h);
    jac acc = jac_infinity();
#pragma unroll 1
    for (int c = COMB_D - 1; c >= 0; c--) {
        if (c < COMB_D - 1) acc = jac_dbl_fast(acc);
        acc = comb_add_column(acc, r, c, endo, tab);
    }
    if (r.even) {
        fe x = endo ? fe_mul(pxs, ec_beta()) : pxs;
        fe y = r.neg ? pys : fe_neg(pys);
        acc = jac_add_aff(acc, x, y, 0);
    }
    return acc;
}

// scratch words per item: the signer's comb area or the verifier's three window tables (h and nul for h*s - nul*c, pk for
// G*s - pk*c: separate, so that the two may run side by side), whichever is larger
#define VB_ITEM_WORDS (COMB_AREA_WORDS > 3 * VB_TAB_WORDS ? COMB_AREA_WORDS : 3 * VB_TAB_WORDS)

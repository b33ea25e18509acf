// stages_team.cuh -- the stages of stages.cuh for SMALL batches: a team of 2 or 4 neighbouring lanes per item.
//
// The throughput kernels give one thread one item (or one scalar), which is what fills the machine at 10^5 items and more.
// A batch of one -- the call shape of the reference's sign_v1 / sign_v2 / verify (rust-k256/src/lib.rs:149-156, :99) -- then
// runs ~3 700 field operations one after the other on one lane of one warp.  These bodies shorten that chain by giving the
// independent parts of a stage to neighbouring lanes and adding the lanes' points with shuffles at the end:
//
//   hash_to_curve          2 lanes: the two SSWU maps (one 254-squaring exponentiation each); also plume_hash_to_curve_batch
//   g^r, g^sk              2 lanes: one scalar each
//   h^r, h^sk (comb)       4 lanes: (scalar, GLV half); all the doublings and half the additions each
//   h*s - nul*c            4 lanes: one of the four half-scalars each (all doublings, a quarter of the additions); the two
//                          window tables built side by side by two of them, in the same kernel
//   G*s - pk*c             4 lanes: the two GLV halves of -c*pk, and the two halves of the generator windows of s; on the
//                          second stream from the start of the call
//
// Results are the same points in another Jacobian representation, so the outputs (affine, canonical) are bit-identical to
// the throughput path's; tests/test_gpu_parity.py runs both on the same inputs against the oracle.  Device code only.
#pragma once
#include "stages.cuh"

// `mask`: the warp's lanes whose team is inside the batch (teams are whole: the kernels are launched on T * n threads).

PLUME_DEV void sign_stage_fixed_team(uint32_t idx, const sign_args& a) {   // 2 lanes, no exchange
    const uint32_t i = idx >> 1, j = idx & 1;
    sc r = ld_sc_be(a.r + (size_t)i * 32);
    sc sk = ld_sc_be(a.sk + (size_t)i * 32);
    uint8_t st = PLUME_ST_OK;
    if (a.flavour == PLUME_FLAVOUR_ARKWORKS) {
        if (sc_ge_n(sk)) { st = PLUME_ST_BAD_SK; sk = sc_one(); }
        if (sc_ge_n(r)) { st = PLUME_ST_BAD_R; r = sc_one(); }
        aff P;
        if (!ld_point_be(P, a.pk_in + (size_t)i * 64) || P.inf) { st = PLUME_ST_BAD_PK; P = aff_generator(); }
        if (j == 0) {
            a.status[i] = st;
            ws_store_jac(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i, fb_mul(r, a.gtab, a.gw));
        } else {
            ws_store_jac(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i, jac_from_aff(P));
        }
        return;
    }
    if (!sc_is_valid_nonzero(sk)) { st = PLUME_ST_BAD_SK; sk = sc_one(); }
    if (!sc_is_valid_nonzero(r)) { st = PLUME_ST_BAD_R; r = sc_one(); }
    if (j == 0) a.status[i] = st;
    jac P = fb_mul(j == 0 ? r : sk, a.gtab, a.gw);
    st_fe(ws_at(a.ws, a.n, j == 0 ? WS_AX : WS_BX, i), P.x);
    st_fe(ws_at(a.ws, a.n, j == 0 ? WS_AY : WS_BY, i), P.y);
    st_fe(ws_at(a.ws, a.n, j == 0 ? WS_Z0 : WS_Z1, i), P.z);
}

// Workspace of the small-batch signer.  As in the verifier below, h stays Jacobian from hash_to_curve to the last stage (the
// throughput path inverts its Z before the comb table: one more inversion in the chain), so the last batched inversion has
// three elements per item -- Z of h^r, of the nullifier, of h -- in three slots PAST the 14 of the throughput path, with three
// more as its scratch: a small batch addresses the workspace with its own n as the stride, and api.cu lane_workspace sizes
// every workspace so that slots 14..19 of a batch of <= 4 096 items lie inside it.
enum { TS_ZA = WS_SLOTS, TS_ZB = WS_SLOTS + 1, TS_ZH = WS_SLOTS + 2, TS_SCRATCH = WS_SLOTS + 3, TS_SLOTS = WS_SLOTS + 6 };

PLUME_DEV void sign_stage_h2c_team(uint32_t mask, uint32_t idx, const sign_args& a) {   // 2 lanes
    const uint32_t i = idx >> 1, j = idx & 1;
    aff R = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, WS_Z0, i);
    aff K = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, WS_Z1, i);
    if (j == 0) {
        ws_store_aff(a.ws, a.n, WS_RX, WS_RY, i, R);
        ws_store_aff(a.ws, a.n, WS_KX, WS_KY, i, K);
    }
    uint8_t pk33[33];
    uint32_t npk = enc_point33(pk33, K);
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    jac h = h2c_hash_to_curve_team(mask, j, m, len, pk33, npk);
    if (j == 0) ws_store_jac(a.ws, a.n, WS_HX, WS_HY, TS_ZH, i, h);
}

// the comb table of sign_stage_varbase_tab from the Jacobian h = (X, Y, Z): (X, Y) is an affine point of the isomorphic
// curve with denominator Z, on which the table is built as usual; WS_P0 = the table's own denominator Zg (0: no ladders),
// WS_P1 = Zg * Z, what the ladders' results are multiplied by.  One lane per item.
PLUME_DEV void sign_stage_varbase_tab_jac(uint32_t i, const sign_args& a, uint32_t* area) {
    fe hz = ld_fe(ws_at(a.ws, a.n, TS_ZH, i));
    if (fe_is_zero(hz)) {
        a.status[i] = PLUME_ST_H_INF;   // the reference panics here (randomizedsigner.rs:61)
        jac o = jac_infinity();
        ws_store_jac(a.ws, a.n, WS_AX, WS_AY, TS_ZA, i, o);
        ws_store_jac(a.ws, a.n, WS_BX, WS_BY, TS_ZB, i, o);
        st_fe(ws_at(a.ws, a.n, WS_P0, i), fe_zero());
        return;
    }
    fe zg = comb_build_table(ld_fe(ws_at(a.ws, a.n, WS_HX, i)), ld_fe(ws_at(a.ws, a.n, WS_HY, i)), area);
    st_fe(ws_at(a.ws, a.n, WS_P0, i), zg);
    st_fe(ws_at(a.ws, a.n, WS_P1, i), fe_mul(zg, hz));
}

// h^r, h^sk from the comb table of sign_stage_varbase_tab_jac: lane q = 2 * (0: r, 1: sk) + (GLV half)
PLUME_DEV void sign_stage_varbase_lad_team(uint32_t mask, uint32_t idx, const sign_args& a, const uint32_t* vbtab) {
    const uint32_t i = idx >> 2, q = idx & 3, which = q >> 1;
    fe zg = ld_fe(ws_at(a.ws, a.n, WS_P0, i));
    mask = __ballot_sync(mask, !fe_is_zero(zg));
    if (fe_is_zero(zg)) return;   // h was the identity: the table stage stored the results (the whole team leaves)
    fe hx = ld_fe(ws_at(a.ws, a.n, WS_HX, i)), hy = ld_fe(ws_at(a.ws, a.n, WS_HY, i));
    fe zg2 = fe_sqr(zg);
    fe hxs = fe_mul(hx, zg2), hys = fe_mul(hy, fe_mul(zg2, zg));
    sc k = ld_sc_be((which ? a.sk : a.r) + (size_t)i * 32);
    const bool zero_ok = a.flavour == PLUME_FLAVOUR_ARKWORKS;
    if (sc_ge_n(k) || (!zero_ok && sc_is_zero(k))) k = sc_one();
    glv_half h1, h2;
    glv_split(k, h1, h2);
    jac o = comb_ladder_half((q & 1) ? h2 : h1, (q & 1) != 0, vbtab + (size_t)i * VB_ITEM_WORDS, hxs, hys);
    o = jac_team_sum(mask, o, q & 1, 1);
    if (q & 1) return;
    if (!o.inf) o.z = fe_mul(o.z, ld_fe(ws_at(a.ws, a.n, WS_P1, i)));
    if (which == 0) ws_store_jac(a.ws, a.n, WS_AX, WS_AY, TS_ZA, i, o);
    else ws_store_jac(a.ws, a.n, WS_BX, WS_BY, TS_ZB, i, o);
}

// last stage: one lane per item; h^r, the nullifier and h become affine here (TS_Z* hold the inverses)
PLUME_DEV void sign_stage_final_team(uint32_t i, const sign_args& a) {
    aff z = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, TS_ZA, i);
    aff nul = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, TS_ZB, i);
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, TS_ZH, i);
    aff R = ws_load_aff_xy(a.ws, a.n, WS_RX, WS_RY, i);
    aff K = ws_load_aff_xy(a.ws, a.n, WS_KX, WS_KY, i);
    sign_final_finish(i, a, K, h, nul, R, z);
}

// Workspace of the small-batch verifier.  h stays Jacobian from hash_to_curve to the last stage (the throughput path inverts
// its Z right away to build the window tables from an affine h: one more inversion in the chain), so the ONE batched inversion
// of this path has three elements per item, in three slots the throughput path uses for other things: Z of A, of B, of h.
// Its scratch is WS_Z0 .. WS_P0, which are idle here.
enum { TV_ZA = WS_RY, TV_ZB = WS_KX, TV_ZH = WS_KY };
static_assert(TV_ZB == TV_ZA + 1 && TV_ZH == TV_ZA + 2, "contiguous Z slots");

PLUME_DEV void verify_stage_h2c_team(uint32_t mask, uint32_t idx, const verify_args& a) {   // 2 lanes
    const uint32_t i = idx >> 1, j = idx & 1;
    uint8_t pk33[33];
    bool good;
    uint32_t npk = verify_h2c_check(i, a, pk33, good);
    if (j == 0) a.ok[i] = good ? 1 : 0;
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    jac h = h2c_hash_to_curve_team(mask, j, m, len, pk33, npk);
    if (j == 0) ws_store_jac(a.ws, a.n, WS_HX, WS_HY, TV_ZH, i, h);
}

// hash_to_curve alone (h2c_stage_map), 2 lanes
PLUME_DEV void h2c_stage_map_team(uint32_t mask, uint32_t idx, const h2c_args& a) {
    const uint32_t i = idx >> 1, j = idx & 1;
    uint32_t len;
    const uint8_t* m = msg_ptr(a.msgs, i, len);
    jac h;
    if (a.pk33) {
        uint8_t e[33];
#pragma unroll 1
        for (int k = 0; k < 33; k++) e[k] = a.pk33[(size_t)i * 33 + k];
        h = h2c_hash_to_curve_team(mask, j, m, len, e, e[0] == 0 ? 1u : 33u);
    } else {
        h = h2c_hash_to_curve_team(mask, j, m, len, m, 0);
    }
    if (j == 0) ws_store_jac(a.ws, a.n, WS_HX, WS_HY, WS_Z0, i, h);
}

// B = s*h - c*nul, table stage and ladder stage (verify_stage_mul_b1 / _b2) in one: lane q = 2 * (0: s on h, 1: -c on nul) +
// (GLV half).  Lanes 0 and 2 build the window table of their base side by side, each over its own denominator (no rescaling
// of one table onto the other's curve, which the one-accumulator Straus ladder needs); every lane then walks one half-scalar,
// the two halves of a base are added on that base's isomorphic curve, brought back with its Zg, and the two bases added.
// h comes as the Jacobian (X, Y, Z) of hash_to_curve: (X, Y) is an affine point of the isomorphic curve with denominator Z,
// the a = 0 formulas do not see the difference, and Z joins Zg at the end.
// An identity among h, nul (adversarial inputs only) contributes the identity.
PLUME_DEV void verify_stage_mul_b_team(uint32_t mask, uint32_t idx, const verify_args& a, uint32_t* vbtab) {
    const uint32_t i = idx >> 2, q = idx & 3;
    const bool good = a.ok[i] != 0;
    aff P;
    fe zscale = fe_one();
    sc k = sc_one();
    if (q < 2) {
        P.x = ld_fe(ws_at(a.ws, a.n, WS_HX, i));
        P.y = ld_fe(ws_at(a.ws, a.n, WS_HY, i));
        zscale = ld_fe(ws_at(a.ws, a.n, TV_ZH, i));
        P.inf = fe_is_zero(zscale);
    } else if (good) {
        ld_point_be(P, a.nullifier + (size_t)i * 64);
    } else {
        P = aff_generator();
    }
    if (good) {
        k = ld_sc_be((q < 2 ? a.s : a.c) + (size_t)i * 32);
        if (q >= 2) k = sc_neg(k);
    }
    uint32_t* tab = vbtab + (size_t)i * VB_ITEM_WORDS + (q < 2 ? 0 : VB_TAB_WORDS);
    fe zg = fe_one();
    if ((q & 1) == 0 && !P.inf) {
        zg = vb_build_table(P.x, P.y, tab, true);
        if (q == 0) zg = fe_mul(zg, zscale);
    }
    __syncwarp(mask);   // the odd lane reads the table its neighbour wrote
    jac B = jac_infinity();
    if (!P.inf) {
        glv_half h1, h2;
        glv_split(k, h1, h2);
        B = vb_ladder_half((q & 1) ? h2 : h1, (q & 1) != 0, tab);
    }
    B = jac_team_sum(mask, B, q & 1, 1);
    if ((q & 1) == 0 && !B.inf) B.z = fe_mul(B.z, zg);   // (the odd lane's copy of the pair's sum is not used)
    jac o = jac_shfl_xor(mask, B, 2);
    if (q != 0) return;
    ws_store_jac(a.ws, a.n, WS_BX, WS_BY, TV_ZB, i, jac_add(B, o));
}

// A = s*G - c*pk: lanes 0, 1 the GLV halves of -c on pk's window table (lane 0 builds it), lanes 2, 3 the lower and upper
// generator windows of s.  Needs nothing from the other stages (it repeats the input checks instead of reading ok[]), so it
// runs on the second stream from the start of the call.
PLUME_DEV void verify_stage_mul_a_team(uint32_t mask, uint32_t idx, const verify_args& a, uint32_t* tab) {
    const uint32_t i = idx >> 2, q = idx & 3;
    uint8_t pk33[33];
    bool good;
    verify_h2c_check(i, a, pk33, good);
    aff pk;
    sc c = sc_one(), s = sc_one();
    if (good) {
        ld_point_be(pk, a.pk + (size_t)i * 64);
        c = ld_sc_be(a.c + (size_t)i * 32);
        s = ld_sc_be(a.s + (size_t)i * 32);
    } else {
        pk = aff_generator();
    }
    fe zg = fe_one();
    if (q == 0 && !pk.inf) zg = vb_build_table(pk.x, pk.y, tab, true);
    __syncwarp(mask);    // the table is in global scratch: ordered for the team's lane 1 by the barrier
    jac A;
    if (q < 2) {
        A = jac_infinity();
        if (!pk.inf) {
            glv_half h1, h2;
            glv_split(sc_neg(c), h1, h2);
            A = vb_ladder_half(q ? h2 : h1, q != 0, tab);
        }
    } else {
        const int nw = fb_windows(a.gw), mid = nw / 2;
        A = fb_mul_windows(s, a.gtab, a.gw, q == 2 ? 0 : mid, q == 2 ? mid : nw);
    }
    A = jac_team_sum(mask, A, q & 1, 1);
    if (q == 0 && !A.inf) A.z = fe_mul(A.z, zg);   // back from the table's isomorphic curve (only lane 0's sum is used)
    jac o = jac_shfl_xor(mask, A, 2);
    if (q != 0) return;
    A = jac_add(o, A);    // fixed-base part first, as verify_stage_mul_a does
    ws_store_jac(a.ws, a.n, WS_AX, WS_AY, TV_ZA, i, A);
}

// last stage: one lane per item; A, B and h become affine here (TV_Z* hold the inverses)
PLUME_DEV void verify_stage_final_team(uint32_t i, const verify_args& a) {
    if (a.ok[i] == 0) return;
    aff A = ws_load_affine(a.ws, a.n, WS_AX, WS_AY, TV_ZA, i);
    aff B = ws_load_affine(a.ws, a.n, WS_BX, WS_BY, TV_ZB, i);
    aff h = ws_load_affine(a.ws, a.n, WS_HX, WS_HY, TV_ZH, i);
    verify_final_check(i, a, A, B, h);
}

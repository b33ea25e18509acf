// ec.cuh -- secp256k1 group law (y^2 = x^3 + 7) in Jacobian coordinates, SEC1 helpers.
//
// Replaces k256::ProjectivePoint / AffinePoint arithmetic that the reference calls at
// rust-k256/src/randomizedsigner.rs:51,53,67,70 (scalar multiplications) and
// rust-k256/src/lib.rs:101,109 (the verifier's two mul-sub combinations); generator per
// rust-arkworks/src/secp256k1/curves/mod.rs:50-58.
//
// All formulas are complete for the inputs they can meet here: the exceptional cases of the
// addition (P + P, P + (-P), identity operands) branch to the right answer instead of assuming
// they cannot happen, because a verifier is fed adversarial points and scalars.
#pragma once
#include "fe.cuh"

struct aff { fe x, y; uint32_t inf; };
struct jac { fe x, y, z; uint32_t inf; };

PLUME_DEV fe fe_lit(uint32_t w7, uint32_t w6, uint32_t w5, uint32_t w4, uint32_t w3, uint32_t w2, uint32_t w1, uint32_t w0) {
    fe r;
    r.v[0] = w0; r.v[1] = w1; r.v[2] = w2; r.v[3] = w3; r.v[4] = w4; r.v[5] = w5; r.v[6] = w6; r.v[7] = w7;
    return r;
}
PLUME_DEV fe ec_gx() { return fe_lit(0x79BE667Eu, 0xF9DCBBACu, 0x55A06295u, 0xCE870B07u, 0x029BFCDBu, 0x2DCE28D9u, 0x59F2815Bu, 0x16F81798u); }
PLUME_DEV fe ec_gy() { return fe_lit(0x483ADA77u, 0x26A3C465u, 0x5DA4FBFCu, 0x0E1108A8u, 0xFD17B448u, 0xA6855419u, 0x9C47D08Fu, 0xFB10D4B8u); }
// beta: cube root of unity in Fp with lambda*(x, y) = (beta*x, y)
PLUME_DEV fe ec_beta() { return fe_lit(0x7AE96A2Bu, 0x657C0710u, 0x6E64479Eu, 0xAC3434E9u, 0x9CF04975u, 0x12F58995u, 0xC1396C28u, 0x719501EEu); }

PLUME_DEV jac jac_infinity() { jac r; r.x = fe_zero(); r.y = fe_zero(); r.z = fe_zero(); r.inf = 1; return r; }
PLUME_DEV aff aff_infinity() { aff r; r.x = fe_zero(); r.y = fe_zero(); r.inf = 1; return r; }
PLUME_DEV jac jac_from_aff(const aff& p) { jac r; r.x = p.x; r.y = p.y; r.z = fe_one(); r.inf = p.inf; return r; }

// y^2 == x^3 + 7 ?
PLUME_DEV bool aff_on_curve(const fe& x, const fe& y) {
    fe lhs = fe_sqr(y);
    fe rhs = fe_add(fe_mul(fe_sqr(x), x), fe_set_u32(7));
    return fe_eq(lhs, rhs);
}

// Call granularity.  Default: fe_mul / fe_sqr are the out-of-line units and the point formulas are inlined
// into their callers.  -DPLUME_POINT_FN makes jac_dbl / jac_add_aff the out-of-line units instead, with the
// field multiplications inlined into them (no operand marshalling per multiplication, 7-11 per point call).
#ifdef PLUME_POINT_FN
#define EC_MUL fe_mul_inl
#define EC_SQR fe_sqr_inl
#define PLUME_POINTFN PLUME_DEV_NOINLINE
#else
#define EC_MUL fe_mul
#define EC_SQR fe_sqr
#define PLUME_POINTFN PLUME_DEV
#endif

// 2P, a = 0: 2M + 5S  (no point of order two exists: the group order is odd).
// Statement order is chosen for short live ranges (the multiplier is an opaque call to the compiler,
// so it keeps this order): at most five field elements are alive at any point.
PLUME_POINTFN jac jac_dbl(jac p) {
    if (p.inf) return p;
    jac r;
    r.z = fe_dbl(EC_MUL(p.y, p.z));          // Z3 = 2*Y*Z          (Z dead)
    fe A = EC_SQR(p.x);
    fe B = EC_SQR(p.y);                      //                      (Y dead)
    fe t = EC_SQR(fe_add(p.x, B));           //                      (X dead)
    fe C = EC_SQR(B);                        //                      (B dead)
    fe D = fe_dbl(fe_sub(fe_sub(t, A), C));  // 2*((X+B)^2 - A - C)  (t dead)
    fe E = fe_add(fe_dbl(A), A);             // 3*A                  (A dead)
    r.x = fe_sub(EC_SQR(E), fe_dbl(D));      // E^2 - 2*D
    fe C8 = fe_shl<3>(C);
    r.y = fe_sub(EC_MUL(E, fe_sub(D, r.x)), C8);
    r.inf = 0;
    return r;
}

// The same doubling for the ladders: no test of the identity flag.  The identity is always stored as (0, 0, 0) with the
// flag set, and the formulas map (0, 0, 0) to (0, 0, 0), so the flag is simply carried along (the generic version above
// compiles its early return into selects over all 24 result registers).
PLUME_POINTFN jac jac_dbl_fast(jac p) {
    jac r;
    r.z = fe_dbl(EC_MUL(p.y, p.z));
    fe A = EC_SQR(p.x);
    fe B = EC_SQR(p.y);
    fe t = EC_SQR(fe_add(p.x, B));
    fe C = EC_SQR(B);
    fe D = fe_dbl(fe_sub(fe_sub(t, A), C));
    fe E = fe_add(fe_dbl(A), A);
    r.x = fe_sub(EC_SQR(E), fe_dbl(D));
    r.y = fe_sub(EC_MUL(E, fe_sub(D, r.x)), fe_shl<3>(C));
    r.inf = p.inf;
    return r;
}

// P + Q, Q affine: 8M + 3S, ordered for short live ranges as well.
PLUME_POINTFN jac jac_add_aff(jac p, fe qx, fe qy, uint32_t qinf) {
    if (qinf) return p;
    if (p.inf) { jac r; r.x = qx; r.y = qy; r.z = fe_one(); r.inf = 0; return r; }
    fe z2 = EC_SQR(p.z);
    fe H = fe_sub(EC_MUL(qx, z2), p.x);                  // U2 - X1            (qx dead)
    fe R = fe_sub(EC_MUL(qy, EC_MUL(p.z, z2)), p.y);     // S2 - Y1            (qy, z2 dead)
    if (fe_is_zero(H)) {
        if (fe_is_zero(R)) return jac_dbl(p);
        return jac_infinity();
    }
    jac r;
    r.z = EC_MUL(p.z, H);                                //                    (Z1 dead)
    fe H2 = EC_SQR(H);
    fe H3 = EC_MUL(H, H2);                               //                    (H dead)
    fe V = EC_MUL(p.x, H2);                              //                    (X1, H2 dead)
    r.x = fe_sub(fe_sub(EC_SQR(R), H3), fe_dbl(V));
    r.y = fe_sub(EC_MUL(R, fe_sub(V, r.x)), EC_MUL(p.y, H3));
    r.inf = 0;
    return r;
}

// P + Q for the ladders: Q affine and finite (a table entry), P anything.  The only test on the common path is H == 0,
// which the identity (0, 0, 0) also satisfies (H = qx * 0 - 0), so the identity operand, P = Q and P = -Q all leave
// through the same rare branch.
PLUME_POINTFN jac jac_add_aff_fast(jac p, fe qx, fe qy) {
    fe z2 = EC_SQR(p.z);
    fe H = fe_sub(EC_MUL(qx, z2), p.x);
    fe R = fe_sub(EC_MUL(qy, EC_MUL(p.z, z2)), p.y);
    if (fe_is_zero(H)) {
        if (p.inf) { jac r; r.x = qx; r.y = qy; r.z = fe_one(); r.inf = 0; return r; }
        if (fe_is_zero(R)) return jac_dbl(p);
        return jac_infinity();
    }
    jac r;
    r.z = EC_MUL(p.z, H);
    fe H2 = EC_SQR(H);
    fe H3 = EC_MUL(H, H2);
    fe V = EC_MUL(p.x, H2);
    r.x = fe_sub(fe_sub(EC_SQR(R), H3), fe_dbl(V));
    r.y = fe_sub(EC_MUL(R, fe_sub(V, r.x)), EC_MUL(p.y, H3));
    r.inf = 0;
    return r;
}

// Conjugate addition: P + Q and P - Q for Q affine, 9M + 4S for the pair instead of 2 x (8M + 3S), and both results come
// out with the SAME Z (= Z_P * H), which halves the work of bringing a table to a common denominator.  Only for operands
// known to be finite with P != +-Q (the signed-comb table: odd combinations of the teeth of a prime-order point).
struct jac_pair { fe xs, ys, xd, yd, z; };   // sum (xs, ys), difference (xd, yd), shared z
PLUME_DEV jac_pair jac_conj_add_aff(const fe& px, const fe& py, const fe& pz, bool pz_is_one, const fe& qx, const fe& qy) {
    fe H, S2;
    jac_pair r;
    if (pz_is_one) {
        H = fe_sub(qx, px);
        S2 = qy;
        r.z = H;
    } else {
        fe z2 = fe_sqr(pz);
        H = fe_sub(fe_mul(qx, z2), px);
        S2 = fe_mul(qy, fe_mul(pz, z2));
        r.z = fe_mul(pz, H);
    }
    fe H2 = fe_sqr(H);
    fe H3 = fe_mul(H, H2);
    fe V = fe_mul(px, H2);
    fe V2 = fe_dbl(V);
    fe YH3 = fe_mul(py, H3);
    fe Rs = fe_sub(S2, py);                 // S2 - Y1
    fe Rd = fe_neg(fe_add(S2, py));         // -S2 - Y1
    r.xs = fe_sub(fe_sub(fe_sqr(Rs), H3), V2);
    r.ys = fe_sub(fe_mul(Rs, fe_sub(V, r.xs)), YH3);
    r.xd = fe_sub(fe_sub(fe_sqr(Rd), H3), V2);
    r.yd = fe_sub(fe_mul(Rd, fe_sub(V, r.xd)), YH3);
    return r;
}

// P + Q, both Jacobian: 12M + 4S
PLUME_DEV jac jac_add(const jac& p, const jac& q) {
    if (q.inf) return p;
    if (p.inf) return q;
    fe z1z1 = fe_sqr(p.z), z2z2 = fe_sqr(q.z);
    fe u1 = fe_mul(p.x, z2z2), u2 = fe_mul(q.x, z1z1);
    fe s1 = fe_mul(p.y, fe_mul(q.z, z2z2)), s2 = fe_mul(q.y, fe_mul(p.z, z1z1));
    fe H = fe_sub(u2, u1);
    fe R = fe_sub(s2, s1);
    if (fe_is_zero(H)) {
        if (fe_is_zero(R)) return jac_dbl(p);
        return jac_infinity();
    }
    fe H2 = fe_sqr(H);
    fe H3 = fe_mul(H, H2);
    fe V = fe_mul(u1, H2);
    jac r;
    r.x = fe_sub(fe_sub(fe_sqr(R), H3), fe_dbl(V));
    r.y = fe_sub(fe_mul(R, fe_sub(V, r.x)), fe_mul(s1, H3));
    r.z = fe_mul(fe_mul(p.z, q.z), H);
    r.inf = 0;
    return r;
}

PLUME_DEV jac jac_neg(const jac& p) { jac r = p; r.y = fe_neg(p.y); return r; }

#ifndef PLUME_HOSTSIM
// Lane exchange for the small-batch ("team") kernels, in which the 2 or 4 neighbouring lanes of a warp that belong to one
// item each compute a share of a sum of points.  `mask` names the warp's live lanes (whole teams; __ballot_sync at kernel
// entry, before the lanes past the end of the batch leave).
PLUME_DEV fe fe_shfl_xor(uint32_t mask, const fe& a, int m) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(mask, a.v[i], m);
    return r;
}
PLUME_DEV jac jac_shfl_xor(uint32_t mask, const jac& p, int m) {
    jac r;
    r.x = fe_shfl_xor(mask, p.x, m); r.y = fe_shfl_xor(mask, p.y, m); r.z = fe_shfl_xor(mask, p.z, m);
    r.inf = __shfl_xor_sync(mask, p.inf, m);
    return r;
}
// the sum of the points of a team of 2^levels lanes; `q` = the lane's index in its team.  Every lane adds in the order the
// team's lane 0 does (lower lane's point first), so all lanes end with the same representation of the sum.
PLUME_DEV jac jac_team_sum(uint32_t mask, jac p, uint32_t q, int levels) {
    for (int l = 0; l < levels; l++) {
        jac o = jac_shfl_xor(mask, p, 1 << l);
        p = ((q >> l) & 1) ? jac_add(o, p) : jac_add(p, o);
    }
    return p;
}
#endif

// affine from Jacobian given zinv = 1/Z (canonical coordinates)
PLUME_DEV aff aff_from_jac_zinv(const jac& p, const fe& zinv) {
    aff r;
    if (p.inf) return aff_infinity();
    fe zi2 = fe_sqr(zinv);
    r.x = fe_norm(fe_mul(p.x, zi2));
    r.y = fe_norm(fe_mul(p.y, fe_mul(zi2, zinv)));
    r.inf = 0;
    return r;
}

// affine (canonical) equality including the identity
PLUME_DEV bool aff_eq(const aff& a, const aff& b) {
    if (a.inf || b.inf) return (a.inf != 0) == (b.inf != 0);
    return fe_eq(a.x, b.x) && fe_eq(a.y, b.y);
}

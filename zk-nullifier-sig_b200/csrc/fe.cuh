// fe.cuh -- arithmetic in Fp, p = 2^256 - 2^32 - 977 (secp256k1 base field).
//
// Representation: 8 x 32-bit little-endian limbs held in registers, value "weakly reduced":
// any representative in [0, 2^256).  fe_norm() gives the canonical one in [0, p).
// Reduction is the pseudo-Mersenne fold 2^256 == C (mod p), C = 2^32 + 977 -- no Montgomery
// form (SURVEY.md section 7 allows either; the fold costs 8 extra limb-products per multiplication
// instead of 64).  One multiplication = 64 + 8 + 2 IMAD.WIDE.U32 issued as carry chains
// (see ptx.cuh); additions/subtractions ride on the ALU pipe.
//
// This replaces what the reference obtains from k256::FieldElement (external crate k256 ~0.13,
// rust-k256/Cargo.toml:18); constants p / b per rust-arkworks/src/secp256k1/fields/fq.rs:12.
#pragma once
#include "ptx.cuh"

struct fe { uint32_t v[8]; };

#define FE_C0 977u  // C = 2^32 + FE_C0

PLUME_DEV fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
PLUME_DEV fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }
PLUME_DEV fe fe_set_u32(uint32_t x) { fe r = fe_zero(); r.v[0] = x; return r; }

// r = a + b
PLUME_DEV fe fe_add(const fe& a, const fe& b) {
    fe r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    uint32_t co = addc(0, 0);
    // 2^256 == C: add co*C (can itself wrap once more when the operands were non-canonical)
    r.v[0] = add_cc(r.v[0], (0u - co) & FE_C0);
    r.v[1] = addc_cc(r.v[1], co);
#pragma unroll
    for (int i = 2; i < 8; i++) r.v[i] = addc_cc(r.v[i], 0);
    uint32_t co2 = addc(0, 0);
    r.v[0] = add_cc(r.v[0], (0u - co2) & FE_C0);  // wrapped value < C here: touches limbs 0,1 only
    r.v[1] = addc(r.v[1], co2);
    return r;
}

// r = a - b
PLUME_DEV fe fe_sub(const fe& a, const fe& b) {
    fe r;
    r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
    uint32_t bo = subc(0, 0) & 1;  // 0xffffffff -> 1
    r.v[0] = sub_cc(r.v[0], (0u - bo) & FE_C0);
    r.v[1] = subc_cc(r.v[1], bo);
#pragma unroll
    for (int i = 2; i < 8; i++) r.v[i] = subc_cc(r.v[i], 0);
    uint32_t bo2 = subc(0, 0) & 1;
    r.v[0] = sub_cc(r.v[0], (0u - bo2) & FE_C0);
    r.v[1] = subc(r.v[1], bo2);
    return r;
}

PLUME_DEV fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }
PLUME_DEV fe fe_dbl(const fe& a) { return fe_add(a, a); }

// canonical representative in [0, p)
PLUME_DEV fe fe_norm(const fe& a) {
    // t = a + C; if that carries out of 2^256 then a >= p and a - p = t mod 2^256
    fe t;
    t.v[0] = add_cc(a.v[0], FE_C0);
    t.v[1] = addc_cc(a.v[1], 1);
#pragma unroll
    for (int i = 2; i < 8; i++) t.v[i] = addc_cc(a.v[i], 0);
    uint32_t co = addc(0, 0);
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = co ? t.v[i] : a.v[i];
    return r;
}

// a == 0 (mod p), a weakly reduced: a is 0 or p
PLUME_DEV bool fe_is_zero(const fe& a) {
    uint32_t o = a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7];
    uint32_t n = (a.v[0] ^ 0xFFFFFC2Fu) | (a.v[1] ^ 0xFFFFFFFEu) | ~(a.v[2] & a.v[3] & a.v[4] & a.v[5] & a.v[6] & a.v[7]);
    return (o == 0) | (n == 0);
}
PLUME_DEV bool fe_eq(const fe& a, const fe& b) { return fe_is_zero(fe_sub(a, b)); }
// parity of the canonical representative
PLUME_DEV uint32_t fe_is_odd(const fe& a) { return fe_norm(a).v[0] & 1; }
PLUME_DEV fe fe_cmov(const fe& a, const fe& b, bool pick_b) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = pick_b ? b.v[i] : a.v[i];
    return r;
}

// T[0..15] = a * b  (schoolbook; partial products a_i*b_j with i+j even go to E, odd to O, so
// that every 64-bit product lands on an aligned register pair and each row is one carry chain
// of four IMAD.WIDE.U32.X)
PLUME_DEV void fe_mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t E[16], O[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        E[2 * k] = mul_lo(a[2 * k], b[0]);
        E[2 * k + 1] = mul_hi(a[2 * k], b[0]);
        O[2 * k] = mul_lo(a[2 * k + 1], b[0]);
        O[2 * k + 1] = mul_hi(a[2 * k + 1], b[0]);
    }
#pragma unroll
    for (int j = 1; j < 8; j++) {
        const int se = (j & 1) ? j + 1 : j, ie = (j & 1) ? 1 : 0;  // E chain: start limb / first a index
        E[se] = mad_lo_cc(a[ie], b[j], E[se]);
        E[se + 1] = madc_hi_cc(a[ie], b[j], E[se + 1]);
#pragma unroll
        for (int k = 1; k < 4; k++) {
            E[se + 2 * k] = madc_lo_cc(a[ie + 2 * k], b[j], E[se + 2 * k]);
            E[se + 2 * k + 1] = madc_hi_cc(a[ie + 2 * k], b[j], E[se + 2 * k + 1]);
        }
        if (!(j & 1)) E[se + 8] = addc(E[se + 8], 0);  // top lane was not fresh: carry may leave it
        const int so = (j & 1) ? j - 1 : j, io = (j & 1) ? 0 : 1;  // O chain (limb index offset by one)
        O[so] = mad_lo_cc(a[io], b[j], O[so]);
        O[so + 1] = madc_hi_cc(a[io], b[j], O[so + 1]);
#pragma unroll
        for (int k = 1; k < 4; k++) {
            O[so + 2 * k] = madc_lo_cc(a[io + 2 * k], b[j], O[so + 2 * k]);
            O[so + 2 * k + 1] = madc_hi_cc(a[io + 2 * k], b[j], O[so + 2 * k + 1]);
        }
        if (j & 1) O[so + 8] = addc(O[so + 8], 0);
    }
    // T = E + (O << 32)
    T[0] = E[0];
    T[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int i = 2; i < 15; i++) T[i] = addc_cc(E[i], O[i - 1]);
    T[15] = addc(E[15], O[14]);
}

// T[0..15] = a^2: 28 off-diagonal products once, doubled, plus 8 diagonal squares = 36 IMAD.WIDE
PLUME_DEV void fe_sqr_wide(uint32_t* T, const uint32_t* a) {
    // off-diagonal sum S = sum_{i<j} a_i a_j 2^(32(i+j)); E: i+j even, O: i+j odd (offset one limb)
    uint32_t E[16], O[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
    // rows indexed by j = the larger index; row j multiplies a_j by a_i for i < j.
    // O rows (i+j odd): i has parity != j.  E rows (i+j even): i has parity == j, i < j.
    // O chains, start limb i+j-1 for the smallest i of the row, consecutive lanes:
    // j=1: i=0 -> lane 0.          j=2: i=1 -> lane 2.        j=3: i=0,2 -> lanes 2,4.
    // j=4: i=1,3 -> lanes 4,6.     j=5: i=0,2,4 -> lanes 4,6,8.   j=6: i=1,3,5 -> lanes 6,8,10.
    // j=7: i=0,2,4,6 -> lanes 6,8,10,12.
#pragma unroll
    for (int j = 1; j < 8; j++) {
        const int i0 = (j & 1) ? 0 : 1;
        const int cnt = (j + 1) / 2;  // number of i < j with parity != j
        const int s = i0 + j - 1;
        O[s] = mad_lo_cc(a[i0], a[j], O[s]);
        O[s + 1] = madc_hi_cc(a[i0], a[j], O[s + 1]);
#pragma unroll
        for (int k = 1; k < cnt; k++) {
            O[s + 2 * k] = madc_lo_cc(a[i0 + 2 * k], a[j], O[s + 2 * k]);
            O[s + 2 * k + 1] = madc_hi_cc(a[i0 + 2 * k], a[j], O[s + 2 * k + 1]);
        }
        // the row's top lane (limbs 2j-2, 2j-1) is untouched so far: product + carry-in < 2^64, no carry out
    }
    // E chains: j=2: i=0 -> lane 2. j=3: i=1 -> lane 4. j=4: i=0,2 -> lanes 4,6. j=5: i=1,3 -> 6,8.
    // j=6: i=0,2,4 -> 6,8,10. j=7: i=1,3,5 -> 8,10,12.
#pragma unroll
    for (int j = 2; j < 8; j++) {
        const int i0 = (j & 1) ? 1 : 0;
        const int cnt = j / 2;  // number of i < j with parity == j
        const int s = i0 + j;
        E[s] = mad_lo_cc(a[i0], a[j], E[s]);
        E[s + 1] = madc_hi_cc(a[i0], a[j], E[s + 1]);
#pragma unroll
        for (int k = 1; k < cnt; k++) {
            E[s + 2 * k] = madc_lo_cc(a[i0 + 2 * k], a[j], E[s + 2 * k]);
            E[s + 2 * k + 1] = madc_hi_cc(a[i0 + 2 * k], a[j], E[s + 2 * k + 1]);
        }
    }
    // S = E + (O << 32), then T = 2*S + diagonal
    uint32_t S[16];
    S[0] = 0;
    S[1] = O[0];
    S[2] = add_cc(E[2], O[1]);
#pragma unroll
    for (int i = 3; i < 15; i++) S[i] = addc_cc(E[i], O[i - 1]);
    S[15] = addc(E[15], O[14]);
    // double
    T[0] = 0;
    T[1] = add_cc(S[1], S[1]);
#pragma unroll
    for (int i = 2; i < 15; i++) T[i] = addc_cc(S[i], S[i]);
    T[15] = addc(S[15], S[15]);
    // diagonal a_i^2 at limb 2i: one chain of eight wide MADs
    T[0] = mad_lo_cc(a[0], a[0], T[0]);
    T[1] = madc_hi_cc(a[0], a[0], T[1]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        T[2 * i] = madc_lo_cc(a[i], a[i], T[2 * i]);
        T[2 * i + 1] = madc_hi_cc(a[i], a[i], T[2 * i + 1]);
    }
}

// r = T mod p (weakly reduced), T < 2^512
PLUME_DEV fe fe_reduce512(const uint32_t* T) {
    const uint32_t* h = T + 8;
    uint32_t A[9], Q[9];
    // even lanes: A = T_lo + sum_k h_{2k}*977*2^(64k)
    A[0] = mad_lo_cc(h[0], FE_C0, T[0]);
    A[1] = madc_hi_cc(h[0], FE_C0, T[1]);
#pragma unroll
    for (int k = 1; k < 4; k++) {
        A[2 * k] = madc_lo_cc(h[2 * k], FE_C0, T[2 * k]);
        A[2 * k + 1] = madc_hi_cc(h[2 * k], FE_C0, T[2 * k + 1]);
    }
    A[8] = addc(0, 0);
    // odd lanes (offset one limb): Q lane k = h_{2k+1}*977 + (h_{2k} + h_{2k+1}*2^32); the addend is
    // exactly the aligned register pair (h_2k, h_2k+1), i.e. the "T_hi << 32" term comes for free
    Q[0] = mad_lo_cc(h[1], FE_C0, h[0]);
    Q[1] = madc_hi_cc(h[1], FE_C0, h[1]);
#pragma unroll
    for (int k = 1; k < 4; k++) {
        Q[2 * k] = madc_lo_cc(h[2 * k + 1], FE_C0, h[2 * k]);
        Q[2 * k + 1] = madc_hi_cc(h[2 * k + 1], FE_C0, h[2 * k + 1]);
    }
    Q[8] = addc(0, 0);
    // R = A + (Q << 32): 10 limbs
    uint32_t R[10];
    R[0] = A[0];
    R[1] = add_cc(A[1], Q[0]);
#pragma unroll
    for (int i = 2; i < 9; i++) R[i] = addc_cc(A[i], Q[i - 1]);
    R[9] = addc(Q[8], 0);
    // second fold: t = R8 + R9*2^32 (t <= 2^32 + 978); add t*C = R8*977 + (R8 + R9*977)*2^32 + R9*2^64
    uint32_t v = R[8] + R[9] * FE_C0;  // fits: R9 = 1 implies R8 <= 978
    uint32_t u0 = mad_lo_cc(R[8], FE_C0, 0);
    uint32_t u1 = madc_hi_cc(R[8], FE_C0, v);
    uint32_t u2 = addc(R[9], 0);
    fe r;
    r.v[0] = add_cc(R[0], u0);
    r.v[1] = addc_cc(R[1], u1);
    r.v[2] = addc_cc(R[2], u2);
#pragma unroll
    for (int i = 3; i < 8; i++) r.v[i] = addc_cc(R[i], 0);
    uint32_t co = addc(0, 0);
    // third fold (rare): the wrapped value is < 2^66, adding C touches limbs 0..2
    r.v[0] = add_cc(r.v[0], (0u - co) & FE_C0);
    r.v[1] = addc_cc(r.v[1], co);
    r.v[2] = addc(r.v[2], 0);
    return r;
}

PLUME_DEV fe fe_mul(const fe& a, const fe& b) {
    uint32_t T[16];
    fe_mul_wide(T, a.v, b.v);
    return fe_reduce512(T);
}
PLUME_DEV fe fe_sqr(const fe& a) {
    uint32_t T[16];
    fe_sqr_wide(T, a.v);
    return fe_reduce512(T);
}
// a^(2^n)
PLUME_DEV fe fe_sqrn(fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe_sqr(a);
    return a;
}

// r = a * k for a small k (k <= 2^16): one chain + fold
PLUME_DEV fe fe_mul_small(const fe& a, uint32_t k) {
    uint32_t lo[9], hi[9];
    // a*k = sum a_i*k*2^32i : even i on aligned lanes, odd i offset by one
    lo[0] = mul_lo(a.v[0], k); lo[1] = mul_hi(a.v[0], k);
    lo[2] = mul_lo(a.v[2], k); lo[3] = mul_hi(a.v[2], k);
    lo[4] = mul_lo(a.v[4], k); lo[5] = mul_hi(a.v[4], k);
    lo[6] = mul_lo(a.v[6], k); lo[7] = mul_hi(a.v[6], k);
    hi[0] = mul_lo(a.v[1], k); hi[1] = mul_hi(a.v[1], k);
    hi[2] = mul_lo(a.v[3], k); hi[3] = mul_hi(a.v[3], k);
    hi[4] = mul_lo(a.v[5], k); hi[5] = mul_hi(a.v[5], k);
    hi[6] = mul_lo(a.v[7], k); hi[7] = mul_hi(a.v[7], k);
    uint32_t R[9];
    R[0] = lo[0];
    R[1] = add_cc(lo[1], hi[0]);
#pragma unroll
    for (int i = 2; i < 8; i++) R[i] = addc_cc(lo[i], hi[i - 1]);
    R[8] = addc(hi[7], 0);  // < 2^16 + 1
    // fold R8 * C
    uint32_t u0 = mad_lo_cc(R[8], FE_C0, 0);
    uint32_t u1 = madc_hi_cc(R[8], FE_C0, R[8]);
    fe r;
    r.v[0] = add_cc(R[0], u0);
    r.v[1] = addc_cc(R[1], u1);
#pragma unroll
    for (int i = 2; i < 8; i++) r.v[i] = addc_cc(R[i], 0);
    uint32_t co = addc(0, 0);
    r.v[0] = add_cc(r.v[0], (0u - co) & FE_C0);
    r.v[1] = addc_cc(r.v[1], co);
    r.v[2] = addc(r.v[2], 0);
    return r;
}

// x^(2^223 - 1) and the short blocks shared by the three fixed exponents below (chain layout as
// in the classic secp256k1 ladder: runs of ones of length 2,3,6,9,11,22,44,88,176,220,223)
struct fe_pow_blocks { fe x2, x22, x223; };
PLUME_DEV fe_pow_blocks fe_pow_common(const fe& x) {
    fe x2 = fe_mul(fe_sqr(x), x);
    fe x3 = fe_mul(fe_sqr(x2), x);
    fe x6 = fe_mul(fe_sqrn(x3, 3), x3);
    fe x9 = fe_mul(fe_sqrn(x6, 3), x3);
    fe x11 = fe_mul(fe_sqrn(x9, 2), x2);
    fe x22 = fe_mul(fe_sqrn(x11, 11), x11);
    fe x44 = fe_mul(fe_sqrn(x22, 22), x22);
    fe x88 = fe_mul(fe_sqrn(x44, 44), x44);
    fe x176 = fe_mul(fe_sqrn(x88, 88), x88);
    fe x220 = fe_mul(fe_sqrn(x176, 44), x44);
    fe x223 = fe_mul(fe_sqrn(x220, 3), x3);
    fe_pow_blocks b;
    b.x2 = x2; b.x22 = x22; b.x223 = x223;
    return b;
}
// x^(p-2): exponent bits = [223 ones][0][22 ones][0000][101101]
PLUME_DEV fe fe_inv(const fe& x) {
    fe_pow_blocks b = fe_pow_common(x);
    fe t = fe_mul(fe_sqrn(b.x223, 23), b.x22);
    t = fe_mul(fe_sqrn(t, 5), x);
    t = fe_mul(fe_sqrn(t, 3), b.x2);
    t = fe_mul(fe_sqrn(t, 2), x);
    return t;
}
// x^((p-3)/4): exponent bits = [223 ones][0][22 ones][0000][1011]   (RFC 9380 F.2.1.2 c1)
PLUME_DEV fe fe_pow_pm3d4(const fe& x) {
    fe_pow_blocks b = fe_pow_common(x);
    fe t = fe_mul(fe_sqrn(b.x223, 23), b.x22);
    t = fe_mul(fe_sqrn(t, 5), x);
    t = fe_mul(fe_sqrn(t, 3), b.x2);
    return t;
}
// x^((p+1)/4) = a square root of x when x is a QR: bits = [223 ones][0][22 ones][0000][1100]
PLUME_DEV fe fe_sqrt_cand(const fe& x) {
    fe_pow_blocks b = fe_pow_common(x);
    fe t = fe_mul(fe_sqrn(b.x223, 23), b.x22);
    t = fe_mul(fe_sqrn(t, 6), b.x2);
    t = fe_sqrn(t, 2);
    return t;
}

// 32 big-endian bytes <-> limbs.  `be` points at 8 big-endian 32-bit words (byte order of the wire).
PLUME_DEV fe fe_from_be_words(const uint32_t* w) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = bswap32(w[7 - i]);
    return r;
}
PLUME_DEV void fe_to_be_words(uint32_t* w, const fe& a) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[7 - i] = bswap32(a.v[i]);
}

// fe.cuh -- arithmetic in Fp, p = 2^256 - 2^32 - 977 (secp256k1 base field).
//
// Representation: 8 x 32-bit little-endian limbs held in registers, value "weakly reduced":
// any representative in [0, 2^256).  fe_norm() gives the canonical one in [0, p).
// Reduction is the pseudo-Mersenne fold 2^256 == C (mod p), C = 2^32 + 977 -- no Montgomery
// form (SURVEY.md section 7 allows either; the fold costs 10 extra limb-products per multiplication
// instead of 64).
//
// Device code: every carry chain is ONE non-volatile inline-asm statement (a row of four
// IMAD.WIDE.U32.X, an 8-limb add, ...).  That matters: PTX has a single carry flag, and with one
// asm statement per instruction ptxas kept all chains of a multiplication in program order on one
// predicate (profiles/r01_sign_varbase_v1.md: 4.05 "wait" stall cycles per issued instruction).
// Whole-chain statements are independent values to the compiler, so ptxas renames the carry to
// different predicates and interleaves the even-column and odd-column chains of a product, and
// neighbouring additions, on the two integer pipes.
//
// Host-sim build: portable C versions of the same functions (the PTX itself is exercised on the
// GPU by tests/test_gpu_field.py through plume_debug_fe_op).
//
// This replaces what the reference obtains from k256::FieldElement (external crate k256 ~0.13,
// rust-k256/Cargo.toml:18); constants p / b per rust-arkworks/src/secp256k1/fields/fq.rs:12.
#pragma once
#include "ptx.cuh"

#ifdef PLUME_INLINE_MUL
#define PLUME_MULFN PLUME_DEV
#else
#define PLUME_MULFN PLUME_DEV_NOINLINE
#endif

struct fe { uint32_t v[8]; };

#define FE_C0 977u  // C = 2^32 + FE_C0

PLUME_DEV fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
PLUME_DEV fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }
PLUME_DEV fe fe_set_u32(uint32_t x) { fe r = fe_zero(); r.v[0] = x; return r; }

#ifdef PLUME_HOSTSIM
// ------------------------------------------------------------------------------------------------
// portable versions (host-sim only)
// ------------------------------------------------------------------------------------------------
// r = t mod p, weakly reduced (t: 16 limbs): lo + hi * C until nothing is left above limb 7
static inline void fe_host_fold(uint32_t* r, const uint32_t* t16) {
    uint32_t cur[20] = {0};
    for (int i = 0; i < 16; i++) cur[i] = t16[i];
    for (;;) {
        bool any = false;
        for (int i = 8; i < 20; i++) any |= cur[i] != 0;
        if (!any) break;
        uint32_t nxt[20] = {0};
        for (int i = 0; i < 8; i++) nxt[i] = cur[i];
        uint64_t c = 0;
        for (int k = 0; k < 11; k++) {  // += hi * 977
            c += (uint64_t)cur[8 + k] * FE_C0 + nxt[k];
            nxt[k] = (uint32_t)c;
            c >>= 32;
        }
        for (int k = 11; k < 20 && c; k++) { c += nxt[k]; nxt[k] = (uint32_t)c; c >>= 32; }
        c = 0;
        for (int k = 0; k < 11; k++) {  // += hi << 32
            c += (uint64_t)cur[8 + k] + nxt[k + 1];
            nxt[k + 1] = (uint32_t)c;
            c >>= 32;
        }
        for (int k = 12; k < 20 && c; k++) { c += nxt[k]; nxt[k] = (uint32_t)c; c >>= 32; }
        for (int i = 0; i < 20; i++) cur[i] = nxt[i];
    }
    for (int i = 0; i < 8; i++) r[i] = cur[i];
}
PLUME_DEV fe fe_reduce512(const uint32_t* T) { fe r; fe_host_fold(r.v, T); return r; }
PLUME_DEV fe fe_add(const fe& a, const fe& b) {
    uint32_t t[16] = {0};
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    t[8] = (uint32_t)c;
    return fe_reduce512(t);
}
PLUME_DEV fe fe_sub(const fe& a, const fe& b) {
    // on borrow the wrapped value is a - b + 2^256 == a - b + C: subtract C (at most twice)
    uint32_t t[8];
    uint64_t bo = 0;
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - bo; t[i] = (uint32_t)d; bo = (d >> 32) & 1; }
    for (int round = 0; round < 2 && bo; round++) {
        const uint64_t sub[8] = {FE_C0, 1, 0, 0, 0, 0, 0, 0};
        uint64_t b2 = 0;
        for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)t[i] - sub[i] - b2; t[i] = (uint32_t)d; b2 = (d >> 32) & 1; }
        bo = b2;
    }
    fe r;
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return r;
}
PLUME_DEV fe fe_mul(const fe& a, const fe& b) {
    uint32_t t[16] = {0};
    for (int j = 0; j < 8; j++) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] * b.v[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        t[j + 8] = (uint32_t)c;
    }
    return fe_reduce512(t);
}
PLUME_DEV fe fe_sqr(const fe& a) { return fe_mul(a, a); }
PLUME_DEV fe fe_mul_inl(const fe& a, const fe& b) { return fe_mul(a, b); }
PLUME_DEV fe fe_sqr_inl(const fe& a) { return fe_mul(a, a); }

#else
// ------------------------------------------------------------------------------------------------
// device versions
// ------------------------------------------------------------------------------------------------

// r = a + b.  The wrap-around correction (+C when the 256-bit sum carried out) only ripples past
// limb 1 when limb 1 overflows as well (probability ~2^-31 for random data): that case takes a
// branch to an exact slow path instead of paying a 6-limb carry chain on every addition.
PLUME_DEV fe fe_add(const fe& a, const fe& b) {
    fe r;
    uint32_t c1;
    asm("{\n\t.reg .u32 co;\n\t"
        "add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 co, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, co, 977, %0;\n\taddc.cc.u32 %1, %1, co;\n\taddc.u32 %8, 0, 0;\n\t}"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(c1)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    if (c1) {
        uint32_t co2;
        asm("add.cc.u32 %0, %0, 1;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.u32 %6, 0, 0;"
            : "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(co2));
        // second wrap (operands were non-canonical): the wrapped value is < C, touches limbs 0,1 only
        asm("mad.lo.cc.u32 %0, %2, 977, %0;\n\taddc.u32 %1, %1, %2;" : "+r"(r.v[0]), "+r"(r.v[1]) : "r"(co2));
    }
    return r;
}

// r = a - b  (same structure with borrows)
PLUME_DEV fe fe_sub(const fe& a, const fe& b) {
    fe r;
    uint32_t b1;
    asm("{\n\t.reg .u32 bo, t, u;\n\t"
        "sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\tsubc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\tsubc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 bo, 0, 0;\n\t"          // 0 or 0xffffffff
        "and.b32 t, bo, 977;\n\tand.b32 u, bo, 1;\n\t"
        "sub.cc.u32 %0, %0, t;\n\tsubc.cc.u32 %1, %1, u;\n\tsubc.u32 %8, 0, 0;\n\t}"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(b1)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    if (b1) {
        uint32_t bo2;
        asm("sub.cc.u32 %0, %0, 1;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.cc.u32 %2, %2, 0;\n\tsubc.cc.u32 %3, %3, 0;\n\t"
            "subc.cc.u32 %4, %4, 0;\n\tsubc.cc.u32 %5, %5, 0;\n\tsubc.u32 %6, 0, 0;"
            : "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(bo2));
        asm("{\n\t.reg .u32 t, u;\n\tand.b32 t, %2, 977;\n\tand.b32 u, %2, 1;\n\tsub.cc.u32 %0, %0, t;\n\tsubc.u32 %1, %1, u;\n\t}"
            : "+r"(r.v[0]), "+r"(r.v[1]) : "r"(bo2));
    }
    return r;
}

// Rows of the schoolbook product: acc += (a0, a1, ..) * b as chained 64-bit lanes.  Words a row
// touches for the first time are pure outputs, so no accumulator ever needs a zeroing move:
//   fe_row4t : lanes 0..2 accumulate, lane 3 = (acc[6] += .., acc[7] fresh); no carry out of a fresh word
//   fe_row4c : lanes 0..3 accumulate, the carry out is written to the fresh word acc[8]
//   fe_rowNf : N lanes, the first N-1 accumulate, the last lane is entirely fresh (squaring)
//
// A lane can be computed two ways.  "Hard": one IMAD.WIDE.U32.X (multiply + 64-bit add + carry in/out).
// "Soft": a plain IMAD.WIDE.U32 product followed by two IADD3.X inside the same carry chain.  Both IMAD.WIDE
// forms cost the same ~4.2 issue cycles per warp on the B200 and the additions do not overlap with them
// (profiles/r01_imad_rates.md), so the soft form only adds instructions: measured 16 % SLOWER (k_sign_varbase
// 20.3 ms vs 17.4 ms per 2^19 items).  The soft rows are kept behind -DPLUME_SOFT_LANES=1 as the record of that
// experiment; the default is all-hard.
#ifndef PLUME_SOFT_LANES
#define PLUME_SOFT_LANES 0
#endif
PLUME_DEV void fe_prod(uint32_t* acc, uint32_t a, uint32_t b) {  // plain 64-bit product (IMAD.WIDE.U32, no carry)
    uint64_t p = (uint64_t)a * b;
    acc[0] = (uint32_t)p;
    acc[1] = (uint32_t)(p >> 32);
}
#if PLUME_SOFT_LANES
PLUME_DEV void fe_row4t(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    uint32_t p1[2], p3[2];
    fe_prod(p1, a1, b);
    fe_prod(p3, a3, b);
    asm("mad.lo.cc.u32 %0, %8, %14, %0;\n\tmadc.hi.cc.u32 %1, %8, %14, %1;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
        "madc.lo.cc.u32 %4, %9, %14, %4;\n\tmadc.hi.cc.u32 %5, %9, %14, %5;\n\t"
        "addc.cc.u32 %6, %6, %12;\n\taddc.u32 %7, %13, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "=r"(acc[7])
        : "r"(a0), "r"(a2), "r"(p1[0]), "r"(p1[1]), "r"(p3[0]), "r"(p3[1]), "r"(b));
}
PLUME_DEV void fe_row4c(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    uint32_t p1[2], p3[2];
    fe_prod(p1, a1, b);
    fe_prod(p3, a3, b);
    asm("mad.lo.cc.u32 %0, %9, %15, %0;\n\tmadc.hi.cc.u32 %1, %9, %15, %1;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\taddc.cc.u32 %3, %3, %12;\n\t"
        "madc.lo.cc.u32 %4, %10, %15, %4;\n\tmadc.hi.cc.u32 %5, %10, %15, %5;\n\t"
        "addc.cc.u32 %6, %6, %13;\n\taddc.cc.u32 %7, %7, %14;\n\taddc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(acc[8])
        : "r"(a0), "r"(a2), "r"(p1[0]), "r"(p1[1]), "r"(p3[0]), "r"(p3[1]), "r"(b));
}
PLUME_DEV void fe_row2f(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b) {
    uint32_t p1[2];
    fe_prod(p1, a1, b);
    asm("mad.lo.cc.u32 %0, %4, %7, %0;\n\tmadc.hi.cc.u32 %1, %4, %7, %1;\n\taddc.cc.u32 %2, %5, 0;\n\taddc.u32 %3, %6, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]) : "r"(a0), "r"(p1[0]), "r"(p1[1]), "r"(b));
}
PLUME_DEV void fe_row3f(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
    uint32_t p1[2], p2[2];
    fe_prod(p1, a1, b);
    fe_prod(p2, a2, b);
    asm("mad.lo.cc.u32 %0, %6, %11, %0;\n\tmadc.hi.cc.u32 %1, %6, %11, %1;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\t"
        "addc.cc.u32 %4, %9, 0;\n\taddc.u32 %5, %10, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(acc[4]), "=r"(acc[5])
        : "r"(a0), "r"(p1[0]), "r"(p1[1]), "r"(p2[0]), "r"(p2[1]), "r"(b));
}
PLUME_DEV void fe_row4f(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    uint32_t p1[2], p3[2];
    fe_prod(p1, a1, b);
    fe_prod(p3, a3, b);
    asm("mad.lo.cc.u32 %0, %8, %14, %0;\n\tmadc.hi.cc.u32 %1, %8, %14, %1;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
        "madc.lo.cc.u32 %4, %9, %14, %4;\n\tmadc.hi.cc.u32 %5, %9, %14, %5;\n\t"
        "addc.cc.u32 %6, %12, 0;\n\taddc.u32 %7, %13, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
        : "r"(a0), "r"(a2), "r"(p1[0]), "r"(p1[1]), "r"(p3[0]), "r"(p3[1]), "r"(b));
}
#else
PLUME_DEV void fe_row4t(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "=r"(acc[7])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
PLUME_DEV void fe_row4c(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(acc[8])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
PLUME_DEV void fe_row2f(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %6, 0;\n\tmadc.hi.u32 %3, %5, %6, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]) : "r"(a0), "r"(a1), "r"(b));
}
PLUME_DEV void fe_row3f(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %6, %9, %0;\n\tmadc.hi.cc.u32 %1, %6, %9, %1;\n\tmadc.lo.cc.u32 %2, %7, %9, %2;\n\tmadc.hi.cc.u32 %3, %7, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %9, 0;\n\tmadc.hi.u32 %5, %8, %9, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]) : "r"(a0), "r"(a1), "r"(a2), "r"(b));
}
PLUME_DEV void fe_row4f(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, 0;\n\tmadc.hi.u32 %7, %11, %12, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
#endif

// t[0..7] = x[0..7] + y[0..7] + cin, returns the carry out (cin, cout in {0,1}); long additions are
// stitched from these so that every statement stays under the 30-operand limit of inline asm
PLUME_DEV uint32_t fe_add8_c(uint32_t* t, const uint32_t* x, const uint32_t* y, uint32_t cin) {
    uint32_t cout;
    asm("add.cc.u32 %0, %25, 0xffffffff;\n\t"   // regenerate the carry flag from cin
        "addc.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(cout)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(cin));
    return cout;
}
// same without a carry in
PLUME_DEV uint32_t fe_add8(uint32_t* t, const uint32_t* x, const uint32_t* y) {
    uint32_t cout;
    asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(cout)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
    return cout;
}

// T[0..15] = a * b  (schoolbook; partial products a_i*b_j with i+j even go to E, odd to O, so
// that every 64-bit product lands on an aligned register pair and each row is one carry chain
// of four IMAD.WIDE.U32.X; the E rows and the O rows are two independent dependency streams)
PLUME_DEV void fe_mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t E[16], O[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        fe_prod(E + 2 * k, a[2 * k], b[0]);
        fe_prod(O + 2 * k, a[2 * k + 1], b[0]);
    }
    E[8] = 0;  // the only word read before a row has written it (row j = 1 accumulates into limbs 2..8)
#pragma unroll
    for (int j = 1; j < 8; j++) {
        const int se = (j & 1) ? j + 1 : j, ie = (j & 1) ? 1 : 0;  // E chain: start limb / first a index
        const int so = (j & 1) ? j - 1 : j, io = (j & 1) ? 0 : 1;  // O chain (limb index offset by one)
        // A row's top lane holds at most the previous row's carry word, so: odd j -> the E row cannot
        // carry out and its top word is fresh, the O row carries into a fresh word; even j the other way.
        if (j & 1) {
            fe_row4t(E + se, a[ie], a[ie + 2], a[ie + 4], a[ie + 6], b[j]);
            fe_row4c(O + so, a[io], a[io + 2], a[io + 4], a[io + 6], b[j]);
        } else {
            fe_row4c(E + se, a[ie], a[ie + 2], a[ie + 4], a[ie + 6], b[j]);
            fe_row4t(O + so, a[io], a[io + 2], a[io + 4], a[io + 6], b[j]);
        }
    }
    // T = E + (O << 32); E has 16 limbs, O has 15 (limbs 0..14)
    T[0] = E[0];
    uint32_t x[8], y[8], t2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = E[1 + i]; y[i] = O[i]; }
    uint32_t c = fe_add8(T + 1, x, y);
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = (i < 7) ? E[9 + i] : 0; y[i] = (i < 7) ? O[8 + i] : 0; }
    fe_add8_c(t2, x, y, c);
#pragma unroll
    for (int i = 0; i < 7; i++) T[9 + i] = t2[i];
}

// T[0..15] = a^2: 28 off-diagonal products once, doubled, plus 8 diagonal squares = 36 IMAD.WIDE
PLUME_DEV void fe_sqr_wide(uint32_t* T, const uint32_t* a) {
    // off-diagonal sum S = sum_{i<j} a_i a_j 2^(32(i+j)); E: i+j even, O: i+j odd (offset one limb).
    // Row j multiplies a_j by the a_i (i < j) of the right parity; the row's top lane is always
    // untouched so far, so no row carries out.
    uint32_t E[16], O[16];
    // O rows: j=1: i=0 @0 | j=2: i=1 @2 | j=3: i=0,2 @2 | j=4: i=1,3 @4 | j=5: i=0,2,4 @4 | j=6: i=1,3,5 @6 | j=7: i=0,2,4,6 @6
    fe_prod(O + 0, a[0], a[1]);
    fe_prod(O + 2, a[1], a[2]);
    fe_row2f(O + 2, a[0], a[2], a[3]);
    fe_row2f(O + 4, a[1], a[3], a[4]);
    fe_row3f(O + 4, a[0], a[2], a[4], a[5]);
    fe_row3f(O + 6, a[1], a[3], a[5], a[6]);
    fe_row4f(O + 6, a[0], a[2], a[4], a[6], a[7]);
    O[14] = 0;
    // E rows: j=2: i=0 @2 | j=3: i=1 @4 | j=4: i=0,2 @4 | j=5: i=1,3 @6 | j=6: i=0,2,4 @6 | j=7: i=1,3,5 @8
    fe_prod(E + 2, a[0], a[2]);
    fe_prod(E + 4, a[1], a[3]);
    fe_row2f(E + 4, a[0], a[2], a[4]);
    fe_row2f(E + 6, a[1], a[3], a[5]);
    fe_row3f(E + 6, a[0], a[2], a[4], a[6]);
    fe_row3f(E + 8, a[1], a[3], a[5], a[7]);
    E[14] = 0; E[15] = 0;
    // S = E + (O << 32): S[0] = 0, S[1] = O[0], S[2..15] = E[2..15] + O[1..14]
    uint32_t S[16], x[8], y[8], s2[8];
    S[0] = 0;
    S[1] = O[0];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = E[2 + i]; y[i] = O[1 + i]; }
    uint32_t c = fe_add8(S + 2, x, y);
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = (i < 6) ? E[10 + i] : 0; y[i] = (i < 6) ? O[9 + i] : 0; }
    fe_add8_c(s2, x, y, c);
#pragma unroll
    for (int i = 0; i < 6; i++) S[10 + i] = s2[i];
    // 2*S by funnel shifts (no carry chain), then the diagonal a_i^2 at limb 2i as two 4-lane rows
    uint32_t D[16];
    D[0] = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) D[i] = __funnelshift_l(S[i - 1], S[i], 1);
#if PLUME_SOFT_LANES
    {
        uint32_t q1[2], q3[2], q5[2], q7[2];
        fe_prod(q1, a[1], a[1]); fe_prod(q3, a[3], a[3]); fe_prod(q5, a[5], a[5]); fe_prod(q7, a[7], a[7]);
        asm("mad.lo.cc.u32 %0, %9, %9, %0;\n\tmadc.hi.cc.u32 %1, %9, %9, %1;\n\t"
            "addc.cc.u32 %2, %2, %11;\n\taddc.cc.u32 %3, %3, %12;\n\t"
            "madc.lo.cc.u32 %4, %10, %10, %4;\n\tmadc.hi.cc.u32 %5, %10, %10, %5;\n\t"
            "addc.cc.u32 %6, %6, %13;\n\taddc.cc.u32 %7, %7, %14;\n\taddc.u32 %8, 0, 0;"
            : "+r"(D[0]), "+r"(D[1]), "+r"(D[2]), "+r"(D[3]), "+r"(D[4]), "+r"(D[5]), "+r"(D[6]), "+r"(D[7]), "=r"(c)
            : "r"(a[0]), "r"(a[2]), "r"(q1[0]), "r"(q1[1]), "r"(q3[0]), "r"(q3[1]));
        asm("add.cc.u32 %8, %8, 0xffffffff;\n\t"
            "madc.lo.cc.u32 %0, %9, %9, %0;\n\tmadc.hi.cc.u32 %1, %9, %9, %1;\n\t"
            "addc.cc.u32 %2, %2, %11;\n\taddc.cc.u32 %3, %3, %12;\n\t"
            "madc.lo.cc.u32 %4, %10, %10, %4;\n\tmadc.hi.cc.u32 %5, %10, %10, %5;\n\t"
            "addc.cc.u32 %6, %6, %13;\n\taddc.u32 %7, %7, %14;"
            : "+r"(D[8]), "+r"(D[9]), "+r"(D[10]), "+r"(D[11]), "+r"(D[12]), "+r"(D[13]), "+r"(D[14]), "+r"(D[15]), "+r"(c)
            : "r"(a[4]), "r"(a[6]), "r"(q5[0]), "r"(q5[1]), "r"(q7[0]), "r"(q7[1]));
    }
#else
    asm("mad.lo.cc.u32 %0, %9, %9, %0;\n\tmadc.hi.cc.u32 %1, %9, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %10, %2;\n\tmadc.hi.cc.u32 %3, %10, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %11, %4;\n\tmadc.hi.cc.u32 %5, %11, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %12, %6;\n\tmadc.hi.cc.u32 %7, %12, %12, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(D[0]), "+r"(D[1]), "+r"(D[2]), "+r"(D[3]), "+r"(D[4]), "+r"(D[5]), "+r"(D[6]), "+r"(D[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]));
    asm("add.cc.u32 %8, %8, 0xffffffff;\n\t"
        "madc.lo.cc.u32 %0, %9, %9, %0;\n\tmadc.hi.cc.u32 %1, %9, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %10, %2;\n\tmadc.hi.cc.u32 %3, %10, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %11, %4;\n\tmadc.hi.cc.u32 %5, %11, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %12, %6;\n\tmadc.hi.u32 %7, %12, %12, %7;"
        : "+r"(D[8]), "+r"(D[9]), "+r"(D[10]), "+r"(D[11]), "+r"(D[12]), "+r"(D[13]), "+r"(D[14]), "+r"(D[15]), "+r"(c)
        : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#endif
#pragma unroll
    for (int i = 0; i < 16; i++) T[i] = D[i];
}

// r = T mod p (weakly reduced), T < 2^512
PLUME_DEV fe fe_reduce512(const uint32_t* T) {
    const uint32_t* h = T + 8;
    uint32_t A[9], Q[9];
#if PLUME_SOFT_LANES
    {
        // even lanes: A = T_lo + sum_k h_{2k}*977*2^(64k); lanes 1 and 3 as plain products added in
        uint32_t e1[2], e3[2], o1[2], o3[2];
        fe_prod(e1, h[2], FE_C0); fe_prod(e3, h[6], FE_C0);
        asm("mad.lo.cc.u32 %0, %9, 977, %15;\n\tmadc.hi.cc.u32 %1, %9, 977, %16;\n\t"
            "addc.cc.u32 %2, %17, %11;\n\taddc.cc.u32 %3, %18, %12;\n\t"
            "madc.lo.cc.u32 %4, %10, 977, %19;\n\tmadc.hi.cc.u32 %5, %10, 977, %20;\n\t"
            "addc.cc.u32 %6, %21, %13;\n\taddc.cc.u32 %7, %22, %14;\n\taddc.u32 %8, 0, 0;"
            : "=&r"(A[0]), "=&r"(A[1]), "=&r"(A[2]), "=&r"(A[3]), "=&r"(A[4]), "=&r"(A[5]), "=&r"(A[6]), "=&r"(A[7]), "=&r"(A[8])
            : "r"(h[0]), "r"(h[4]), "r"(e1[0]), "r"(e1[1]), "r"(e3[0]), "r"(e3[1]),
              "r"(T[0]), "r"(T[1]), "r"(T[2]), "r"(T[3]), "r"(T[4]), "r"(T[5]), "r"(T[6]), "r"(T[7]));
        // odd lanes (offset one limb): Q lane k = h_{2k+1}*977 + (h_{2k} + h_{2k+1}*2^32)
        fe_prod(o1, h[3], FE_C0); fe_prod(o3, h[7], FE_C0);
        asm("mad.lo.cc.u32 %0, %10, 977, %9;\n\tmadc.hi.cc.u32 %1, %10, 977, %10;\n\t"
            "addc.cc.u32 %2, %11, %17;\n\taddc.cc.u32 %3, %12, %18;\n\t"
            "madc.lo.cc.u32 %4, %14, 977, %13;\n\tmadc.hi.cc.u32 %5, %14, 977, %14;\n\t"
            "addc.cc.u32 %6, %15, %19;\n\taddc.cc.u32 %7, %16, %20;\n\taddc.u32 %8, 0, 0;"
            : "=&r"(Q[0]), "=&r"(Q[1]), "=&r"(Q[2]), "=&r"(Q[3]), "=&r"(Q[4]), "=&r"(Q[5]), "=&r"(Q[6]), "=&r"(Q[7]), "=&r"(Q[8])
            : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]),
              "r"(o1[0]), "r"(o1[1]), "r"(o3[0]), "r"(o3[1]));
    }
#else
    // even lanes: A = T_lo + sum_k h_{2k}*977*2^(64k)
    asm("mad.lo.cc.u32 %0, %9, 977, %13;\n\tmadc.hi.cc.u32 %1, %9, 977, %14;\n\t"
        "madc.lo.cc.u32 %2, %10, 977, %15;\n\tmadc.hi.cc.u32 %3, %10, 977, %16;\n\t"
        "madc.lo.cc.u32 %4, %11, 977, %17;\n\tmadc.hi.cc.u32 %5, %11, 977, %18;\n\t"
        "madc.lo.cc.u32 %6, %12, 977, %19;\n\tmadc.hi.cc.u32 %7, %12, 977, %20;\n\taddc.u32 %8, 0, 0;"
        : "=&r"(A[0]), "=&r"(A[1]), "=&r"(A[2]), "=&r"(A[3]), "=&r"(A[4]), "=&r"(A[5]), "=&r"(A[6]), "=&r"(A[7]), "=&r"(A[8])
        : "r"(h[0]), "r"(h[2]), "r"(h[4]), "r"(h[6]), "r"(T[0]), "r"(T[1]), "r"(T[2]), "r"(T[3]), "r"(T[4]), "r"(T[5]), "r"(T[6]), "r"(T[7]));
    // odd lanes (offset one limb): Q lane k = h_{2k+1}*977 + (h_{2k} + h_{2k+1}*2^32); the addend is
    // exactly the aligned register pair (h_2k, h_2k+1), i.e. the "T_hi << 32" term comes for free
    asm("mad.lo.cc.u32 %0, %10, 977, %9;\n\tmadc.hi.cc.u32 %1, %10, 977, %10;\n\t"
        "madc.lo.cc.u32 %2, %12, 977, %11;\n\tmadc.hi.cc.u32 %3, %12, 977, %12;\n\t"
        "madc.lo.cc.u32 %4, %14, 977, %13;\n\tmadc.hi.cc.u32 %5, %14, 977, %14;\n\t"
        "madc.lo.cc.u32 %6, %16, 977, %15;\n\tmadc.hi.cc.u32 %7, %16, 977, %16;\n\taddc.u32 %8, 0, 0;"
        : "=&r"(Q[0]), "=&r"(Q[1]), "=&r"(Q[2]), "=&r"(Q[3]), "=&r"(Q[4]), "=&r"(Q[5]), "=&r"(Q[6]), "=&r"(Q[7]), "=&r"(Q[8])
        : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]));
#endif
    // R = A + (Q << 32): 10 limbs
    uint32_t R[10];
    R[0] = A[0];
    uint32_t c = fe_add8(R + 1, A + 1, Q);
    R[9] = Q[8] + c;
    // second fold: t = R8 + R9*2^32 (t <= 2^32 + 978); add t*C = R8*977 + (R8 + R9*977)*2^32 + R9*2^64
    uint32_t v = R[8] + R[9] * FE_C0;  // fits: R9 = 1 implies R8 <= 978
    fe r;
    uint32_t c2;
    asm("{\n\t.reg .u32 u0, u1, u2;\n\t"
        "mad.lo.cc.u32 u0, %4, 977, 0;\n\tmadc.hi.cc.u32 u1, %4, 977, %5;\n\taddc.u32 u2, %6, 0;\n\t"
        "add.cc.u32 %0, %7, u0;\n\taddc.cc.u32 %1, %8, u1;\n\taddc.cc.u32 %2, %9, u2;\n\taddc.u32 %3, 0, 0;\n\t}"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(c2)
        : "r"(R[8]), "r"(v), "r"(R[9]), "r"(R[0]), "r"(R[1]), "r"(R[2]));
#pragma unroll
    for (int i = 3; i < 8; i++) r.v[i] = R[i];
    if (c2) {  // rare: the carry leaves limb 2
        uint32_t co;
        asm("add.cc.u32 %0, %0, 1;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\taddc.u32 %5, 0, 0;"
            : "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(co));
        // third fold: the wrapped value is < 2^66, adding C touches limbs 0..2
        asm("mad.lo.cc.u32 %0, %3, 977, %0;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 %2, %2, 0;" : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]) : "r"(co));
    }
    return r;
}

// The two big bodies are real functions (not inlined): every point formula calls them 7-11 times and
// the instruction cache, not the register file, is what the inlined version ran out of
// (profiles/r01_sign_varbase_v0.md: 2.6 "no instruction" stall cycles per issued instruction).
PLUME_DEV fe fe_mul_inl(const fe& a, const fe& b) {
    uint32_t T[16];
    fe_mul_wide(T, a.v, b.v);
    return fe_reduce512(T);
}
PLUME_DEV fe fe_sqr_inl(const fe& a) {
    uint32_t T[16];
    fe_sqr_wide(T, a.v);
    return fe_reduce512(T);
}
PLUME_MULFN fe fe_mul(fe a, fe b) { return fe_mul_inl(a, b); }
PLUME_MULFN fe fe_sqr(fe a) { return fe_sqr_inl(a); }
// a^(2^n) as ONE call: the exponentiation ladders (inversion, square roots) run 250+ squarings in a row, and
// paying the call marshalling per squaring is a fifth of their cost
PLUME_MULFN fe fe_sqrn(fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        uint32_t T[16];
        fe_sqr_wide(T, a.v);
        a = fe_reduce512(T);
    }
    return a;
}
#endif  // device versions

PLUME_DEV fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }
PLUME_DEV fe fe_dbl(const fe& a) { return fe_add(a, a); }

// r = a * 2^K (K = 1..3) as ONE pass: eight funnel shifts (no carry chain) and the K bits shifted out of the top folded
// back as t * C (t < 8) into limbs 0 and 1; like fe_add, the carry leaves limb 1 only with probability ~2^-31 and then
// takes a branch to the exact ripple.  The doubling formula's 8*C was three fe_dbl (51 instructions) before this.
template <int K>
PLUME_DEV fe fe_shl(const fe& a) {
    fe r;
    const uint32_t t = a.v[7] >> (32 - K);
#ifdef PLUME_HOSTSIM
    uint32_t T[16];
    for (int i = 0; i < 16; i++) T[i] = 0;
    T[0] = a.v[0] << K;
    for (int i = 1; i < 8; i++) T[i] = (a.v[i] << K) | (a.v[i - 1] >> (32 - K));
    T[8] = t;
    r = fe_reduce512(T);
#else
    r.v[0] = a.v[0] << K;
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], K);
    uint32_t c1;
    asm("mad.lo.cc.u32 %0, %3, 977, %0;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 %2, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "=r"(c1) : "r"(t));
    if (c1) {
        uint32_t co2;
        asm("add.cc.u32 %0, %0, 1;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.u32 %6, 0, 0;"
            : "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(co2));
        // second wrap: the wrapped value is tiny (limbs 2..7 are zero), adding C touches limbs 0, 1 only
        asm("mad.lo.cc.u32 %0, %2, 977, %0;\n\taddc.u32 %1, %1, %2;" : "+r"(r.v[0]), "+r"(r.v[1]) : "r"(co2));
    }
#endif
    return r;
}

// canonical representative in [0, p)   (plain 64-bit arithmetic: not on the hot path)
PLUME_DEV fe fe_norm(const fe& a) {
    // t = a + C; if that carries out of 2^256 then a >= p and a - p = t mod 2^256
    fe t;
    uint64_t c = (uint64_t)a.v[0] + FE_C0;
    t.v[0] = (uint32_t)c; c >>= 32;
    c += (uint64_t)a.v[1] + 1;
    t.v[1] = (uint32_t)c; c >>= 32;
#pragma unroll
    for (int i = 2; i < 8; i++) { c += a.v[i]; t.v[i] = (uint32_t)c; c >>= 32; }
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? t.v[i] : a.v[i];
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

#ifdef PLUME_HOSTSIM
// a^(2^n)
PLUME_DEV fe fe_sqrn(fe a, int n) {
    for (int i = 0; i < n; i++) a = fe_sqr(a);
    return a;
}
#endif

// r = a * k for a small k   (plain 64-bit arithmetic: a handful of uses in the SSWU map)
PLUME_DEV fe fe_mul_small(const fe& a, uint32_t k) {
    uint32_t T[16];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] * k; T[i] = (uint32_t)c; c >>= 32; }
    T[8] = (uint32_t)c;
#pragma unroll
    for (int i = 9; i < 16; i++) T[i] = 0;
    return fe_reduce512(T);
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

// fe29.h -- experiment: Fp element as 9 limbs of 29 bits, column sums with plain IMAD.WIDE.U32 (no carry flags).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define F29 __host__ __device__ __forceinline__
#else
#define F29 static inline
#endif
struct fe9 { uint32_t v[9]; };
#define M29 0x1FFFFFFFu
#define R0_29 31264u            // 2^261 = 2^5 * (2^32 + 977) = 31264 + 256 * 2^29 (mod p)
#define R1_29 256u

F29 void fe9_reduce(fe9& r, const uint64_t* c) {
    uint32_t p0[17], p1[17], p2[17], t[18];
#pragma unroll
    for (int k = 0; k < 17; k++) {
        p0[k] = (uint32_t)c[k] & M29;
        p1[k] = (uint32_t)(c[k] >> 29) & M29;
        p2[k] = (uint32_t)(c[k] >> 58);
    }
    t[0] = p0[0];
    t[1] = p0[1] + p1[0];
#pragma unroll
    for (int k = 2; k < 17; k++) t[k] = p0[k] + p1[k - 1] + p2[k - 2];
    t[17] = p1[16] + p2[15];   // p2[16] = 0: the top column is a single product of two top limbs
    uint64_t u[9];
    u[0] = (uint64_t)t[0] + (uint64_t)R0_29 * t[9] + (uint64_t)(R1_29 * R0_29) * t[17];
    u[1] = (uint64_t)t[1] + (uint64_t)R0_29 * t[10] + (uint64_t)R1_29 * t[9] + (uint64_t)(R1_29 * R1_29) * t[17];
#pragma unroll
    for (int k = 2; k < 8; k++) u[k] = (uint64_t)t[k] + (uint64_t)R0_29 * t[k + 9] + (uint64_t)R1_29 * t[k + 8];
    u[8] = (uint64_t)t[8] + (uint64_t)R0_29 * t[17] + (uint64_t)R1_29 * t[16];
    uint32_t s = (uint32_t)(u[8] >> 24);          // everything at or above 2^256
    uint32_t top = (uint32_t)u[8] & 0xFFFFFFu;
    u[0] += (uint64_t)977u * s;                    // 2^256 = 977 + 8 * 2^29 (mod p)
    u[1] += (uint64_t)8u * s;
    r.v[0] = (uint32_t)u[0] & M29;
#pragma unroll
    for (int k = 1; k < 8; k++) r.v[k] = ((uint32_t)u[k] & M29) + (uint32_t)(u[k - 1] >> 29);
    r.v[8] = top + (uint32_t)(u[7] >> 29);
}
F29 fe9 fe9_mul(const fe9& a, const fe9& b) {
    uint64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) {
        uint64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            int j = k - i;
            if (j >= 0 && j < 9) acc += (uint64_t)a.v[i] * b.v[j];
        }
        c[k] = acc;
    }
    fe9 r;
    fe9_reduce(r, c);
    return r;
}
F29 fe9 fe9_sqr(const fe9& a) {
    uint32_t d[9];
#pragma unroll
    for (int i = 0; i < 9; i++) d[i] = a.v[i] * 2;
    uint64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) {
        uint64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            int j = k - i;
            if (j > i && j < 9) acc += (uint64_t)d[i] * a.v[j];
        }
        if ((k & 1) == 0) acc += (uint64_t)a.v[k / 2] * a.v[k / 2];
        c[k] = acc;
    }
    fe9 r;
    fe9_reduce(r, c);
    return r;
}

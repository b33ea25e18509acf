// imad_rates.cu -- scratch microbenchmark (not a test): issue rates of the integer instructions the field
// multiplier can be built from, per SM per clock, on the box's GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_rates imad_rates.cu && ./imad_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

// A: plain 64-bit multiply-accumulate, 8 independent accumulators (IMAD.WIDE.U32 Rd, Ra, Rb, Rc)
__global__ void kA(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u;
    unsigned long long a[8];
    for (int i = 0; i < 8; i++) a[i] = x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(x), "r"(y));
    }
    unsigned long long t = 0;
    for (int i = 0; i < 8; i++) t ^= a[i];
    if (t == 0x1234567) sink[0] = (uint32_t)t;
}
// B: carry-chain form, NCH independent chains of four IMAD.WIDE.U32.X (mad.lo.cc / madc.hi.cc pairs)
template <int NCH>
__global__ void kB(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u;
    uint32_t a[NCH][8];
    for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) a[c][i] = x + c * 8 + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++)
                asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                    "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                    : "+r"(a[c][0]), "+r"(a[c][1]), "+r"(a[c][2]), "+r"(a[c][3]), "+r"(a[c][4]), "+r"(a[c][5]), "+r"(a[c][6]), "+r"(a[c][7])
                    : "r"(x), "r"(y));
    }
    uint32_t t = 0;
    for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) t ^= a[c][i];
    if (t == 0x1234567) sink[0] = t;
}
// C: carry-out only (mad.lo.cc + madc.hi with the carry consumed by a separate add): IMAD.WIDE.U32 Rd, P, ...
__global__ void kC(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u;
    uint32_t a[8][3];
    for (int c = 0; c < 8; c++) for (int i = 0; i < 3; i++) a[c][i] = x + c * 3 + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                    : "+r"(a[c][0]), "+r"(a[c][1]), "+r"(a[c][2]) : "r"(x), "r"(y));
    }
    uint32_t t = 0;
    for (int c = 0; c < 8; c++) for (int i = 0; i < 3; i++) t ^= a[c][i];
    if (t == 0x1234567) sink[0] = t;
}
// D: IADD3.X chains: 4 independent 8-limb additions
__global__ void kD(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1;
    uint32_t a[4][8];
    for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) a[c][i] = x + c * 8 + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int c = 0; c < 4; c++)
                asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\t"
                    "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %8;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %8;"
                    : "+r"(a[c][0]), "+r"(a[c][1]), "+r"(a[c][2]), "+r"(a[c][3]), "+r"(a[c][4]), "+r"(a[c][5]), "+r"(a[c][6]), "+r"(a[c][7]) : "r"(x));
    }
    uint32_t t = 0;
    for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) t ^= a[c][i];
    if (t == 0x1234567) sink[0] = t;
}
// E: 32-bit IMAD (mad.lo.u32), 16 independent
__global__ void kE(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u;
    uint32_t a[16];
    for (int i = 0; i < 16; i++) a[i] = x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
    }
    uint32_t t = 0;
    for (int i = 0; i < 16; i++) t ^= a[i];
    if (t == 0x1234567) sink[0] = t;
}
// F: IMAD.HI (mad.hi.u32), 16 independent
__global__ void kF(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u;
    uint32_t a[16];
    for (int i = 0; i < 16; i++) a[i] = x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
    }
    uint32_t t = 0;
    for (int i = 0; i < 16; i++) t ^= a[i];
    if (t == 0x1234567) sink[0] = t;
}
// G: mixed: chain-form IMAD.WIDE.X (4 chains) interleaved with an equal number of independent IADD3
__global__ void kG(uint32_t* sink) {
    uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u;
    uint32_t a[4][8], s[16];
    for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) a[c][i] = x + c * 8 + i;
    for (int i = 0; i < 16; i++) s[i] = y + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                    "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                    : "+r"(a[c][0]), "+r"(a[c][1]), "+r"(a[c][2]), "+r"(a[c][3]), "+r"(a[c][4]), "+r"(a[c][5]), "+r"(a[c][6]), "+r"(a[c][7])
                    : "r"(x), "r"(y));
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(x));
        }
    }
    uint32_t t = 0;
    for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) t ^= a[c][i];
    for (int i = 0; i < 16; i++) t ^= s[i];
    if (t == 0x1234567) sink[0] = t;
}


// H: DFMA (fma.rz.f64), 16 independent accumulators
__global__ void kH(uint32_t* sink) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 1.0 - threadIdx.x * 1e-9;
    double a[16];
    for (int i = 0; i < 16; i++) a[i] = x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(x), "d"(y));
    }
    double t = 0;
    for (int i = 0; i < 16; i++) t += a[i];
    if (t == 0.12345) sink[0] = 1;
}
// I: 16 DFMA + 16 plain IMAD.WIDE interleaved (do the FP64 and FMA-heavy pipes overlap?)
__global__ void kI(uint32_t* sink) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 1.0 - threadIdx.x * 1e-9;
    uint32_t xi = threadIdx.x * 2654435761u + 1, yi = xi ^ 0x9E3779B9u;
    double a[16];
    unsigned long long b[16];
    for (int i = 0; i < 16; i++) { a[i] = x + i; b[i] = xi + i; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(x), "d"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b[i]) : "r"(xi), "r"(yi));
        }
    }
    double t = 0; unsigned long long tb = 0;
    for (int i = 0; i < 16; i++) { t += a[i]; tb ^= b[i]; }
    if (t == 0.12345 || tb == 0x1234567) sink[0] = 1;
}
// J: 16 plain IMAD.WIDE + 16 plain IADD3 (no carry) interleaved
__global__ void kJ(uint32_t* sink) {
    uint32_t xi = threadIdx.x * 2654435761u + 1, yi = xi ^ 0x9E3779B9u;
    unsigned long long b[16];
    uint32_t s[16];
    for (int i = 0; i < 16; i++) { b[i] = xi + i; s[i] = yi + i; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b[i]) : "r"(xi), "r"(yi));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(xi));
        }
    }
    unsigned long long tb = 0;
    for (int i = 0; i < 16; i++) { tb ^= b[i]; tb ^= s[i]; }
    if (tb == 0x1234567) sink[0] = 1;
}
// K: 16 plain IMAD.WIDE + 16 SHF (funnel shift) + 16 LOP3 interleaved
__global__ void kK(uint32_t* sink) {
    uint32_t xi = threadIdx.x * 2654435761u + 1, yi = xi ^ 0x9E3779B9u;
    unsigned long long b[16];
    uint32_t s[16], q[16];
    for (int i = 0; i < 16; i++) { b[i] = xi + i; s[i] = yi + i; q[i] = yi * i; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b[i]) : "r"(xi), "r"(yi));
            asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(s[i]) : "r"(xi));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(xi), "r"(yi));
        }
    }
    unsigned long long tb = 0;
    for (int i = 0; i < 16; i++) { tb ^= b[i]; tb ^= s[i]; tb ^= q[i]; }
    if (tb == 0x1234567) sink[0] = 1;
}
// L: plain IADD3 (no carry), 32 independent
__global__ void kL(uint32_t* sink) {
    uint32_t xi = threadIdx.x * 2654435761u + 1;
    uint32_t s[32];
    for (int i = 0; i < 32; i++) s[i] = xi + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 32; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(xi));
    }
    uint32_t t = 0;
    for (int i = 0; i < 32; i++) t ^= s[i];
    if (t == 0x1234567) sink[0] = 1;
}
// M: 16 DFMA + 16 IADD3 pairs (64-bit integer add of the bit patterns, add.cc/addc)
__global__ void kM(uint32_t* sink) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 1.0 - threadIdx.x * 1e-9;
    uint32_t xi = threadIdx.x * 2654435761u + 1;
    double a[16];
    uint32_t lo[16], hi[16];
    for (int i = 0; i < 16; i++) { a[i] = x + i; lo[i] = xi + i; hi[i] = xi * i; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(x), "d"(y));
            asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %2;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(xi));
        }
    }
    double t = 0; uint32_t tb = 0;
    for (int i = 0; i < 16; i++) { t += a[i]; tb ^= lo[i] ^ hi[i]; }
    if (t == 0.12345 || tb == 0x1234567) sink[0] = 1;
}
// N: all three: 16 DFMA + 16 IMAD.WIDE + 16 IADD3
__global__ void kN(uint32_t* sink) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 1.0 - threadIdx.x * 1e-9;
    uint32_t xi = threadIdx.x * 2654435761u + 1, yi = xi ^ 0x9E3779B9u;
    double a[16];
    unsigned long long b[16];
    uint32_t s[16];
    for (int i = 0; i < 16; i++) { a[i] = x + i; b[i] = xi + i; s[i] = yi + i; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(x), "d"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b[i]) : "r"(xi), "r"(yi));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(xi));
        }
    }
    double t = 0; unsigned long long tb = 0;
    for (int i = 0; i < 16; i++) { t += a[i]; tb ^= b[i] ^ s[i]; }
    if (t == 0.12345 || tb == 0x1234567) sink[0] = 1;
}
// O: 64-bit add pairs alone (add.cc + addc), 16 independent pairs
__global__ void kO(uint32_t* sink) {
    uint32_t xi = threadIdx.x * 2654435761u + 1;
    uint32_t lo[16], hi[16];
    for (int i = 0; i < 16; i++) { lo[i] = xi + i; hi[i] = xi * i; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %2;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(xi));
    }
    uint32_t tb = 0;
    for (int i = 0; i < 16; i++) tb ^= lo[i] ^ hi[i];
    if (tb == 0x1234567) sink[0] = 1;
}
// P: native 64-bit integer add (add.u64), 16 independent
__global__ void kP(uint32_t* sink) {
    unsigned long long xi = threadIdx.x * 2654435761ull + 1;
    unsigned long long s[16];
    for (int i = 0; i < 16; i++) s[i] = xi * (i + 3);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("add.u64 %0, %0, %1;" : "+l"(s[i]) : "l"(xi));
    }
    unsigned long long tb = 0;
    for (int i = 0; i < 16; i++) tb ^= s[i];
    if (tb == 0x1234567) sink[0] = 1;
}
// Q: DADD alone (add.rz.f64), 16 independent
__global__ void kQ(uint32_t* sink) {
    double x = 1.0 + threadIdx.x * 1e-9;
    double a[16];
    for (int i = 0; i < 16; i++) a[i] = x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(x));
    }
    double t = 0;
    for (int i = 0; i < 16; i++) t += a[i];
    if (t == 0.12345) sink[0] = 1;
}

template <class K>
static double run(K kern, uint32_t* sink, int blocks, int threads) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<blocks, threads>>>(sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        kern<<<blocks, threads>>>(sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

static void sweep(uint32_t* sink, int sms, double clk, int threads, int blocks_per_sm) {
    int blocks = sms * blocks_per_sm;
    double thr = (double)blocks * threads;
    struct { const char* name; double ops_per_thread; double t; } R[32];
    int n = 0;
    R[n++] = {"A  IMAD.WIDE.U32 plain (64-bit acc), 8 independent", (double)ITERS * 32, run(kA, sink, blocks, threads)};
    R[n++] = {"B1 IMAD.WIDE.U32.X carry chains, 1 chain", (double)ITERS * 2 * 1 * 4, run(kB<1>, sink, blocks, threads)};
    R[n++] = {"B2 IMAD.WIDE.U32.X carry chains, 2 chains", (double)ITERS * 2 * 2 * 4, run(kB<2>, sink, blocks, threads)};
    R[n++] = {"B4 IMAD.WIDE.U32.X carry chains, 4 chains", (double)ITERS * 2 * 4 * 4, run(kB<4>, sink, blocks, threads)};
    R[n++] = {"B8 IMAD.WIDE.U32.X carry chains, 8 chains", (double)ITERS * 2 * 8 * 4, run(kB<8>, sink, blocks, threads)};
    R[n++] = {"C  IMAD.WIDE.U32 +carry-out, +1 IADD3.X each (count = IMAD.WIDE)", (double)ITERS * 32, run(kC, sink, blocks, threads)};
    R[n++] = {"D  IADD3.X 8-limb chains, 4 chains (count = adds)", (double)ITERS * 2 * 4 * 8, run(kD, sink, blocks, threads)};
    R[n++] = {"E  IMAD 32-bit lo, 16 independent", (double)ITERS * 32, run(kE, sink, blocks, threads)};
    R[n++] = {"F  IMAD.HI.U32, 16 independent", (double)ITERS * 32, run(kF, sink, blocks, threads)};
    R[n++] = {"G  4 carry chains + equal number of IADD (count = IMAD.WIDE)", (double)ITERS * 2 * 4 * 4, run(kG, sink, blocks, threads)};
    R[n++] = {"H  DFMA (fma.rz.f64), 16 independent", (double)ITERS * 32, run(kH, sink, blocks, threads)};
    R[n++] = {"I  DFMA + plain IMAD.WIDE 1:1 (count = DFMA)", (double)ITERS * 16, run(kI, sink, blocks, threads)};
    R[n++] = {"J  plain IMAD.WIDE + IADD3 1:1 (count = IMAD.WIDE)", (double)ITERS * 16, run(kJ, sink, blocks, threads)};
    R[n++] = {"K  plain IMAD.WIDE + SHF + LOP3 1:1:1 (count = IMAD.WIDE)", (double)ITERS * 16, run(kK, sink, blocks, threads)};
    R[n++] = {"L  IADD3 plain, 32 independent", (double)ITERS * 32, run(kL, sink, blocks, threads)};
    R[n++] = {"M  DFMA + 64-bit int add (add.cc/addc) 1:1 (count = DFMA)", (double)ITERS * 16, run(kM, sink, blocks, threads)};
    R[n++] = {"N  DFMA + IMAD.WIDE + IADD3 1:1:1 (count = DFMA)", (double)ITERS * 16, run(kN, sink, blocks, threads)};
    R[n++] = {"O  64-bit int add as add.cc/addc pairs (count = pairs)", (double)ITERS * 32, run(kO, sink, blocks, threads)};
    R[n++] = {"P  add.u64 (count = adds)", (double)ITERS * 32, run(kP, sink, blocks, threads)};
    R[n++] = {"Q  DADD (add.rz.f64), 16 independent", (double)ITERS * 32, run(kQ, sink, blocks, threads)};
    printf("--- %d threads/block, %d blocks/SM = %d warps/SM\n", threads, blocks_per_sm, threads * blocks_per_sm / 32);
    for (int i = 0; i < n; i++)
        printf("%-70s %8.3f ms  %7.1f /clk/SM   %.3e /s\n", R[i].name, R[i].t * 1e3, R[i].ops_per_thread * thr / R[i].t / clk / sms, R[i].ops_per_thread * thr / R[i].t);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = p.clockRate * 1e3;  // Hz (max)
    uint32_t* sink; cudaMalloc(&sink, 64);
    printf("GPU %s, %d SMs, max clock %.0f MHz; rates in thread-instructions per clock per SM (at max clock)\n", p.name, sms, clk / 1e6);
    sweep(sink, sms, clk, 256, 8);
    sweep(sink, sms, clk, 128, 4);
    return 0;
}

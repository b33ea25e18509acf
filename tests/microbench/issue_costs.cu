// issue_costs.cu -- scratch microbenchmark (not a test): issue cost (SM-sub-partition cycles per warp instruction)
// of the integer instructions a field multiplier is made of, alone and in pairs, to see which ones overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_costs issue_costs.cu && ./issue_costs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 16

// one kernel per instruction pattern: BODY is executed for i = 0..15 on independent registers
#define KERNEL(name, DECL, BODY, FOLD)                                         \
    __global__ void name(uint32_t* sink) {                                     \
        uint32_t x = threadIdx.x * 2654435761u + 1, y = x ^ 0x9E3779B9u, sh = (x & 1) + (sink != 0);       \
        DECL                                                                   \
        _Pragma("unroll 1") for (int it = 0; it < ITERS; it++) {               \
            _Pragma("unroll") for (int u = 0; u < 2; u++) {                    \
                _Pragma("unroll") for (int i = 0; i < NACC; i++) { BODY }      \
            }                                                                  \
        }                                                                      \
        uint32_t t = 0;                                                        \
        for (int i = 0; i < NACC; i++) { FOLD }                                \
        if (t == 0x1234567) sink[0] = t;                                       \
    }

#define D32 uint32_t a[NACC]; for (int i = 0; i < NACC; i++) a[i] = x + i;
#define D32B uint32_t a[NACC], b[NACC]; for (int i = 0; i < NACC; i++) { a[i] = x + i; b[i] = y * i + 1; }
#define D64 unsigned long long w[NACC]; for (int i = 0; i < NACC; i++) w[i] = x + i;
#define D64_32 unsigned long long w[NACC]; uint32_t a[NACC]; for (int i = 0; i < NACC; i++) { w[i] = x + i; a[i] = y + i; }
#define D64_32B unsigned long long w[NACC]; uint32_t a[NACC], b[NACC]; for (int i = 0; i < NACC; i++) { w[i] = x + i; a[i] = y + i; b[i] = y * i; }

KERNEL(k_iadd3, D32, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));, t ^= a[i];)
KERNEL(k_iadd3_3in, D32, asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(x), "r"(y));, t ^= a[i];)
KERNEL(k_lop_and_imm, D32, asm volatile("lop3.b32 %0, %0, %1, 0x1ffffff7, 0x6a;" : "+r"(a[i]) : "r"(x));, t ^= a[i];)
KERNEL(k_lop3_reg, D32, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(x), "r"(y));, t ^= a[i];)
KERNEL(k_shr_imm, D32, asm volatile("shr.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(sh));, t ^= a[i];)
KERNEL(k_shl_imm, D32, asm volatile("shl.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(sh));, t ^= a[i];)
KERNEL(k_shf_funnel, D32, asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(a[i]) : "r"(x));, t ^= a[i];)
KERNEL(k_mov_sel, D32, asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.u32 %0, %0, %1, p;}" : "+r"(a[i]) : "r"(x), "r"(y));, t ^= a[i];)
KERNEL(k_prmt, D32, asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(a[i]) : "r"(x));, t ^= a[i];)
KERNEL(k_lea, D32, asm volatile("{.reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1;}" : "+r"(a[i]) : "r"(x));, t ^= a[i];)
KERNEL(k_imad_lo, D32, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));, t ^= a[i];)
KERNEL(k_imad_hi, D32, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));, t ^= a[i];)
KERNEL(k_imad_lo_imm, D32, asm volatile("mad.lo.u32 %0, %0, 977, %1;" : "+r"(a[i]) : "r"(x));, t ^= a[i];)
KERNEL(k_imad_wide, D64, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y;, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);)
KERNEL(k_mul_wide, D64_32, asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(y)); a[i] ^= (uint32_t)w[i];, t ^= a[i];)
KERNEL(k_imad_wide_shf, D64_32, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(a[i]) : "r"(x));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ a[i];)
KERNEL(k_imad_wide_and, D64_32, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("lop3.b32 %0, %0, %1, 0x1ffffff7, 0x6a;" : "+r"(a[i]) : "r"(x));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ a[i];)
KERNEL(k_imad_wide_shr, D64_32, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("shr.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(sh));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ a[i];)
KERNEL(k_imad_wide_2iadd, D64_32B, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(y));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ a[i] ^ b[i];)
KERNEL(k_imadx_shf, D32B, if ((i & 7) == 0) asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\tmadc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;" : "+r"(a[i]), "+r"(a[i+1]), "+r"(a[i+2]), "+r"(a[i+3]), "+r"(a[i+4]), "+r"(a[i+5]), "+r"(a[i+6]), "+r"(a[i+7]) : "r"(x), "r"(y)); if ((i & 3) == 0) asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(b[i]) : "r"(x));, t ^= a[i] ^ b[i];)
KERNEL(k_ffma, float f[NACC]; for (int i = 0; i < NACC; i++) f[i] = x + i; float fx = 1.0001f; , asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fx));, t ^= __float_as_uint(f[i]);)
KERNEL(k_imad_wide_ffma, D64 float f[NACC]; for (int i = 0; i < NACC; i++) f[i] = x + i; float fx = 1.0001f; , w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fx));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ __float_as_uint(f[i]);)
KERNEL(k_iadd3_ffma, D32 float f[NACC]; for (int i = 0; i < NACC; i++) f[i] = x + i; float fx = 1.0001f; , asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fx));, t ^= a[i] ^ __float_as_uint(f[i]);)
KERNEL(k_iadd3_shf, D32B, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(b[i]) : "r"(x));, t ^= a[i] ^ b[i];)
KERNEL(k_iadd3_and, D32B, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); asm volatile("lop3.b32 %0, %0, %1, 0x1ffffff7, 0x6a;" : "+r"(b[i]) : "r"(x));, t ^= a[i] ^ b[i];)


KERNEL(k_imadx_row, D32B, if ((i & 7) == 0) { asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\tmadc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;" : "+r"(a[i]), "+r"(a[i+1]), "+r"(a[i+2]), "+r"(a[i+3]), "+r"(a[i+4]), "+r"(a[i+5]), "+r"(a[i+6]), "+r"(a[i+7]) : "r"(b[i]), "r"(y)); b[i] ^= a[i+7]; }, t ^= a[i] ^ b[i];)
KERNEL(k_add64, D64, asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((unsigned long long)x << 7));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);)
KERNEL(k_imad_wide_add64, D64 unsigned long long v[NACC]; for (int i = 0; i < NACC; i++) v[i] = y + i;, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("add.u64 %0, %0, %1;" : "+l"(v[i]) : "l"((unsigned long long)x << 7));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)v[i] ^ (uint32_t)(v[i] >> 32);)
KERNEL(k_dfma_dep, double f[NACC]; for (int i = 0; i < NACC; i++) f[i] = 1.0 + x * 1e-12 + i; double fy = 1.0 - 1e-9;, asm volatile("fma.rz.f64 %0, %0, %1, %1;" : "+d"(f[i]) : "d"(fy));, t ^= (uint32_t)__double2loint(f[i]);)
KERNEL(k_dfma_imad_wide, D64 double f[NACC]; for (int i = 0; i < NACC; i++) f[i] = 1.0 + x * 1e-12 + i; double fy = 1.0 - 1e-9;, w[i] += (unsigned long long)(uint32_t)w[(i + 1) & (NACC - 1)] * y; asm volatile("fma.rz.f64 %0, %0, %1, %1;" : "+d"(f[i]) : "d"(fy));, t ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(f[i]);)

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

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = p.clockRate * 1e3;
    uint32_t* sink; cudaMalloc(&sink, 64);
    printf("GPU %s, %d SMs, max clock %.0f MHz; cost = sub-partition cycles per loop body (one of each instruction named), 16 warps/SM\n", p.name, sms, clk / 1e6);
    int threads = 128, blocks = sms * 4;
    double warps_per_smsp = 4.0 * threads / 32 / 4;   // 4 blocks/SM, over 4 sub-partitions
#define RUN(k, label, bodies_per_iter) { double t = run(k, sink, blocks, threads); \
        printf("%-58s %8.3f ms   %6.2f cycles/body\n", label, t * 1e3, t * clk / (warps_per_smsp * ITERS * (bodies_per_iter))); }
    RUN(k_iadd3, "IADD3 (2 inputs)", 32.0)
    RUN(k_iadd3_3in, "IADD3 (3 inputs)", 32.0)
    RUN(k_lop_and_imm, "LOP3 (a ^ (x & imm))", 32.0)
    RUN(k_lop3_reg, "LOP3 three registers", 32.0)
    RUN(k_shr_imm, "SHF (shr.u32 imm)", 32.0)
    RUN(k_shl_imm, "SHL (shl.b32 imm)", 32.0)
    RUN(k_shf_funnel, "SHF funnel (two registers)", 32.0)
    RUN(k_mov_sel, "SETP + SEL", 32.0)
    RUN(k_prmt, "PRMT", 32.0)
    RUN(k_lea, "shl+add (LEA)", 32.0)
    RUN(k_imad_lo, "IMAD lo 32", 32.0)
    RUN(k_imad_hi, "IMAD.HI", 32.0)
    RUN(k_imad_lo_imm, "IMAD lo 32, immediate multiplier", 32.0)
    RUN(k_imad_wide, "IMAD.WIDE.U32 plain (64-bit addend)", 32.0)
    RUN(k_mul_wide, "mul.wide.u32 (no addend) + xor", 32.0)
    RUN(k_imad_wide_shf, "IMAD.WIDE + SHF funnel", 32.0)
    RUN(k_imad_wide_and, "IMAD.WIDE + LOP3", 32.0)
    RUN(k_imad_wide_shr, "IMAD.WIDE + SHR imm", 32.0)
    RUN(k_imad_wide_2iadd, "IMAD.WIDE + 2 IADD3", 32.0)
    RUN(k_imadx_shf, "row of 4 IMAD.WIDE.X + 2 SHF (per row)", 4.0)
    RUN(k_imadx_row, "row of 4 carry-form IMAD.WIDE (1 .CC + 3 .X), per row", 4.0)
    RUN(k_add64, "add.u64", 32.0)
    RUN(k_imad_wide_add64, "IMAD.WIDE + add.u64", 32.0)
    RUN(k_dfma_dep, "DFMA", 32.0)
    RUN(k_dfma_imad_wide, "DFMA + IMAD.WIDE", 32.0)
    RUN(k_ffma, "FFMA", 32.0)
    RUN(k_imad_wide_ffma, "IMAD.WIDE + FFMA", 32.0)
    RUN(k_iadd3_ffma, "IADD3 + FFMA", 32.0)
    RUN(k_iadd3_shf, "IADD3 + SHF funnel", 32.0)
    RUN(k_iadd3_and, "IADD3 + LOP3", 32.0)
    return 0;
}

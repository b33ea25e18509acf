// k_misc.cu -- hash-to-curve-only pipeline, batched inversion, generator table, IMAD microbenchmark.
#include "launch.h"

__global__ void __launch_bounds__(128, PLUME_H2C_MINBLOCKS) k_h2c_map(h2c_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2c_stage_map(i, a);
}
__global__ void __launch_bounds__(128, PLUME_H2C_MINBLOCKS) k_h2cw_map(h2cw_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2cw_stage_map(i, a);
}
__global__ void __launch_bounds__(128) k_h2cw_sum(h2cw_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2cw_stage_sum(i, a);
}
__global__ void __launch_bounds__(128) k_h2cw_out(h2cw_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2cw_stage_out(i, a);
}
__global__ void __launch_bounds__(128) k_fbmul_map(fbmul_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) fbmul_stage_map(i, a);
}
__global__ void __launch_bounds__(128) k_fbmul_out(fbmul_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) fbmul_stage_out(i, a);
}
__global__ void __launch_bounds__(256) k_registers(uint32_t n, const uint8_t* in32, uint64_t* out4) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) registers_body(i, in32, out4);
}
__global__ void __launch_bounds__(128) k_h2c_out(h2c_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2c_stage_out(i, a);
}
__global__ void __launch_bounds__(128) k_binv(uint32_t* Z, uint32_t* scratch, uint32_t m, uint32_t T) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) binv_body<false>(t, T, Z, scratch, m);
}
__global__ void __launch_bounds__(128) k_binv_var(uint32_t* Z, uint32_t* scratch, uint32_t m, uint32_t T) {   // small batches
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) binv_body<true>(t, T, Z, scratch, m);
}
__global__ void k_gtab_bases(uint32_t* bases, int w) {
    if (blockIdx.x == 0 && threadIdx.x == 0) gtab_bases_body(bases, w);
}
__global__ void __launch_bounds__(128) k_gtab_entries(uint32_t ne, uint32_t* tab, uint32_t* zs, const uint32_t* bases, int w) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ne) gtab_entry_body(e, tab, zs, bases, w);
}
__global__ void __launch_bounds__(128) k_gtab_norm(uint32_t ne, uint32_t* tab, const uint32_t* zs) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ne) gtab_norm_body(e, tab, zs);
}

__global__ void __launch_bounds__(128) k_sec1_compress(uint32_t n, const uint8_t* in64, uint8_t* out33) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sec1_compress_body(i, in64, out33);
}
__global__ void __launch_bounds__(128, 4) k_sec1_decompress(uint32_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sec1_decompress_body(i, in33, out64, ok);
}
__global__ void k_and_flags(uint32_t n, uint8_t* ok, const uint8_t* f0, const uint8_t* f1, const uint8_t* f2, const uint8_t* f3) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) and_flags_body(i, ok, f0, f1, f2, f3);
}
// element-wise field operation on raw little-endian limb arrays (test hook: plume_debug_fe_op)
__global__ void __launch_bounds__(128) k_debug_fe_op(int op, uint32_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x = ld_fe(a + (size_t)i * 8), y = ld_fe(b + (size_t)i * 8), r;
    switch (op) {
        case 0: r = fe_mul(x, y); break;
        case 1: r = fe_sqr(x); break;
        case 2: r = fe_add(x, y); break;
        case 3: r = fe_sub(x, y); break;
        case 4: r = fe_inv(x); break;
        case 14: r = fe_inv_var(x); break;
        case 5: r = fe_norm(x); break;
        case 6: r = fe_mul_small(x, y.v[0]); break;
        case 7: r = fe_pow_pm3d4(x); break;
        case 8: r = fe_neg(x); break;
        case 9: r = fe_sqrt_cand(x); break;
        case 10: r = fe_set_u32(fe_is_zero(x) ? 1u : 0u); break;
        case 11: r = fe_shl<3>(x); break;
        case 12: r = fe_shl<1>(x); break;
        case 13: r = fe_shl<2>(x); break;
        default: r = fe_zero();
    }
    st_fe(out + (size_t)i * 8, r);
}

// Integer-pipe microbenchmarks: the denominator of the roofline (SURVEY.md 8d "Peak").  A limb product is one
// 32x32->64-bit multiply-accumulate.  Two forms are timed and the faster one is the peak:
//   form 0 "columns": plain IMAD.WIDE.U32 Rd, Ra, Rb, Rc (64-bit addend, no carry flag) -- an 8x8 block of products
//                     accumulated into 8 independent column sums per pass;
//   form 1 "rows":    the carry-chain form the shipped multiplier uses, rows of four IMAD.WIDE.U32(.X) fused by
//                     ptxas from mad.lo.cc / madc.hi.cc pairs, 8 independent rows per pass.
// The multiplicands are refreshed from the accumulators on every pass (8 LOP3 per 64 products): with
// loop-invariant operands ptxas hoists the products out of the loop and what remains is 64-bit additions
// (that is what the round-1 version of this kernel measured; profiles/r02_issue_costs.md).
__global__ void k_imad_peak_cols(uint32_t* sink, int iters) {
    uint32_t a[8], b[8];
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = (threadIdx.x + 1) * 2654435761u + i * 40503u + blockIdx.x;
        b[i] = a[i] ^ (0x9E3779B9u * (i + 1));
        acc[i] = a[i];
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[(i + j) & 7] += (unsigned long long)a[i] * b[j];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] ^= (uint32_t)(acc[i] >> 32);
    }
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t ^= acc[i];
    if (t == 0x123456789ull) sink[0] = (uint32_t)t;  // keep the chains alive
}
__global__ void k_imad_peak_rows(uint32_t* sink, int iters) {
    // loop-invariant multiplicands are fine here: ptxas keeps the fused carry-form instructions (SASS of this
    // kernel: 16 IMAD.WIDE.U32 + 48 IMAD.WIDE.U32.X per pass, profiles/r02_issue_costs.md)
    uint32_t x = (threadIdx.x + 1) * 2654435761u + blockIdx.x, y = x ^ 0x9E3779B9u;
    uint32_t r[8][8];
#pragma unroll
    for (int c = 0; c < 8; c++)
#pragma unroll
        for (int k = 0; k < 8; k++) r[c][k] = x + c * 8 + k;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                    "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                    : "+r"(r[c][0]), "+r"(r[c][1]), "+r"(r[c][2]), "+r"(r[c][3]), "+r"(r[c][4]), "+r"(r[c][5]), "+r"(r[c][6]), "+r"(r[c][7])
                    : "r"(x), "r"(y));
    }
    uint32_t t = 0;
#pragma unroll
    for (int c = 0; c < 8; c++)
#pragma unroll
        for (int k = 0; k < 8; k++) t ^= r[c][k];
    if (t == 0x1234567u) sink[0] = t;
}

static inline unsigned grid_for(uint32_t n, unsigned b) { return (n + b - 1) / b; }

cudaError_t launch_h2c_map(const h2c_args& a, cudaStream_t s) {
    k_h2c_map<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_h2cw(int stage, const h2cw_args& a, cudaStream_t s) {
    if (stage == 0) k_h2cw_map<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    else if (stage == 1) k_h2cw_sum<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    else k_h2cw_out<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_fbmul(int stage, const fbmul_args& a, cudaStream_t s) {
    if (stage == 0) k_fbmul_map<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    else k_fbmul_out<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_registers(uint32_t n, const uint8_t* in32, uint64_t* out4, cudaStream_t s) {
    k_registers<<<grid_for(n, 256), 256, 0, s>>>(n, in32, out4);
    return cudaGetLastError();
}
cudaError_t launch_h2c_out(const h2c_args& a, cudaStream_t s) {
    k_h2c_out<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_binv(uint32_t* Z, uint32_t* scratch, uint32_t m, uint32_t per_thread, cudaStream_t s, bool var) {
    if (per_thread < 1) per_thread = 1;
    uint32_t T = (m + per_thread - 1) / per_thread;
    T = (T + 31) / 32 * 32;  // whole warps so that the strided accesses stay coalesced
    if (var) k_binv_var<<<grid_for(T, 128), 128, 0, s>>>(Z, scratch, m, T);
    else k_binv<<<grid_for(T, 128), 128, 0, s>>>(Z, scratch, m, T);
    return cudaGetLastError();
}
cudaError_t launch_gtab_bases(uint32_t* bases, int w, cudaStream_t s) {
    k_gtab_bases<<<1, 32, 0, s>>>(bases, w);
    return cudaGetLastError();
}
cudaError_t launch_gtab_entries(uint32_t ne, uint32_t* tab, uint32_t* zs, const uint32_t* bases, int w, cudaStream_t s) {
    k_gtab_entries<<<grid_for(ne, 128), 128, 0, s>>>(ne, tab, zs, bases, w);
    return cudaGetLastError();
}
cudaError_t launch_gtab_norm(uint32_t ne, uint32_t* tab, const uint32_t* zs, cudaStream_t s) {
    k_gtab_norm<<<grid_for(ne, 128), 128, 0, s>>>(ne, tab, zs);
    return cudaGetLastError();
}
cudaError_t launch_sec1_compress(uint32_t n, const uint8_t* in64, uint8_t* out33, cudaStream_t s) {
    k_sec1_compress<<<grid_for(n, 128), 128, 0, s>>>(n, in64, out33);
    return cudaGetLastError();
}
cudaError_t launch_sec1_decompress(uint32_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok, cudaStream_t s) {
    k_sec1_decompress<<<grid_for(n, 128), 128, 0, s>>>(n, in33, out64, ok);
    return cudaGetLastError();
}
cudaError_t launch_and_flags(uint32_t n, uint8_t* ok, const uint8_t* f0, const uint8_t* f1, const uint8_t* f2, const uint8_t* f3, cudaStream_t s) {
    k_and_flags<<<grid_for(n, 256), 256, 0, s>>>(n, ok, f0, f1, f2, f3);
    return cudaGetLastError();
}
cudaError_t launch_debug_fe_op(int op, uint32_t n, const uint32_t* a, const uint32_t* b, uint32_t* out, cudaStream_t s) {
    k_debug_fe_op<<<grid_for(n, 128), 128, 0, s>>>(op, n, a, b, out);
    return cudaGetLastError();
}
cudaError_t launch_imad_peak(uint32_t* sink, int iters, int blocks, int threads, int form, cudaStream_t s) {
    if (form == 0) k_imad_peak_cols<<<blocks, threads, 0, s>>>(sink, iters);
    else k_imad_peak_rows<<<blocks, threads, 0, s>>>(sink, iters);
    return cudaGetLastError();
}

cudaError_t kernels_init_sign();
cudaError_t kernels_init_verify();
cudaError_t kernels_init() {
    cudaError_t e = kernels_init_sign();
    if (e != cudaSuccess) return e;
    return kernels_init_verify();
}

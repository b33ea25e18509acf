// k_misc.cu -- hash-to-curve-only pipeline, batched inversion, generator table, IMAD microbenchmark.
#include "launch.h"

__global__ void __launch_bounds__(128, PLUME_H2C_MINBLOCKS) k_h2c_map(h2c_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2c_stage_map(i, a);
}
__global__ void __launch_bounds__(128) k_h2c_out(h2c_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) h2c_stage_out(i, a);
}
__global__ void __launch_bounds__(128) k_binv(uint32_t* Z, uint32_t* scratch, uint32_t m, uint32_t T) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) binv_body(t, T, Z, scratch, m);
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
        case 5: r = fe_norm(x); break;
        case 6: r = fe_mul_small(x, y.v[0]); break;
        case 7: r = fe_pow_pm3d4(x); break;
        case 8: r = fe_neg(x); break;
        case 9: r = fe_sqrt_cand(x); break;
        case 10: r = fe_set_u32(fe_is_zero(x) ? 1u : 0u); break;
        default: r = fe_zero();
    }
    st_fe(out + (size_t)i * 8, r);
}

// 8 independent 64-bit accumulators per thread, each fed by mad.wide.u32 (IMAD.WIDE.U32):
// iters * 8 * 8 multiply-adds per thread, no memory traffic inside the loop.
__global__ void k_imad_peak(uint32_t* sink, int iters) {
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x + 1, y = x ^ 0x9E3779B9u;
    unsigned long long a0 = x, a1 = y, a2 = x + 3, a3 = y + 5, a4 = x + 7, a5 = y + 11, a6 = x + 13, a7 = y + 17;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"(y), "r"(x));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"(x), "r"(x));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"(y), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a4) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a5) : "r"(y), "r"(x));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a6) : "r"(x), "r"(x));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a7) : "r"(y), "r"(y));
        }
    }
    unsigned long long t = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (t == 0x123456789ull) sink[0] = (uint32_t)t;  // keep the chains alive
}

static inline unsigned grid_for(uint32_t n, unsigned b) { return (n + b - 1) / b; }

cudaError_t launch_h2c_map(const h2c_args& a, cudaStream_t s) {
    k_h2c_map<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_h2c_out(const h2c_args& a, cudaStream_t s) {
    k_h2c_out<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_binv(uint32_t* Z, uint32_t* scratch, uint32_t m, uint32_t per_thread, cudaStream_t s) {
    if (per_thread < 1) per_thread = 1;
    uint32_t T = (m + per_thread - 1) / per_thread;
    T = (T + 31) / 32 * 32;  // whole warps so that the strided accesses stay coalesced
    k_binv<<<grid_for(T, 128), 128, 0, s>>>(Z, scratch, m, T);
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
cudaError_t launch_imad_peak(uint32_t* sink, int iters, int blocks, int threads, cudaStream_t s) {
    k_imad_peak<<<blocks, threads, 0, s>>>(sink, iters);
    return cudaGetLastError();
}

cudaError_t kernels_init_sign();
cudaError_t kernels_init_verify();
cudaError_t kernels_init() {
    cudaError_t e = kernels_init_sign();
    if (e != cudaSuccess) return e;
    return kernels_init_verify();
}

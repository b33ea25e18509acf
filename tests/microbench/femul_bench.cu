// femul_bench.cu -- scratch experiment (not a test): throughput of the shipped 8x32 carry-chain multiplier (fe.cuh)
// against a 9x29 carry-free candidate (fe29.h), as chains of dependent multiplications / squarings per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../zk-nullifier-sig_b200/csrc -o femul_bench femul_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fe.cuh"
#include "fe29.h"

static __device__ __noinline__ fe9 fe9_mul_fn(fe9 a, fe9 b) { return fe9_mul(a, b); }
static __device__ __noinline__ fe9 fe9_sqr_fn(fe9 a) { return fe9_sqr(a); }

// MODE 0: x = x*y; y = y^2 (calls)   MODE 1: same, inlined bodies
template <int MODE>
__global__ void k_fe8(uint32_t* out, int iters) {
    fe x, y;
    for (int i = 0; i < 8; i++) { x.v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; y.v[i] = x.v[i] ^ (0x9E3779B9u * (i + 1)); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { x = fe_mul(x, y); y = fe_sqr(y); x = fe_mul(x, y); y = fe_sqr(y); }
        else { x = fe_mul_inl(x, y); y = fe_sqr_inl(y); x = fe_mul_inl(x, y); y = fe_sqr_inl(y); }
    }
    uint32_t t = 0;
    for (int i = 0; i < 8; i++) t ^= x.v[i] ^ y.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE>
__global__ void k_fe9(uint32_t* out, int iters) {
    fe9 x, y;
    for (int i = 0; i < 9; i++) { x.v[i] = (threadIdx.x * 2654435761u + i * 40503u + blockIdx.x) & (i == 8 ? 0xFFFFFFu : M29); y.v[i] = (x.v[i] ^ (0x9E3779B9u * (i + 1))) & (i == 8 ? 0xFFFFFFu : M29); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { x = fe9_mul_fn(x, y); y = fe9_sqr_fn(y); x = fe9_mul_fn(x, y); y = fe9_sqr_fn(y); }
        else { x = fe9_mul(x, y); y = fe9_sqr(y); x = fe9_mul(x, y); y = fe9_sqr(y); }
    }
    uint32_t t = 0;
    for (int i = 0; i < 9; i++) t ^= x.v[i] ^ y.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// only multiplications / only squarings (inlined), to separate the two costs
template <int WHICH>
__global__ void k_fe8_one(uint32_t* out, int iters) {
    fe x, y;
    for (int i = 0; i < 8; i++) { x.v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; y.v[i] = x.v[i] ^ (0x9E3779B9u * (i + 1)); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (WHICH == 0) { x = fe_mul_inl(x, y); y = fe_mul_inl(y, x); x = fe_mul_inl(x, y); y = fe_mul_inl(y, x); }
        else { x = fe_sqr_inl(x); y = fe_sqr_inl(y); x = fe_sqr_inl(x); y = fe_sqr_inl(y); }
    }
    uint32_t t = 0;
    for (int i = 0; i < 8; i++) t ^= x.v[i] ^ y.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int WHICH>
__global__ void k_fe9_one(uint32_t* out, int iters) {
    fe9 x, y;
    for (int i = 0; i < 9; i++) { x.v[i] = (threadIdx.x * 2654435761u + i * 40503u + blockIdx.x) & (i == 8 ? 0xFFFFFFu : M29); y.v[i] = (x.v[i] ^ (0x9E3779B9u * (i + 1))) & (i == 8 ? 0xFFFFFFu : M29); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (WHICH == 0) { x = fe9_mul(x, y); y = fe9_mul(y, x); x = fe9_mul(x, y); y = fe9_mul(y, x); }
        else { x = fe9_sqr(x); y = fe9_sqr(y); x = fe9_sqr(x); y = fe9_sqr(y); }
    }
    uint32_t t = 0;
    for (int i = 0; i < 9; i++) t ^= x.v[i] ^ y.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <class K>
static void run(const char* name, K kern, uint32_t* out, int sms, double clk, int threads, int bps, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = sms * bps;
    kern<<<blocks, threads>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        kern<<<blocks, threads>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    double ops = 4.0 * iters * blocks * threads;                   // field operations executed
    double warps_per_smsp = (double)bps * threads / 32 / 4;
    double cyc = best * 1e-3 * clk / (4.0 * iters * warps_per_smsp);  // sub-partition cycles per field operation per warp
    printf("%-44s %2d warps/SM  %8.3f ms  %.3e ops/s  %7.1f cycles/op/warp\n", name, bps * threads / 32, best, ops / (best * 1e-3), cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = p.clockRate * 1e3;
    uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * sms * 16 * 256);
    const int iters = 512;
    printf("GPU %s, %d SMs, %.0f MHz; op = one field multiplication or squaring\n", p.name, sms, clk / 1e6);
    for (int bps = 2; bps <= 8; bps *= 2) {
        run("8x32 carry chain, mul+sqr, calls", k_fe8<0>, out, sms, clk, 128, bps, iters);
        run("8x32 carry chain, mul+sqr, inlined", k_fe8<1>, out, sms, clk, 128, bps, iters);
        run("8x32 carry chain, mul only, inlined", k_fe8_one<0>, out, sms, clk, 128, bps, iters);
        run("8x32 carry chain, sqr only, inlined", k_fe8_one<1>, out, sms, clk, 128, bps, iters);
        run("9x29 carry free,  mul+sqr, calls", k_fe9<0>, out, sms, clk, 128, bps, iters);
        run("9x29 carry free,  mul+sqr, inlined", k_fe9<1>, out, sms, clk, 128, bps, iters);
        run("9x29 carry free,  mul only, inlined", k_fe9_one<0>, out, sms, clk, 128, bps, iters);
        run("9x29 carry free,  sqr only, inlined", k_fe9_one<1>, out, sms, clk, 128, bps, iters);
    }
    return 0;
}

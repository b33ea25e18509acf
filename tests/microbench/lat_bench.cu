// lat_bench.cu -- scratch experiment (not a test): LATENCY of the field operations on a lone warp (what a batch of one waits
// for), as dependent chains: out-of-line calls (the shipped form), inlined bodies, and two independent operations per call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../zk-nullifier-sig_b200/csrc -o lat_bench lat_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fe.cuh"

struct fe2 { fe a, b; };
static __device__ __noinline__ fe2 fe_sqr2(fe a, fe b) { fe2 r; r.a = fe_sqr_inl(a); r.b = fe_sqr_inl(b); return r; }
static __device__ __noinline__ fe2 fe_mul2(fe a, fe b, fe c, fe d) { fe2 r; r.a = fe_mul_inl(a, b); r.b = fe_mul_inl(c, d); return r; }

template <int MODE>
__global__ void k_lat(uint32_t* out, long long* cyc, int iters) {
    fe x, y;
    for (int i = 0; i < 8; i++) { x.v[i] = threadIdx.x * 2654435761u + i * 40503u + 7; y.v[i] = x.v[i] ^ (0x9E3779B9u * (i + 1)); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { x = fe_sqr(x); }                                  // one chain of squarings, calls
        if (MODE == 1) { x = fe_mul(x, y); }                               // one chain of multiplications, calls
        if (MODE == 2) { x = fe_sqr(x); y = fe_sqr(y); }                   // two independent chains, one call each
        if (MODE == 3) { fe2 r = fe_sqr2(x, y); x = r.a; y = r.b; }        // two independent chains, one call for both
        if (MODE == 4) { x = fe_mul(x, x); y = fe_mul(y, y); }
        if (MODE == 5) { fe2 r = fe_mul2(x, x, y, y); x = r.a; y = r.b; }
        if (MODE == 6) { x = fe_sqr_inl(x); }
        if (MODE == 7) { x = fe_sqr_inl(x); y = fe_sqr_inl(y); }           // two chains, inlined: ptxas free to interleave
    }
    long long t1 = clock64();
    uint32_t t = 0;
    for (int i = 0; i < 8; i++) t ^= x.v[i] ^ y.v[i];
    out[threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 8);
    const int iters = 20000;
    const char* names[] = {"sqr chain (call)", "mul chain (call)", "2 sqr chains, 2 calls", "2 sqr chains, fe_sqr2", "2 mul chains, 2 calls",
                           "2 mul chains, fe_mul2", "sqr chain inlined", "2 sqr chains inlined"};
    const int ops[] = {1, 1, 2, 2, 2, 2, 1, 2};
    for (int lanes : {1, 32}) {
        for (int m = 0; m < 8; m++) {
            for (int rep = 0; rep < 2; rep++) {
                switch (m) {
                    case 0: k_lat<0><<<1, lanes>>>(out, cyc, iters); break;
                    case 1: k_lat<1><<<1, lanes>>>(out, cyc, iters); break;
                    case 2: k_lat<2><<<1, lanes>>>(out, cyc, iters); break;
                    case 3: k_lat<3><<<1, lanes>>>(out, cyc, iters); break;
                    case 4: k_lat<4><<<1, lanes>>>(out, cyc, iters); break;
                    case 5: k_lat<5><<<1, lanes>>>(out, cyc, iters); break;
                    case 6: k_lat<6><<<1, lanes>>>(out, cyc, iters); break;
                    case 7: k_lat<7><<<1, lanes>>>(out, cyc, iters); break;
                }
                cudaDeviceSynchronize();
            }
            printf("lanes %2d  %-26s %7.1f cycles per iteration, %7.1f per field operation\n", lanes, names[m], (double)cyc[0] / iters,
                   (double)cyc[0] / iters / ops[m]);
        }
    }
    return 0;
}

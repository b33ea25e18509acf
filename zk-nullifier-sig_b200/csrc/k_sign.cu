// k_sign.cu -- stage kernels of the signing pipeline (bodies in stages.cuh).
#include "launch.h"

__global__ void __launch_bounds__(128) k_sign_fixed(sign_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) sign_stage_fixed(i, a);
}
__global__ void __launch_bounds__(128, PLUME_H2C_MINBLOCKS) k_sign_h2c(sign_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) sign_stage_h2c(i, a);
}
__global__ void __launch_bounds__(VB_BLOCK, PLUME_VB_MINBLOCKS) k_sign_varbase(sign_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
#ifdef PLUME_SIGN_WINDOWED
    if (i < a.n) sign_stage_varbase(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS);          // windowed ladder per scalar (132 doublings each)
#else
    if (i < a.n) sign_stage_varbase_comb(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS);     // signed comb, shared teeth
#endif
}
#ifndef PLUME_ST_MINBLOCKS
#define PLUME_ST_MINBLOCKS 6   // comb table per 2^20 items: 4 blocks/SM 9.94 ms, 5: 9.86, 6: 9.92; 6 keeps the host path's 227 328-item chunks at whole waves
#endif
#ifndef PLUME_SL_MINBLOCKS
#define PLUME_SL_MINBLOCKS 6   // comb ladders (2^21 threads): 4 blocks/SM 16.06 ms, 5: 15.62, 6: 15.69 (whole waves per host chunk, as above)
#endif
__global__ void __launch_bounds__(VB_BLOCK, PLUME_ST_MINBLOCKS) k_sign_comb_tab(sign_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) sign_stage_varbase_tab(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS);
}
__global__ void __launch_bounds__(VB_BLOCK, PLUME_SL_MINBLOCKS) k_sign_comb_lad(sign_args a) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * a.n) sign_stage_varbase_lad(idx, a, a.vbtab);
}
__global__ void __launch_bounds__(128) k_sign_final(sign_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) sign_stage_final(i, a);
}

static inline unsigned grid_for(uint32_t n, unsigned b) { return (n + b - 1) / b; }

cudaError_t launch_sign_fixed(const sign_args& a, cudaStream_t s) {
    k_sign_fixed<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_sign_h2c(const sign_args& a, cudaStream_t s) {
    k_sign_h2c<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_sign_varbase(const sign_args& a, cudaStream_t s) {
    k_sign_varbase<<<grid_for(a.n, VB_BLOCK), VB_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_sign_comb_tab(const sign_args& a, cudaStream_t s) {
    k_sign_comb_tab<<<grid_for(a.n, VB_BLOCK), VB_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_sign_comb_lad(const sign_args& a, cudaStream_t s) {
    k_sign_comb_lad<<<grid_for(2 * a.n, VB_BLOCK), VB_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_sign_final(const sign_args& a, cudaStream_t s) {
    k_sign_final<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t kernels_init_sign() { return cudaSuccess; }

// k_verify.cu -- stage kernels of the verification pipeline (bodies in stages.cuh).
#include "launch.h"

__global__ void __launch_bounds__(128, PLUME_H2C_MINBLOCKS) k_verify_h2c(verify_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) verify_stage_h2c(i, a);
}
#ifndef PLUME_VA_MINBLOCKS
#define PLUME_VA_MINBLOCKS 6   // k_verify_mul_a per 2^20 items: 4 blocks/SM 19.92 ms, 5: 19.35, 6: 19.45; the ladder of mul_b is best at 4
#endif
#ifndef PLUME_VB2_MINBLOCKS
#define PLUME_VB2_MINBLOCKS 6   // the ladder of h*s - nul*c per 2^20 items (round 2, 128-byte table entries): 4 blocks/SM 23.50 ms, 5: 23.03, 6: 22.90
#endif
__global__ void __launch_bounds__(PLUME_VM_BLOCK, PLUME_VM_MINBLOCKS) k_verify_tab_b(verify_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n)
        verify_stage_mul_b1(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS, a.vbtab + (size_t)i * VB_ITEM_WORDS + VB_TAB_WORDS);
}
__global__ void __launch_bounds__(PLUME_VM_BLOCK, PLUME_VB2_MINBLOCKS) k_verify_lad_b(verify_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n)
        verify_stage_mul_b2(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS, a.vbtab + (size_t)i * VB_ITEM_WORDS + VB_TAB_WORDS);
}
__global__ void __launch_bounds__(PLUME_VM_BLOCK, PLUME_VA_MINBLOCKS) k_verify_mul_a(verify_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) verify_stage_mul_a(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS + 2 * VB_TAB_WORDS);
}
__global__ void __launch_bounds__(128) k_verify_final(verify_args a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) verify_stage_final(i, a);
}

static inline unsigned grid_for(uint32_t n, unsigned b) { return (n + b - 1) / b; }

cudaError_t launch_verify_h2c(const verify_args& a, cudaStream_t s) {
    k_verify_h2c<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_verify_tab_b(const verify_args& a, cudaStream_t s) {
    k_verify_tab_b<<<grid_for(a.n, PLUME_VM_BLOCK), PLUME_VM_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_verify_lad_b(const verify_args& a, cudaStream_t s) {
    k_verify_lad_b<<<grid_for(a.n, PLUME_VM_BLOCK), PLUME_VM_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_verify_mul_a(const verify_args& a, cudaStream_t s) {
    k_verify_mul_a<<<grid_for(a.n, PLUME_VM_BLOCK), PLUME_VM_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_verify_final(const verify_args& a, cudaStream_t s) {
    k_verify_final<<<grid_for(a.n, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t kernels_init_verify() {
    return cudaSuccess;
}

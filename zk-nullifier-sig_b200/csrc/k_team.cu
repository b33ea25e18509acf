// k_team.cu -- the small-batch kernels: 2 or 4 neighbouring lanes per item (bodies in stages_team.cuh).
#include "launch.h"
#include "stages_team.cuh"

// T lanes per item on T * n threads; `mask` = the warp's lanes inside the batch (whole teams, T divides 32)
#define TEAM_PROLOGUE(T, n)                                              \
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;         \
    const bool live = idx < (T) * (n);                                   \
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, live);              \
    if (!live) return;

__global__ void __launch_bounds__(128) k_sign_fixed_team(sign_args a) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * a.n) sign_stage_fixed_team(idx, a);
}
__global__ void __launch_bounds__(128) k_sign_h2c_team(sign_args a) {
    TEAM_PROLOGUE(2, a.n)
    sign_stage_h2c_team(mask, idx, a);
}
__global__ void __launch_bounds__(128) k_sign_comb_lad_team(sign_args a) {
    TEAM_PROLOGUE(4, a.n)
    sign_stage_varbase_lad_team(mask, idx, a, a.vbtab);
}
__global__ void __launch_bounds__(128) k_verify_h2c_team(verify_args a) {
    TEAM_PROLOGUE(2, a.n)
    verify_stage_h2c_team(mask, idx, a);
}
__global__ void __launch_bounds__(128) k_verify_lad_b_team(verify_args a) {
    TEAM_PROLOGUE(4, a.n)
    verify_stage_mul_b2_team(mask, idx, a, a.vbtab);
}
__global__ void __launch_bounds__(128) k_verify_mul_a_team(verify_args a) {
    TEAM_PROLOGUE(4, a.n)
    verify_stage_mul_a_team(mask, idx, a, a.vbtab + (size_t)(idx >> 2) * VB_ITEM_WORDS + 2 * VB_TAB_WORDS);
}

static inline unsigned grid_for(uint32_t n, unsigned b) { return (n + b - 1) / b; }

#define TEAM_LAUNCH(name, kernel, args_t, T)                                  \
    cudaError_t name(const args_t& a, cudaStream_t s) {                       \
        kernel<<<grid_for((T) * a.n, 128), 128, 0, s>>>(a);                   \
        return cudaGetLastError();                                            \
    }
TEAM_LAUNCH(launch_sign_fixed_team, k_sign_fixed_team, sign_args, 2)
TEAM_LAUNCH(launch_sign_h2c_team, k_sign_h2c_team, sign_args, 2)
TEAM_LAUNCH(launch_sign_comb_lad_team, k_sign_comb_lad_team, sign_args, 4)
TEAM_LAUNCH(launch_verify_h2c_team, k_verify_h2c_team, verify_args, 2)
TEAM_LAUNCH(launch_verify_lad_b_team, k_verify_lad_b_team, verify_args, 4)
TEAM_LAUNCH(launch_verify_mul_a_team, k_verify_mul_a_team, verify_args, 4)

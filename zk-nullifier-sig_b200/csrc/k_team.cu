// k_team.cu -- the small-batch kernels: 2 or 4 neighbouring lanes per item (bodies in stages_team.cuh).
// This translation unit: hash_to_curve (multiplier out of line, as everywhere else).  k_team_lad.cu: the ladder and table
// kernels with the multiplier INLINED -- a lone warp waits for latency, not for issue slots or the instruction cache, and the
// operand marshalling around the out-of-line multiplier is 10-15 % of a ladder's chain (batch of one: h*s - nul*c 422 -> 371 us,
// G*s - pk*c 506 -> 445 us, comb ladders 172 -> 156 us; hash_to_curve gets slower inlined, 113 -> 121 us, and stays here).
#include "team.h"

__global__ void __launch_bounds__(128) k_sign_h2c_team(sign_args a) {
    TEAM_PROLOGUE(2, a.n)
    sign_stage_h2c_team(mask, idx, a);
}
__global__ void __launch_bounds__(128) k_verify_h2c_team(verify_args a) {
    TEAM_PROLOGUE(2, a.n)
    verify_stage_h2c_team(mask, idx, a);
}
__global__ void __launch_bounds__(128) k_h2c_map_team(h2c_args a) {
    TEAM_PROLOGUE(2, a.n)
    h2c_stage_map_team(mask, idx, a);
}
__global__ void __launch_bounds__(128) k_verify_final_team(verify_args a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) verify_stage_final_team(i, a);
}
__global__ void __launch_bounds__(128) k_sign_final_team(sign_args a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) sign_stage_final_team(i, a);
}
TEAM_LAUNCH(launch_sign_final_team, k_sign_final_team, sign_args, 1)
TEAM_LAUNCH(launch_verify_final_team, k_verify_final_team, verify_args, 1)
TEAM_LAUNCH(launch_h2c_map_team, k_h2c_map_team, h2c_args, 2)
TEAM_LAUNCH(launch_sign_h2c_team, k_sign_h2c_team, sign_args, 2)
TEAM_LAUNCH(launch_verify_h2c_team, k_verify_h2c_team, verify_args, 2)

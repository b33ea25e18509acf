// k_team_lad.cu -- the small-batch kernels whose time is point arithmetic, compiled with the field multiplier inlined
// (see k_team.cu).  Besides the team ladders this holds the one-lane-per-item table builder of the small-batch signer.
#define PLUME_INLINE_MUL
// The signer's comb with 4 teeth here (5 in the throughput kernels): the chain a lone lane walks is 99 + 33 doublings, 7
// conjugate additions for the 8 entries and 33 additions, against 104 + 26, 15 and 26 -- fewer operations in a row, more in
// total.  Table and ladders of a small batch both come from this translation unit; the per-item scratch stride is the
// (larger) one the library allocates for 5 teeth.
#ifndef PLUME_TEAM_COMB_T
#define PLUME_TEAM_COMB_T 4
#endif
#define PLUME_COMB_T PLUME_TEAM_COMB_T
#include "team.h"

__global__ void __launch_bounds__(128) k_sign_fixed_team(sign_args a) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * a.n) sign_stage_fixed_team(idx, a);
}
__global__ void __launch_bounds__(128) k_sign_comb_tab_small(sign_args a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) sign_stage_varbase_tab_jac(i, a, a.vbtab + (size_t)i * VB_ITEM_WORDS);
}
__global__ void __launch_bounds__(128) k_sign_comb_lad_team(sign_args a) {
    TEAM_PROLOGUE(4, a.n)
    sign_stage_varbase_lad_team(mask, idx, a, a.vbtab);
}
__global__ void __launch_bounds__(128) k_verify_mul_b_team(verify_args a) {   // tables and ladders of h*s - nul*c
    TEAM_PROLOGUE(4, a.n)
    verify_stage_mul_b_team(mask, idx, a, a.vbtab);
}
__global__ void __launch_bounds__(128) k_verify_mul_a_team(verify_args a) {
    TEAM_PROLOGUE(4, a.n)
    verify_stage_mul_a_team(mask, idx, a, a.vbtab + (size_t)(idx >> 2) * VB_ITEM_WORDS + 2 * VB_TAB_WORDS);
}
TEAM_LAUNCH(launch_sign_fixed_team, k_sign_fixed_team, sign_args, 2)
TEAM_LAUNCH(launch_sign_comb_tab_small, k_sign_comb_tab_small, sign_args, 1)
TEAM_LAUNCH(launch_sign_comb_lad_team, k_sign_comb_lad_team, sign_args, 4)
TEAM_LAUNCH(launch_verify_mul_b_team, k_verify_mul_b_team, verify_args, 4)
TEAM_LAUNCH(launch_verify_mul_a_team, k_verify_mul_a_team, verify_args, 4)

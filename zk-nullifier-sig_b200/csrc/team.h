// team.h -- shared by the two translation units of the small-batch kernels (k_team.cu, k_team_lad.cu).
#pragma once
#include "launch.h"
#include "stages_team.cuh"

// T lanes per item on T * n threads; `mask` = the warp's lanes inside the batch (whole teams, T divides 32)
#define TEAM_PROLOGUE(T, n)                                              \
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;         \
    const bool live = idx < (T) * (n);                                   \
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, live);              \
    if (!live) return;

#define TEAM_LAUNCH(name, kernel, args_t, T)                                          \
    cudaError_t name(const args_t& a, cudaStream_t s) {                               \
        kernel<<<((T) * a.n + 127) / 128, 128, 0, s>>>(a);                            \
        return cudaGetLastError();                                                    \
    }

// launch.h -- host-side launchers of the stage kernels (one translation unit per kernel family so
// that nvcc compiles them in parallel).  Every launcher enqueues on `stream` and returns the
// cudaError_t of the launch.
#pragma once
#include <cuda_runtime.h>
#include "stages.cuh"

// Variable-base kernels: threads per block and minimum resident blocks per SM (register cap).  The per-thread tables live
// in a global scratch array (L1/L2 resident, one 128-byte line per entry).  Register caps, measured on the B200 per 2^20
// items (round 1): k_sign_varbase 4 blocks/SM (128 registers) 30.14 ms, 5 (96) 29.67, 6 (80) 29.43 -- the spills of the
// table-building code cost less than the extra resident warps bring.
#ifndef PLUME_VB_BLOCK
#define PLUME_VB_BLOCK 128
#endif
#ifndef PLUME_VB_MINBLOCKS
#define PLUME_VB_MINBLOCKS 6
#endif
#define VB_BLOCK PLUME_VB_BLOCK
// hash-to-curve kernels: register cap (blocks of 128 threads per SM)
#ifndef PLUME_H2C_MINBLOCKS
#define PLUME_H2C_MINBLOCKS 4
#endif
// the verifier's double-base ladder keeps two tables per thread: always in global scratch
#ifndef PLUME_VM_BLOCK
#define PLUME_VM_BLOCK 128
#endif
#ifndef PLUME_VM_MINBLOCKS
#define PLUME_VM_MINBLOCKS 3
#endif

cudaError_t launch_sign_fixed(const sign_args& a, cudaStream_t s);
cudaError_t launch_sign_h2c(const sign_args& a, cudaStream_t s);
cudaError_t launch_sign_varbase(const sign_args& a, cudaStream_t s);    // table and ladders of h^r, h^sk in one kernel ...
cudaError_t launch_sign_comb_tab(const sign_args& a, cudaStream_t s);   // ... or as two: the comb table (n threads)
cudaError_t launch_sign_comb_lad(const sign_args& a, cudaStream_t s);   //     and the ladders (2n threads)
cudaError_t launch_sign_final(const sign_args& a, cudaStream_t s);
cudaError_t launch_verify_h2c(const verify_args& a, cudaStream_t s);
cudaError_t launch_verify_tab_b(const verify_args& a, cudaStream_t s);   // the two window tables of h*s - nul*c (first: reads the inverted Z of h) ...
cudaError_t launch_verify_lad_b(const verify_args& a, cudaStream_t s);   // ... and its double-base ladder
cudaError_t launch_verify_mul_a(const verify_args& a, cudaStream_t s);   // G*s - pk*c
cudaError_t launch_verify_final(const verify_args& a, cudaStream_t s);
// small batches (k_team.cu, k_team_lad.cu): 2 or 4 neighbouring lanes per item, same workspace conventions as the kernels
// above; `_small`: one lane per item, the signer's table builder compiled for latency
cudaError_t launch_sign_comb_tab_small(const sign_args& a, cudaStream_t s);
cudaError_t launch_h2c_map_team(const h2c_args& a, cudaStream_t s);
cudaError_t launch_sign_fixed_team(const sign_args& a, cudaStream_t s);
cudaError_t launch_sign_h2c_team(const sign_args& a, cudaStream_t s);
cudaError_t launch_sign_comb_lad_team(const sign_args& a, cudaStream_t s);
cudaError_t launch_verify_h2c_team(const verify_args& a, cudaStream_t s);
cudaError_t launch_verify_mul_b_team(const verify_args& a, cudaStream_t s);    // tables and ladders of h*s - nul*c in one
cudaError_t launch_verify_mul_a_team(const verify_args& a, cudaStream_t s);
cudaError_t launch_verify_final_team(const verify_args& a, cudaStream_t s);
cudaError_t launch_sign_final_team(const sign_args& a, cudaStream_t s);
cudaError_t launch_h2c_map(const h2c_args& a, cudaStream_t s);
cudaError_t launch_h2c_out(const h2c_args& a, cudaStream_t s);
cudaError_t launch_h2cw(int stage, const h2cw_args& a, cudaStream_t s);   // 0 map, 1 sum, 2 out
cudaError_t launch_fbmul(int stage, const fbmul_args& a, cudaStream_t s);   // 0 map, 1 out
cudaError_t launch_registers(uint32_t n, const uint8_t* in32, uint64_t* out4, cudaStream_t s);
// batched inversion of m elements at Z (scratch same size), `per_thread` elements per thread; `var`: the division-step
// inversion (inv.cuh; small batches, where the one inversion per thread is what the caller waits for)
cudaError_t launch_binv(uint32_t* Z, uint32_t* scratch, uint32_t m, uint32_t per_thread, cudaStream_t s, bool var = false);
// generator table
cudaError_t launch_gtab_bases(uint32_t* bases, int w, cudaStream_t s);
cudaError_t launch_gtab_entries(uint32_t ne, uint32_t* tab, uint32_t* zs, const uint32_t* bases, int w, cudaStream_t s);
cudaError_t launch_gtab_norm(uint32_t ne, uint32_t* tab, const uint32_t* zs, cudaStream_t s);
// integer-pipe microbenchmark: every thread runs `iters` * 64 independent-chain IMAD.WIDE.U32
cudaError_t launch_imad_peak(uint32_t* sink, int iters, int blocks, int threads, int form, cudaStream_t s);  // 64 limb products per thread per iteration
cudaError_t launch_sec1_compress(uint32_t n, const uint8_t* in64, uint8_t* out33, cudaStream_t s);
cudaError_t launch_sec1_decompress(uint32_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok, cudaStream_t s);
cudaError_t launch_and_flags(uint32_t n, uint8_t* ok, const uint8_t* f0, const uint8_t* f1, const uint8_t* f2, const uint8_t* f3, cudaStream_t s);
cudaError_t launch_debug_fe_op(int op, uint32_t n, const uint32_t* a, const uint32_t* b, uint32_t* out, cudaStream_t s);
cudaError_t kernels_init();  // opt-in shared memory sizes

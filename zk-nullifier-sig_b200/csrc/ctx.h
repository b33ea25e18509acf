// ctx.h -- the context object behind include/plume_b200.h, shared by api.cu (single-device execution) and
// api_multi.cu (the multi-device context: one sub-context and one worker thread per GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/plume_b200.h"

enum Stage { ST_SIGN_FIXED, ST_SIGN_H2C, ST_SIGN_VARBASE, ST_SIGN_FINAL, ST_VERIFY_H2C, ST_VERIFY_FINAL, ST_H2C_MAP, ST_H2C_OUT,
             ST_BINV, ST_SEC1_COMPRESS, ST_SEC1_DECOMPRESS, ST_VERIFY_MUL_A, ST_VERIFY_MUL_B, ST_H2C_WITNESS, ST_REGISTERS,
             ST_VERIFY_TAB_B, ST_FIXED_MUL, ST_SIGN_TAB, ST_COUNT };

struct PendingCopy { void* dst; const void* src; size_t bytes; };

// One execution lane: a stream, the inter-stage workspace, the per-item table scratch and the I/O arenas of one chunk.
// Lanes 0 and 1 alternate under the host-pointer entry points (chunk k+1 uploads and computes while chunk k downloads);
// lane 2 belongs to the `_device` entry points, which run on the caller's stream.
struct Lane {
    cudaStream_t stream = nullptr;
    uint32_t* ws = nullptr;          // WS_SLOTS * ws_items * 32 bytes
    uint32_t* vbtab = nullptr;       // ws_items * VB_ITEM_WORDS words of table scratch in HBM/L2
    size_t ws_items = 0;
    uint8_t* d_io = nullptr;         // device arena for inputs and outputs of one chunk
    size_t d_io_cap = 0, d_io_used = 0;
    uint8_t* h_stage = nullptr;      // pinned staging arena (same layout as d_io), allocated when a pageable pointer shows up
    size_t h_cap = 0;
    std::vector<PendingCopy> pending;                       // staged outputs to hand to the caller after the stream drains
    std::vector<std::pair<uint8_t*, size_t>> host_wipes;    // staged secrets (sk, r) to zero after the stream drains
    bool busy = false;
};

struct Worker;     // api_multi.cu
struct CopyPool;   // api.cu: the staging copy threads of pageable callers

struct plume_ctx {
    int device = 0;
    int gw = 0;
    uint32_t* gtab = nullptr;
    size_t chunk = 0;        // largest n of one pass (what the `_device` entry points accept)
    size_t host_chunk = 0;   // pipelining granularity of the host-pointer entry points
    uint32_t binv_k = 32;
    bool binv_var = false;      // division-step inversion in the batched inversion of LARGE batches too (PLUME_BINV_VAR=1; measurement)
    uint32_t team_max = 4096;   // batches of at most this many items run the small-batch kernels (k_team.cu); 0: never
    int stage_threads = 8;   // threads of a staging memcpy (pageable callers)
    CopyPool* copy_pool = nullptr;   // created when the first pageable pointer shows up
    Lane lanes[3];
    cudaEvent_t dev_done = nullptr;   // completion of the last `_device` call (it owns lane 2's workspace until then)
    bool dev_used = false;
    bool dev_split = true;            // large `_device` batches as two half-batches on two streams (PLUME_DEVICE_SPLIT=0: off)
    cudaStream_t aux_stream = nullptr;        // second stream of small batches (independent stages side by side)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    uint64_t launches = 0;
    bool profiling = false;
    struct Ev { int stage; cudaEvent_t a, b; };
    std::vector<Ev> events;
    double stage_ms[ST_COUNT] = {0};
    uint64_t stage_n[ST_COUNT] = {0};
    // multi-device context: no lanes of its own, one sub-context and one worker thread per GPU
    std::vector<plume_ctx*> subs;
    std::vector<Worker*> workers;
};

inline bool ctx_is_multi(const plume_ctx* c) { return c && !c->subs.empty(); }

// api_multi.cu: run f(sub, first, count) on every sub-context concurrently, sub g owning items [g n / G, (g+1) n / G)
int multi_split(plume_ctx* ctx, size_t n, const std::function<int(plume_ctx*, size_t, size_t)>& f);
void multi_destroy(plume_ctx* ctx);
// api.cu
int ctx_fail(plume_ctx* c, int code, const std::string& msg);
int ctx_create_single(plume_ctx** out, int device, int fixed_window_bits, const uint32_t* gtab_from, int gtab_from_device);
size_t ctx_gtab_bytes(int w);

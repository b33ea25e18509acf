// api.cu -- the C ABI of include/plume_b200.h on ONE device: context, chunked double-buffered execution of the stage
// pipelines, host staging, per-stage event timing.  (api_multi.cu adds the multi-device context on top.)  No CPU
// fallback lives here: without a CUDA device every entry point fails.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "ctx.h"
#include "launch.h"
#ifndef PLUME_NO_NVTX
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached
#define PLUME_RANGE_PUSH(name) nvtxRangePushA(name)
#define PLUME_RANGE_POP() nvtxRangePop()
#else
#define PLUME_RANGE_PUSH(name) ((void)0)
#define PLUME_RANGE_POP() ((void)0)
#endif

namespace {

const char* const kStageNames[ST_COUNT] = {"sign_fixed", "sign_h2c", "sign_varbase", "sign_final", "verify_h2c", "verify_final",
                                           "h2c_map", "h2c_out", "binv", "sec1_compress", "sec1_decompress", "verify_mul_a",
                                           "verify_mul_b", "h2c_witness", "registers", "verify_tab_b", "fixed_mul", "sign_tab"};

thread_local std::string g_create_error;   // last plume_ctx_create* failure of THIS thread

}  // namespace

int ctx_fail(plume_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

namespace {

#define fail ctx_fail
// a CUDA call that fails returns PLUME_E_CUDA -- or PLUME_E_NOMEM when it was an allocation that did not fit
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            if (e__ == cudaErrorMemoryAllocation) cudaGetLastError();                              \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? PLUME_E_NOMEM : PLUME_E_CUDA,      \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                      \
        }                                                                                          \
    } while (0)

size_t env_size(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    long long x = atoll(v);
    return x > 0 ? (size_t)x : dflt;
}

struct ScopedDevice {
    int prev = -1;
    explicit ScopedDevice(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~ScopedDevice() { if (prev >= 0) cudaSetDevice(prev); }
};

// one stage launch, counted, inside an NVTX range and (optionally) bracketed by an event pair on the launching stream
template <class F>
int run_stage(plume_ctx* ctx, int stage, cudaStream_t s, F&& launch) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (ctx->profiling) {
        CU(cudaEventCreate(&a));
        CU(cudaEventCreate(&b));
        CU(cudaEventRecord(a, s));
    }
    PLUME_RANGE_PUSH(kStageNames[stage]);
    cudaError_t le = launch();
    PLUME_RANGE_POP();
    CU(le);
    ctx->launches++;
    if (ctx->profiling) {
        CU(cudaEventRecord(b, s));
        ctx->events.push_back({stage, a, b});
    }
    return PLUME_OK;
}
#define RUN(stage, expr)                                                             \
    do {                                                                             \
        int rc__ = run_stage(ctx, stage, s, [&]() -> cudaError_t { return (expr); }); \
        if (rc__ != PLUME_OK) return rc__;                                           \
    } while (0)

// batched inversion of m workspace elements starting at slot `slot` (WS_Z0: both Z arrays when m = 2n; WS_Z1: the second)
int binv(plume_ctx* ctx, uint32_t* ws, uint32_t n, uint32_t m, cudaStream_t s, int slot = WS_Z0, int scratch_slot = WS_P0) {
    // small batches: short chains (4 elements per inversion) and the division-step inversion -- latency, not throughput
    const bool small = n <= ctx->team_max;
    RUN(ST_BINV, launch_binv(ws + (size_t)slot * n * 8, ws + (size_t)scratch_slot * n * 8, m, small ? 4u : ctx->binv_k, s, small || ctx->binv_var));
    return PLUME_OK;
}

// Small batches (n <= team_max) leave the GPU almost idle with one thread per item; what the caller waits for is the
// length of the dependent chain, so the stages that have independent parts run them on 2 or 4 neighbouring lanes
// (k_team.cu, stages_team.cuh).  Same workspace conventions, same results.
int enqueue_sign_team(plume_ctx* ctx, sign_args a, cudaStream_t s) {
    RUN(ST_SIGN_FIXED, launch_sign_fixed_team(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_SIGN_H2C, launch_sign_h2c_team(a, s));                 // h stays Jacobian: no inversion before the table
    RUN(ST_SIGN_TAB, launch_sign_comb_tab_small(a, s));
    RUN(ST_SIGN_VARBASE, launch_sign_comb_lad_team(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 3 * a.n, s, WS_SLOTS, WS_SLOTS + 3)) return rc;   // TS_ZA, TS_ZB, TS_ZH; scratch TS_SCRATCH
    RUN(ST_SIGN_FINAL, launch_sign_final_team(a, s));
    return PLUME_OK;
}
int enqueue_sign(plume_ctx* ctx, sign_args a, cudaStream_t s) {
    if (a.n <= ctx->team_max) return enqueue_sign_team(ctx, a, s);
    RUN(ST_SIGN_FIXED, launch_sign_fixed(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_SIGN_H2C, launch_sign_h2c(a, s));
    if (int rc = binv(ctx, a.ws, a.n, a.n, s)) return rc;
#ifdef PLUME_SIGN_ONE_KERNEL
    RUN(ST_SIGN_VARBASE, launch_sign_varbase(a, s));
#else
    RUN(ST_SIGN_TAB, launch_sign_comb_tab(a, s));
    RUN(ST_SIGN_VARBASE, launch_sign_comb_lad(a, s));
#endif
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_SIGN_FINAL, launch_sign_final(a, s));
    return PLUME_OK;
}
// Batches of at most this many items leave most of the GPU idle (one thread per item): their latency is the length of the
// dependent chain of stages, so independent stages are put on two streams (verify: G*s - pk*c next to h*s - nul*c).
const uint32_t kSmallBatch = 8192;

// The small-batch verifier (stages_team.cuh): G*s - pk*c on the second stream from the start (it needs only the inputs),
// hash_to_curve, then tables and ladders of h*s - nul*c in one kernel on the Jacobian h, ONE batched inversion (Z of A, B, h).
int enqueue_verify_team(plume_ctx* ctx, verify_args a, cudaStream_t s) {
    const bool fork = ctx->aux_stream != nullptr;
    if (fork) {
        cudaStream_t sa = ctx->aux_stream;
        CU(cudaEventRecord(ctx->ev_fork, s));
        CU(cudaStreamWaitEvent(sa, ctx->ev_fork, 0));
        if (int rc = run_stage(ctx, ST_VERIFY_MUL_A, sa, [&]() -> cudaError_t { return launch_verify_mul_a_team(a, sa); })) return rc;
        CU(cudaEventRecord(ctx->ev_join, sa));
    }
    RUN(ST_VERIFY_H2C, launch_verify_h2c_team(a, s));
    RUN(ST_VERIFY_MUL_B, launch_verify_mul_b_team(a, s));
    if (fork) CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    else RUN(ST_VERIFY_MUL_A, launch_verify_mul_a_team(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 3 * a.n, s, WS_RY, WS_Z0)) return rc;   // TV_ZA, TV_ZB, TV_ZH; scratch: the idle WS_Z0..WS_P0
    RUN(ST_VERIFY_FINAL, launch_verify_final_team(a, s));
    return PLUME_OK;
}

int enqueue_verify(plume_ctx* ctx, verify_args a, cudaStream_t s) {
    if (a.n <= ctx->team_max) return enqueue_verify_team(ctx, a, s);
    const bool fork = a.n <= kSmallBatch && ctx->aux_stream != nullptr;
    RUN(ST_VERIFY_H2C, launch_verify_h2c(a, s));
    if (fork) {   // A needs the input checks of the first stage (ok[]) and nothing else: start it next to the stages of B
        cudaStream_t sa = ctx->aux_stream;
        CU(cudaEventRecord(ctx->ev_fork, s));
        CU(cudaStreamWaitEvent(sa, ctx->ev_fork, 0));
        if (int rc = run_stage(ctx, ST_VERIFY_MUL_A, sa, [&]() -> cudaError_t { return launch_verify_mul_a(a, sa); })) return rc;
        CU(cudaEventRecord(ctx->ev_join, sa));
    }
    if (int rc = binv(ctx, a.ws, a.n, a.n, s, WS_Z1)) return rc;
    // separate kernels, each with its own register budget: one fused kernel needs 168 registers (12 warps/SM), the
    // ladders alone run at 128 or fewer (16-24 warps/SM); 18 % faster in total (round 1)
    RUN(ST_VERIFY_TAB_B, launch_verify_tab_b(a, s));
    RUN(ST_VERIFY_MUL_B, launch_verify_lad_b(a, s));
    if (fork) CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    else RUN(ST_VERIFY_MUL_A, launch_verify_mul_a(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_VERIFY_FINAL, launch_verify_final(a, s));
    return PLUME_OK;
}
int enqueue_h2c(plume_ctx* ctx, h2c_args a, cudaStream_t s) {
    RUN(ST_H2C_MAP, a.n <= ctx->team_max ? launch_h2c_map_team(a, s) : launch_h2c_map(a, s));
    if (int rc = binv(ctx, a.ws, a.n, a.n, s)) return rc;
    RUN(ST_H2C_OUT, launch_h2c_out(a, s));
    return PLUME_OK;
}
int enqueue_h2cw(plume_ctx* ctx, h2cw_args a, cudaStream_t s) {
    RUN(ST_H2C_WITNESS, launch_h2cw(0, a, s));
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_H2C_WITNESS, launch_h2cw(1, a, s));
    if (int rc = binv(ctx, a.ws, a.n, a.n, s)) return rc;
    RUN(ST_H2C_WITNESS, launch_h2cw(2, a, s));
    return PLUME_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// memcpy between a caller's pageable arrays and the pinned staging arena.  One thread moves ~10 GB/s, which a 2^20-item V1
// batch (0.8 GB in and out per sign + verify) would feel, so the copies of a chunk are cut into slices and shared between the
// calling thread and a small persistent pool (threads are not spawned per copy: a chunk has seven arrays).
}  // namespace
struct CopyPool {
    struct Job { char* dst; const char* src; size_t len; };
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<Job> q;
    size_t in_flight = 0;
    bool quit = false;
    explicit CopyPool(int workers) {
        for (int i = 0; i < workers; i++) th.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(mu); quit = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
    bool take(Job& j) {   // mu held
        if (q.empty()) return false;
        j = q.front();
        q.pop_front();
        return true;
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [this] { return quit || !q.empty(); });
            if (quit) return;
            Job j;
            while (take(j)) {
                lk.unlock();
                memcpy(j.dst, j.src, j.len);
                lk.lock();
                if (--in_flight == 0) cv_done.notify_all();
            }
        }
    }
    // copy everything in `jobs` (sliced), the caller working along; returns when all of it is in place
    void run(const std::vector<Job>& jobs) {
        const size_t kSlice = (size_t)2 << 20;
        {
            std::lock_guard<std::mutex> g(mu);
            for (const Job& j : jobs)
                for (size_t o = 0; o < j.len; o += kSlice) {
                    q.push_back({j.dst + o, j.src + o, j.len - o < kSlice ? j.len - o : kSlice});
                    in_flight++;
                }
        }
        cv.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        Job j;
        while (take(j)) {
            lk.unlock();
            memcpy(j.dst, j.src, j.len);
            lk.lock();
            --in_flight;
        }
        cv_done.wait(lk, [this] { return in_flight == 0; });
    }
};
namespace {
void stage_copy_many(plume_ctx* ctx, const std::vector<CopyPool::Job>& jobs) {
    size_t total = 0;
    for (const auto& j : jobs) total += j.len;
    if (total == 0) return;
    if (total < ((size_t)4 << 20) || ctx->stage_threads <= 1) {
        for (const auto& j : jobs) memcpy(j.dst, j.src, j.len);
        return;
    }
    if (!ctx->copy_pool) ctx->copy_pool = new CopyPool(ctx->stage_threads - 1);
    ctx->copy_pool->run(jobs);
}
void stage_copy(plume_ctx* ctx, void* dst, const void* src, size_t bytes) {
    stage_copy_many(ctx, {{(char*)dst, (const char*)src, bytes}});
}

// ---- lane storage ---------------------------------------------------------------------------------------
// workspace and table scratch for `items` items: grown on demand (powers of two from 4096 up to the context's chunk), so
// that a context used for single signatures does not hold the gigabytes a 2^20-item pass needs
int lane_workspace(plume_ctx* ctx, Lane& L, size_t items) {
    if (items <= L.ws_items) return PLUME_OK;
    size_t cap = 4096;
    while (cap < items) cap <<= 1;
    if (cap > ctx->chunk) cap = ctx->chunk;
    if (L.ws) { cudaFree(L.ws); L.ws = nullptr; }
    if (L.vbtab) { cudaFree(L.vbtab); L.vbtab = nullptr; }
    L.ws_items = 0;
    // The small-batch signer (<= team_max <= 4 096 items, its own n as the slot stride) uses 6 slots past WS_SLOTS
    // (stages_team.cuh TS_*): they lie inside the allocation of any workspace for more than 5 851 items, smaller ones get the room.
    const size_t small_items = cap < 4096 ? cap : 4096;
    const size_t elems = std::max((size_t)WS_SLOTS * cap, (size_t)(WS_SLOTS + 6) * small_items);
    CU(cudaMalloc(&L.ws, elems * 32));
    CU(cudaMalloc(&L.vbtab, cap * (size_t)VB_ITEM_WORDS * 4));   // comb area (sign) / two window tables (verify) per item
    L.ws_items = cap;
    return PLUME_OK;
}
int lane_reserve(plume_ctx* ctx, Lane& L, size_t bytes) {
    if (bytes <= L.d_io_cap) return PLUME_OK;
    size_t cap = bytes + bytes / 4;
    if (L.d_io) { cudaFree(L.d_io); L.d_io = nullptr; }
    if (L.h_stage) { cudaFreeHost(L.h_stage); L.h_stage = nullptr; }
    L.d_io_cap = L.h_cap = 0;
    CU(cudaMalloc(&L.d_io, cap));
    L.d_io_cap = cap;
    return PLUME_OK;
}
int lane_host(plume_ctx* ctx, Lane& L) {   // the pinned staging arena, only for callers with pageable memory
    if (L.h_cap >= L.d_io_cap) return PLUME_OK;
    if (L.h_stage) { cudaFreeHost(L.h_stage); L.h_stage = nullptr; }
    L.h_cap = 0;
    CU(cudaHostAlloc(&L.h_stage, L.d_io_cap, cudaHostAllocDefault));
    L.h_cap = L.d_io_cap;
    return PLUME_OK;
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// carve `bytes` out of the lane arena; returns the offset
size_t lane_take(Lane& L, size_t bytes) {
    size_t off = L.d_io_used;
    L.d_io_used = align256(off + bytes);
    return off;
}
// device copy of a host input array; `secret`: wipe the staged copy once the chunk is done
int lane_input(plume_ctx* ctx, Lane& L, const void* host, size_t bytes, uint8_t** dev, bool secret = false) {
    size_t off = lane_take(L, bytes);
    *dev = L.d_io + off;
    if (bytes == 0) return PLUME_OK;
    const void* src = host;
    if (!is_pinned(host)) {
        if (int rc = lane_host(ctx, L)) return rc;
        stage_copy(ctx, L.h_stage + off, host, bytes);
        src = L.h_stage + off;
        if (secret) L.host_wipes.push_back({L.h_stage + off, bytes});
    }
    CU(cudaMemcpyAsync(*dev, src, bytes, cudaMemcpyHostToDevice, L.stream));
    return PLUME_OK;
}
uint8_t* lane_output(Lane& L, size_t bytes, size_t* off_out) {
    size_t off = lane_take(L, bytes);
    *off_out = off;
    return L.d_io + off;
}
int lane_fetch(plume_ctx* ctx, Lane& L, void* host, size_t off, size_t bytes) {
    if (bytes == 0 || host == nullptr) return PLUME_OK;
    if (is_pinned(host)) {
        CU(cudaMemcpyAsync(host, L.d_io + off, bytes, cudaMemcpyDeviceToHost, L.stream));
    } else {
        if (int rc = lane_host(ctx, L)) return rc;
        CU(cudaMemcpyAsync(L.h_stage + off, L.d_io + off, bytes, cudaMemcpyDeviceToHost, L.stream));
        L.pending.push_back({host, L.h_stage + off, bytes});
    }
    return PLUME_OK;
}
void lane_wipe_host(Lane& L) {
    for (auto& w : L.host_wipes) memset(w.first, 0, w.second);
    L.host_wipes.clear();
}
int lane_finish(plume_ctx* ctx, Lane& L) {
    if (!L.busy) return PLUME_OK;
    CU(cudaStreamSynchronize(L.stream));
    if (!L.pending.empty()) {   // all staged outputs of the chunk in one parallel copy
        std::vector<CopyPool::Job> jobs;
        for (const PendingCopy& p : L.pending) jobs.push_back({(char*)p.dst, (const char*)p.src, p.bytes});
        stage_copy_many(ctx, jobs);
    }
    L.pending.clear();
    lane_wipe_host(L);
    L.busy = false;
    return PLUME_OK;
}
// error exit of a host-pointer call: nothing of this call may stay in flight or be delivered later (the caller is about to
// release its buffers), so drain both lanes, drop the staged results and clear the staged secrets
void lanes_abort(plume_ctx* ctx) {
    for (int k = 0; k < 2; k++) {
        Lane& L = ctx->lanes[k];
        if (L.stream) cudaStreamSynchronize(L.stream);
        cudaGetLastError();
        L.pending.clear();
        lane_wipe_host(L);
        L.busy = false;
    }
}

// stage the messages of items [i0, i0+cn): returns device views
int lane_msgs(plume_ctx* ctx, Lane& L, const uint8_t* msgs, const uint64_t* offs, size_t msg_len, size_t i0, size_t cn,
              msg_view* view) {
    uint8_t* d_msgs = nullptr;
    if (offs) {
        uint64_t b0 = offs[i0], b1 = offs[i0 + cn];
        if (int rc = lane_input(ctx, L, msgs + b0, (size_t)(b1 - b0), &d_msgs)) return rc;
        // rebased offsets
        if (int rc = lane_host(ctx, L)) return rc;
        size_t off = lane_take(L, (cn + 1) * sizeof(uint64_t));
        uint64_t* h = reinterpret_cast<uint64_t*>(L.h_stage + off);
        for (size_t i = 0; i <= cn; i++) h[i] = offs[i0 + i] - b0;
        CU(cudaMemcpyAsync(L.d_io + off, h, (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, L.stream));
        view->base = d_msgs;
        view->offs = reinterpret_cast<const uint64_t*>(L.d_io + off);
        view->fixed_len = 0;
    } else {
        if (int rc = lane_input(ctx, L, msgs + i0 * msg_len, cn * msg_len, &d_msgs)) return rc;
        view->base = d_msgs;
        view->offs = nullptr;
        view->fixed_len = (uint32_t)msg_len;
    }
    return PLUME_OK;
}
size_t msgs_bytes(const uint64_t* offs, size_t msg_len, size_t i0, size_t cn) {
    return offs ? (size_t)(offs[i0 + cn] - offs[i0]) + (cn + 1) * 8 + 512 : cn * msg_len + 256;
}

// Length of the chunk of a host-pointer call that starts at item i0.  The first chunk is a third of the others (one wave
// of the 4-blocks-per-SM kernels): its upload is the part of the call no computation can hide, so it is kept short.
size_t host_chunk_len(const plume_ctx* ctx, size_t n, size_t i0) {
    size_t want = ctx->host_chunk;
    if (i0 == 0 && n > ctx->host_chunk && ctx->host_chunk >= 3 * 128) want = ctx->host_chunk / 3;
    return (n - i0 < want) ? n - i0 : want;
}

// argument checks shared by every entry point that takes messages.  host_offsets: the offsets are host memory and are
// walked here (non-decreasing, every message shorter than 4 GiB: the kernels keep lengths in 32 bits).
int check_common(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* offs, size_t msg_len, bool host_offsets) {
    if (!ctx) return PLUME_E_ARG;
    if (n > 0xFFFFFFFFull) return fail(ctx, PLUME_E_ARG, "n does not fit 32 bits");
    if (!offs && msg_len > 0xFFFFFFFFull) return fail(ctx, PLUME_E_ARG, "msg_len too large");
    if (offs && host_offsets) {
        for (size_t i = 0; i < n; i++) {
            if (offs[i + 1] < offs[i]) return fail(ctx, PLUME_E_ARG, "msg_offsets must be non-decreasing");
            if (offs[i + 1] - offs[i] > 0xFFFFFFFFull) return fail(ctx, PLUME_E_ARG, "a message is longer than 4 GiB - 1");
        }
        if (n > 0 && !msgs && offs[n] != offs[0]) return fail(ctx, PLUME_E_ARG, "msgs is null");
    } else if (n > 0 && !msgs && (offs || msg_len != 0)) {
        return fail(ctx, PLUME_E_ARG, "msgs is null");
    }
    return PLUME_OK;
}

// The two-lane pipeline of every host-pointer entry point: body(L, i0, cn) reserves, uploads, enqueues and schedules
// the downloads of one chunk on lane L.  On any failure the lanes are drained and nothing is delivered late.
template <class Body>
int run_chunks(plume_ctx* ctx, size_t n, size_t fixed_step, Body&& body) {
    ScopedDevice sd(ctx->device);
    int rc = PLUME_OK;
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n && rc == PLUME_OK; i0 += cn, k++) {
        cn = fixed_step ? (n - i0 < fixed_step ? n - i0 : fixed_step) : host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        rc = lane_finish(ctx, L);
        if (rc != PLUME_OK) break;
        L.d_io_used = 0;
        L.busy = true;
        rc = body(L, i0, cn);
    }
    for (int q = 0; q < 2 && rc == PLUME_OK; q++) rc = lane_finish(ctx, ctx->lanes[q]);
    if (rc != PLUME_OK) lanes_abort(ctx);
    return rc;
}

// `_device` entry points: lane 2's workspace on the caller's stream.  A call still running on another stream owns that
// workspace, so this one is ordered after it.
int dev_begin(plume_ctx* ctx, size_t n, cudaStream_t s) {
    if (int rc = lane_workspace(ctx, ctx->lanes[2], n)) return rc;
    if (ctx->dev_used) CU(cudaStreamWaitEvent(s, ctx->dev_done, 0));
    return PLUME_OK;
}
int dev_end(plume_ctx* ctx, cudaStream_t s) {
    CU(cudaEventRecord(ctx->dev_done, s));
    ctx->dev_used = true;
    return PLUME_OK;
}

// A large device-resident batch runs as TWO half-batches, the second on the context's auxiliary stream: every stage kernel
// ends with a partly filled last wave (2^20 items are 9.2 waves of the 6-blocks-per-SM kernels), and with two independent
// kernel sequences in flight the blocks of one fill the SMs the other's tail leaves idle -- the same effect the two lanes of
// the host-pointer path get for free.  Not while per-stage profiling is on (the event pairs would time overlapping kernels).
const size_t kSplitMin = 1 << 16;
template <class F>
int dev_run(plume_ctx* ctx, size_t n, cudaStream_t s, F&& part) {
    if (int rc = dev_begin(ctx, n, s)) return rc;
    Lane& L = ctx->lanes[2];
    if (!ctx->dev_split || ctx->profiling || n < kSplitMin) {
        if (int rc = part(0, n, L.ws, L.vbtab, s)) return rc;
    } else {
        const size_t n1 = ((n / 2) + 127) & ~(size_t)127;
        cudaStream_t sa = ctx->aux_stream;
        CU(cudaEventRecord(ctx->ev_fork, s));
        CU(cudaStreamWaitEvent(sa, ctx->ev_fork, 0));
        if (int rc = part(0, n1, L.ws, L.vbtab, s)) return rc;
        if (int rc = part(n1, n - n1, L.ws + (size_t)WS_SLOTS * n1 * 8, L.vbtab + n1 * (size_t)VB_ITEM_WORDS, sa)) return rc;
        CU(cudaEventRecord(ctx->ev_join, sa));
        CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    }
    return dev_end(ctx, s);
}

template <class T> T* at(T* p, size_t first, size_t width) { return p ? p + first * width : nullptr; }
// message arguments of the items from `first` on: fixed-length records move the base, an offsets array moves the offsets
const uint8_t* msgs_at(const uint8_t* msgs, const uint64_t* offs, size_t msg_len, size_t first) { return (offs || !msgs) ? msgs : msgs + first * msg_len; }
const uint64_t* offs_at(const uint64_t* offs, size_t first) { return offs ? offs + first : nullptr; }

struct DevBuf {   // temporaries of plume_ctx_create: freed on every exit path
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

size_t ctx_gtab_bytes(int w) { return ((size_t)((256 + w - 1) / w) << w) * 64; }

// gtab_from: a finished generator table of the same window width on device gtab_from_device (peer copy instead of
// rebuilding it; api_multi.cu's broadcast option), or null to build it here.
int ctx_create_single(plume_ctx** out, int device, int fixed_window_bits, const uint32_t* gtab_from, int gtab_from_device) {
    plume_ctx* ctx = nullptr;  // CU() reports into g_create_error while ctx is null
    if (!out) return fail(nullptr, PLUME_E_ARG, "out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PLUME_E_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, PLUME_E_NO_DEVICE, "device ordinal out of range");
    // default 22 bits: 12 windows x 2^22 affine points = 3.2 GB of HBM, 12 additions per fixed-base multiplication (180 GB
    // per GPU is there to be used).  Measured per 2^20 items, sign_fixed / verify_mul_a: 16 bits (64 MiB, 16 additions)
    // 3.44 / 19.46 ms, 18 bits 3.22 / 19.05, 20 bits (872 MB, 13) 2.78 / 18.81, 22 bits 2.55 / 18.69.
    int w = fixed_window_bits ? fixed_window_bits : (int)env_size("PLUME_FIXED_WINDOW", 22);
    if (w < 4 || w > 24) return fail(nullptr, PLUME_E_ARG, "fixed_window_bits must be in 4..24");
    ScopedDevice sd(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, PLUME_E_NO_DEVICE, std::string("device is not sm_100 class: ") + prop.name);
    CU(kernels_init());
    plume_ctx* c = new plume_ctx();
    c->device = device;
    c->gw = w;
    // capacity of one pass (what the `_device` entry points accept): large, so that a device-resident batch is one launch
    // per stage with a negligible tail
    c->chunk = env_size("PLUME_CHUNK_ITEMS", (size_t)1 << 20);
    // granularity of the host-pointer entry points: 3 full waves of the kernels that hold 4 blocks of 128 threads per SM
    // = 4 full waves of the one that holds 3 (k_verify_tab_b), i.e. no tail, and short enough that the first upload and
    // the last download of a call -- the part the two lanes cannot overlap -- are a small fraction of it
    c->host_chunk = env_size("PLUME_HOST_CHUNK_ITEMS", (size_t)prop.multiProcessorCount * 128 * 12);
    if (c->host_chunk > c->chunk) c->host_chunk = c->chunk;
    c->binv_k = (uint32_t)env_size("PLUME_BINV_K", 32);   // elements per inversion: 8 -> 0.42 ms per 2^21 elements, 16 -> 0.27, 32 -> 0.21, 64 -> 0.21
    c->stage_threads = (int)env_size("PLUME_STAGE_THREADS", 8);
    c->binv_var = env_size("PLUME_BINV_VAR", 0) != 0;
    if (const char* e = getenv("PLUME_TEAM_MAX")) c->team_max = (uint32_t)strtoul(e, nullptr, 10);   // 0 switches the small-batch kernels off
    if (c->team_max > 4096) c->team_max = 4096;   // the small-batch signer's extra workspace slots (lane_workspace)
    { const char* v = getenv("PLUME_DEVICE_SPLIT"); c->dev_split = !(v && v[0] == '0'); }
    struct Guard { plume_ctx* c; ~Guard() { if (c) plume_ctx_destroy(c); } } guard{c};
    for (int k = 0; k < 2; k++) CU(cudaStreamCreateWithFlags(&c->lanes[k].stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->dev_done, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    const int nwin = (256 + w - 1) / w;
    const size_t ne = (size_t)nwin << w;
    CU(cudaMalloc(&c->gtab, ne * 64));
    cudaStream_t s = c->lanes[0].stream;
    if (gtab_from) {
        CU(cudaMemcpyPeerAsync(c->gtab, device, gtab_from, gtab_from_device, ne * 64, s));
    } else {
        // generator table: entries -> batched inversion -> affine
        DevBuf bases, zs, scratch;
        CU(cudaMalloc(&bases.p, (size_t)nwin * 64));
        CU(cudaMalloc(&zs.p, ne * 32));
        CU(cudaMalloc(&scratch.p, ne * 32));
        CU(launch_gtab_bases((uint32_t*)bases.p, w, s));
        CU(launch_gtab_entries((uint32_t)ne, c->gtab, (uint32_t*)zs.p, (const uint32_t*)bases.p, w, s));
        CU(launch_binv((uint32_t*)zs.p, (uint32_t*)scratch.p, (uint32_t)ne, 16, s));
        CU(launch_gtab_norm((uint32_t)ne, c->gtab, (const uint32_t*)zs.p, s));
        CU(cudaStreamSynchronize(s));
    }
    CU(cudaStreamSynchronize(s));
    guard.c = nullptr;
    *out = c;
    return PLUME_OK;
}

extern "C" {

int plume_version(void) { return PLUME_ABI_VERSION; }

const char* plume_last_error(const plume_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

size_t plume_ctx_chunk_items(const plume_ctx* ctx) {
    if (!ctx) return 0;
    return ctx_is_multi(ctx) ? ctx->subs[0]->chunk : ctx->chunk;
}
uint64_t plume_ctx_launch_count(const plume_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t t = ctx->launches;
    for (const plume_ctx* s : ctx->subs) t += s->launches;
    return t;
}
int plume_ctx_device_count(const plume_ctx* ctx) { return !ctx ? 0 : ctx_is_multi(ctx) ? (int)ctx->subs.size() : 1; }
plume_ctx* plume_ctx_sub(plume_ctx* ctx, int i) {
    if (!ctx) return nullptr;
    if (!ctx_is_multi(ctx)) return i == 0 ? ctx : nullptr;
    return (i >= 0 && i < (int)ctx->subs.size()) ? ctx->subs[i] : nullptr;
}

void plume_ctx_destroy(plume_ctx* ctx) {
    if (!ctx) return;
    if (ctx_is_multi(ctx)) { multi_destroy(ctx); return; }
    ScopedDevice sd(ctx->device);
    cudaDeviceSynchronize();
    for (auto& e : ctx->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (Lane& L : ctx->lanes) {
        // the arenas held secret keys and nonces (already wiped chunk by chunk; once more, whatever path left them)
        if (L.d_io) { cudaMemset(L.d_io, 0, L.d_io_cap); cudaFree(L.d_io); }
        if (L.h_stage) { memset(L.h_stage, 0, L.h_cap); cudaFreeHost(L.h_stage); }
        if (L.ws) cudaFree(L.ws);
        if (L.vbtab) cudaFree(L.vbtab);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    delete ctx->copy_pool;
    if (ctx->dev_done) cudaEventDestroy(ctx->dev_done);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->gtab) cudaFree(ctx->gtab);
    delete ctx;
}

int plume_ctx_create(plume_ctx** out, int device, int fixed_window_bits) {
    return ctx_create_single(out, device, fixed_window_bits, nullptr, 0);
}

int plume_ctx_set_profiling(plume_ctx* ctx, int on) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) {
        for (plume_ctx* s : ctx->subs) if (int rc = plume_ctx_set_profiling(s, on)) return rc;
        return PLUME_OK;
    }
    ScopedDevice sd(ctx->device);
    cudaDeviceSynchronize();
    for (auto& e : ctx->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    ctx->events.clear();
    for (int i = 0; i < ST_COUNT; i++) { ctx->stage_ms[i] = 0; ctx->stage_n[i] = 0; }
    ctx->profiling = on != 0;
    return PLUME_OK;
}

double plume_ctx_stage_ms(plume_ctx* ctx, const char* stage, uint64_t* launches) {
    if (!ctx || !stage) return -1.0;
    int id = -1;
    for (int i = 0; i < ST_COUNT; i++) if (strcmp(stage, kStageNames[i]) == 0) id = i;
    if (id < 0) return -1.0;
    if (ctx_is_multi(ctx)) {   // summed over the devices
        double t = 0;
        uint64_t k = 0, kk = 0;
        for (plume_ctx* s : ctx->subs) { t += plume_ctx_stage_ms(s, stage, &kk); k += kk; }
        if (launches) *launches = k;
        return t;
    }
    ScopedDevice sd(ctx->device);
    cudaDeviceSynchronize();
    for (auto& e : ctx->events) {  // fold finished event pairs into the sums
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { ctx->stage_ms[e.stage] += ms; ctx->stage_n[e.stage]++; }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    ctx->events.clear();
    if (launches) *launches = ctx->stage_n[id];
    return ctx->stage_ms[id];
}

int plume_measure_imad_rates(plume_ctx* ctx, int iters, double* plain_lp_per_s, double* carry_lp_per_s) {
    if (!ctx || iters <= 0) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return plume_measure_imad_rates(ctx->subs[0], iters, plain_lp_per_s, carry_lp_per_s);
    ScopedDevice sd(ctx->device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    DevBuf sink;
    CU(cudaMalloc(&sink.p, 64));
    cudaStream_t s = ctx->lanes[0].stream;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double best[2] = {0, 0};
    for (int form = 0; form < 2; form++) {
        CU(launch_imad_peak((uint32_t*)sink.p, iters, blocks, threads, form, s));  // warm-up
        for (int rep = 0; rep < 5; rep++) {
            CU(cudaEventRecord(a, s));
            CU(launch_imad_peak((uint32_t*)sink.p, iters, blocks, threads, form, s));
            CU(cudaEventRecord(b, s));
            CU(cudaEventSynchronize(b));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, a, b));
            double rate = (double)blocks * threads * (double)iters * 64.0 / (ms * 1e-3);
            if (rate > best[form]) best[form] = rate;
        }
    }
    ctx->launches += 12;
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (plain_lp_per_s) *plain_lp_per_s = best[0];
    if (carry_lp_per_s) *carry_lp_per_s = best[1];
    return PLUME_OK;
}

int plume_measure_imad_peak(plume_ctx* ctx, int iters, double* lp_per_s) {
    if (!lp_per_s) return PLUME_E_ARG;
    double plain = 0, carry = 0;
    int rc = plume_measure_imad_rates(ctx, iters, &plain, &carry);
    if (rc != PLUME_OK) return rc;
    *lp_per_s = plain > carry ? plain : carry;
    return PLUME_OK;
}

int plume_debug_fe_op(plume_ctx* ctx, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    if (!ctx || !a || !b || !out) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return plume_debug_fe_op(ctx->subs[0], op, n, a, b, out);
    if (n == 0) return PLUME_OK;
    ScopedDevice sd(ctx->device);
    DevBuf da, db, dout;
    cudaStream_t s = ctx->lanes[0].stream;
    CU(cudaMalloc(&da.p, n * 32));
    CU(cudaMalloc(&db.p, n * 32));
    CU(cudaMalloc(&dout.p, n * 32));
    CU(cudaMemcpyAsync(da.p, a, n * 32, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(db.p, b, n * 32, cudaMemcpyHostToDevice, s));
    CU(launch_debug_fe_op(op, (uint32_t)n, (const uint32_t*)da.p, (const uint32_t*)db.p, (uint32_t*)dout.p, s));
    ctx->launches++;
    CU(cudaMemcpyAsync(out, dout.p, n * 32, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return PLUME_OK;
}

// test hook: the staged copies of the secrets of the last host-pointer calls.  Reads back lane `lane`'s device arena and
// pinned staging arena (host pointers; either may be null) -- tests/ check that no secret key or nonce survives a call.
int plume_debug_read_arena(plume_ctx* ctx, int lane, uint8_t* dev_copy, uint8_t* host_copy, size_t cap, size_t* dev_bytes,
                           size_t* host_bytes) {
    if (!ctx || ctx_is_multi(ctx) || lane < 0 || lane > 1) return PLUME_E_ARG;
    ScopedDevice sd(ctx->device);
    Lane& L = ctx->lanes[lane];
    CU(cudaStreamSynchronize(L.stream));
    size_t nd = L.d_io_cap < cap ? L.d_io_cap : cap, nh = L.h_cap < cap ? L.h_cap : cap;
    if (dev_copy && nd) CU(cudaMemcpy(dev_copy, L.d_io, nd, cudaMemcpyDeviceToHost));
    if (host_copy && nh) memcpy(host_copy, L.h_stage, nh);
    if (dev_bytes) *dev_bytes = dev_copy ? nd : 0;
    if (host_bytes) *host_bytes = host_copy ? nh : 0;
    return PLUME_OK;
}

// ---- device-pointer variants ---------------------------------------------------------------------------------
}  // extern "C" (reopened below)

namespace {
const char* const kMultiDevice = "a multi-device context has no device-pointer entry points: use plume_ctx_sub(ctx, i)";

int sign_device(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                const uint8_t* pk_in, const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nullifier, uint8_t* c, uint8_t* s_out,
                uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status, void* stream) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return fail(ctx, PLUME_E_ARG, kMultiDevice);
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, false)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (n > ctx->chunk) return fail(ctx, PLUME_E_ARG, "n exceeds plume_ctx_chunk_items()");
    if (!sk || !r || !nullifier || !c || !s_out || !status) return fail(ctx, PLUME_E_ARG, "null array");
    if (flavour == PLUME_FLAVOUR_ARKWORKS ? !pk_in : !pk) return fail(ctx, PLUME_E_ARG, "null pk array");
    ScopedDevice sd(ctx->device);
    return dev_run(ctx, n, (cudaStream_t)stream, [&](size_t f, size_t k, uint32_t* ws, uint32_t* vbtab, cudaStream_t s) -> int {
        sign_args a{};
        a.version = version; a.flavour = flavour; a.n = (uint32_t)k;
        a.msgs.base = msgs_at(msgs, msg_offsets, msg_len, f); a.msgs.offs = offs_at(msg_offsets, f); a.msgs.fixed_len = (uint32_t)msg_len;
        a.sk = at(sk, f, 32); a.r = at(r, f, 32); a.pk = at(pk, f, 64); a.pk_in = at(pk_in, f, 64); a.nullifier = at(nullifier, f, 64);
        a.c = at(c, f, 32); a.s = at(s_out, f, 32); a.r_point = at(r_point, f, 64);
        a.hashed_to_curve_r = at(hashed_to_curve_r, f, 64); a.status = at(status, f, 1);
        a.ws = ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = vbtab;
        return enqueue_sign(ctx, a, s);
    });
}
int verify_device(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                  const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s_in, const uint8_t* r_point,
                  const uint8_t* hashed_to_curve_r, uint8_t* ok, void* stream) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return fail(ctx, PLUME_E_ARG, kMultiDevice);
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, false)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (n > ctx->chunk) return fail(ctx, PLUME_E_ARG, "n exceeds plume_ctx_chunk_items()");
    if (!pk || !nullifier || !c || !s_in || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    if ((version == 1 || flavour == PLUME_FLAVOUR_ARKWORKS) && (!r_point || !hashed_to_curve_r))
        return fail(ctx, PLUME_E_ARG, "r_point and hashed_to_curve_r are required");
    ScopedDevice sd(ctx->device);
    return dev_run(ctx, n, (cudaStream_t)stream, [&](size_t f, size_t k, uint32_t* ws, uint32_t* vbtab, cudaStream_t s) -> int {
        verify_args a{};
        a.version = version; a.flavour = flavour; a.n = (uint32_t)k;
        a.msgs.base = msgs_at(msgs, msg_offsets, msg_len, f); a.msgs.offs = offs_at(msg_offsets, f); a.msgs.fixed_len = (uint32_t)msg_len;
        a.pk = at(pk, f, 64); a.nullifier = at(nullifier, f, 64); a.c = at(c, f, 32); a.s = at(s_in, f, 32);
        a.r_point = at(r_point, f, 64); a.hashed_to_curve_r = at(hashed_to_curve_r, f, 64);
        a.ok = at(ok, f, 1); a.ws = ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = vbtab;
        return enqueue_verify(ctx, a, s);
    });
}
}  // namespace

extern "C" {

int plume_sign_batch_device(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                            size_t msg_len, const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nullifier, uint8_t* c,
                            uint8_t* s_out, uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status, void* stream) {
    return sign_device(ctx, PLUME_FLAVOUR_K256, version, n, msgs, msg_offsets, msg_len, nullptr, sk, r, pk, nullifier, c, s_out, r_point,
                       hashed_to_curve_r, status, stream);
}
int plume_ark_sign_batch_device(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                                size_t msg_len, const uint8_t* pk, const uint8_t* sk, const uint8_t* r, uint8_t* nullifier,
                                uint8_t* digest_private, uint8_t* s_out, uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status,
                                void* stream) {
    return sign_device(ctx, PLUME_FLAVOUR_ARKWORKS, version, n, msgs, msg_offsets, msg_len, pk, sk, r, nullptr, nullifier,
                       digest_private, s_out, r_point, hashed_to_curve_r, status, stream);
}
int plume_verify_batch_device(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                              size_t msg_len, const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c,
                              const uint8_t* s_in, const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok,
                              void* stream) {
    return verify_device(ctx, PLUME_FLAVOUR_K256, version, n, msgs, msg_offsets, msg_len, pk, nullifier, c, s_in, r_point,
                         hashed_to_curve_r, ok, stream);
}
int plume_ark_verify_batch_device(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                                  size_t msg_len, const uint8_t* pk, const uint8_t* nullifier, const uint8_t* digest_private,
                                  const uint8_t* s_in, const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok,
                                  void* stream) {
    return verify_device(ctx, PLUME_FLAVOUR_ARKWORKS, version, n, msgs, msg_offsets, msg_len, pk, nullifier, digest_private, s_in,
                         r_point, hashed_to_curve_r, ok, stream);
}

int plume_hash_to_curve_batch_device(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                                     size_t msg_len, uint8_t* out, void* stream) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return fail(ctx, PLUME_E_ARG, kMultiDevice);
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, false)) return rc;
    if (n == 0) return PLUME_OK;
    if (n > ctx->chunk) return fail(ctx, PLUME_E_ARG, "n exceeds plume_ctx_chunk_items()");
    if (!out) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    return dev_run(ctx, n, (cudaStream_t)stream, [&](size_t f, size_t k, uint32_t* ws, uint32_t*, cudaStream_t s) -> int {
        h2c_args a{};
        a.n = (uint32_t)k;
        a.msgs.base = msgs_at(msgs, msg_offsets, msg_len, f); a.msgs.offs = offs_at(msg_offsets, f); a.msgs.fixed_len = (uint32_t)msg_len;
        a.out = at(out, f, 64); a.ws = ws;
        return enqueue_h2c(ctx, a, s);
    });
}

// ---- host-pointer variants: chunked, two lanes in flight --------------------------------------------------------
}  // extern "C" (reopened below)

namespace {

int sign_host(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
              const uint8_t* pk_in, const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nullifier, uint8_t* c, uint8_t* s_out,
              uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return sign_host(sub, flavour, version, k, msgs_at(msgs, msg_offsets, msg_len, f), offs_at(msg_offsets, f), msg_len,
                             at(pk_in, f, 64), at(sk, f, 32), at(r, f, 32), at(pk, f, 64), at(nullifier, f, 64), at(c, f, 32),
                             at(s_out, f, 32), at(r_point, f, 64), at(hashed_to_curve_r, f, 64), at(status, f, 1));
        });
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, true)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!sk || !r || !nullifier || !c || !s_out || !status) return fail(ctx, PLUME_E_ARG, "null array");
    if (flavour == PLUME_FLAVOUR_ARKWORKS ? !pk_in : !pk) return fail(ctx, PLUME_E_ARG, "null pk array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 64 * 4 + 64 + 1) + 4096)) return rc;
        sign_args a{};
        a.version = version; a.flavour = flavour; a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        uint8_t *d_sk, *d_r, *d_pk = nullptr;
        if (int rc = lane_input(ctx, L, sk + i0 * 32, cn * 32, &d_sk, true)) return rc;
        if (int rc = lane_input(ctx, L, r + i0 * 32, cn * 32, &d_r, true)) return rc;
        if (pk_in) if (int rc = lane_input(ctx, L, pk_in + i0 * 64, cn * 64, &d_pk)) return rc;
        a.sk = d_sk; a.r = d_r; a.pk_in = d_pk;
        size_t o_pk = 0, o_nul, o_c, o_s, o_rp = 0, o_hr = 0, o_st;
        a.pk = pk ? lane_output(L, cn * 64, &o_pk) : nullptr;
        a.nullifier = lane_output(L, cn * 64, &o_nul);
        a.c = lane_output(L, cn * 32, &o_c);
        a.s = lane_output(L, cn * 32, &o_s);
        a.r_point = r_point ? lane_output(L, cn * 64, &o_rp) : nullptr;
        a.hashed_to_curve_r = hashed_to_curve_r ? lane_output(L, cn * 64, &o_hr) : nullptr;
        a.status = lane_output(L, cn, &o_st);
        a.ws = L.ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = L.vbtab;
        if (int rc = enqueue_sign(ctx, a, L.stream)) return rc;
        // the device copies of the secret keys and nonces are dead now (the reference zeroises its witness,
        // javascript/src/lib.rs:64-71,82): clear them before anything else reuses the arena
        CU(cudaMemsetAsync(d_sk, 0, cn * 32, L.stream));
        CU(cudaMemsetAsync(d_r, 0, cn * 32, L.stream));
        if (pk) if (int rc = lane_fetch(ctx, L, pk + i0 * 64, o_pk, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, nullifier + i0 * 64, o_nul, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, c + i0 * 32, o_c, cn * 32)) return rc;
        if (int rc = lane_fetch(ctx, L, s_out + i0 * 32, o_s, cn * 32)) return rc;
        if (r_point) if (int rc = lane_fetch(ctx, L, r_point + i0 * 64, o_rp, cn * 64)) return rc;
        if (hashed_to_curve_r) if (int rc = lane_fetch(ctx, L, hashed_to_curve_r + i0 * 64, o_hr, cn * 64)) return rc;
        return lane_fetch(ctx, L, status + i0, o_st, cn);
    });
}

int verify_host(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s_in, const uint8_t* r_point,
                const uint8_t* hashed_to_curve_r, uint8_t* ok) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return verify_host(sub, flavour, version, k, msgs_at(msgs, msg_offsets, msg_len, f), offs_at(msg_offsets, f), msg_len,
                               at(pk, f, 64), at(nullifier, f, 64), at(c, f, 32), at(s_in, f, 32), at(r_point, f, 64),
                               at(hashed_to_curve_r, f, 64), at(ok, f, 1));
        });
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, true)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!pk || !nullifier || !c || !s_in || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    const bool need_points = version == 1 || flavour == PLUME_FLAVOUR_ARKWORKS;
    if (need_points && (!r_point || !hashed_to_curve_r)) return fail(ctx, PLUME_E_ARG, "r_point and hashed_to_curve_r are required");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 * 4 + 64 + 1) + 4096)) return rc;
        verify_args a{};
        a.version = version; a.flavour = flavour; a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        uint8_t *d_pk, *d_nul, *d_c, *d_s, *d_rp = nullptr, *d_hr = nullptr;
        if (int rc = lane_input(ctx, L, pk + i0 * 64, cn * 64, &d_pk)) return rc;
        if (int rc = lane_input(ctx, L, nullifier + i0 * 64, cn * 64, &d_nul)) return rc;
        if (int rc = lane_input(ctx, L, c + i0 * 32, cn * 32, &d_c)) return rc;
        if (int rc = lane_input(ctx, L, s_in + i0 * 32, cn * 32, &d_s)) return rc;
        if (need_points) {
            if (int rc = lane_input(ctx, L, r_point + i0 * 64, cn * 64, &d_rp)) return rc;
            if (int rc = lane_input(ctx, L, hashed_to_curve_r + i0 * 64, cn * 64, &d_hr)) return rc;
        }
        a.pk = d_pk; a.nullifier = d_nul; a.c = d_c; a.s = d_s; a.r_point = d_rp; a.hashed_to_curve_r = d_hr;
        size_t o_ok;
        a.ok = lane_output(L, cn, &o_ok);
        a.ws = L.ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = L.vbtab;
        if (int rc = enqueue_verify(ctx, a, L.stream)) return rc;
        return lane_fetch(ctx, L, ok + i0, o_ok, cn);
    });
}
}  // namespace

extern "C" {

int plume_sign_batch(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                     size_t msg_len, const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nullifier, uint8_t* c,
                     uint8_t* s_out, uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status) {
    return sign_host(ctx, PLUME_FLAVOUR_K256, version, n, msgs, msg_offsets, msg_len, nullptr, sk, r, pk, nullifier, c, s_out, r_point,
                     hashed_to_curve_r, status);
}
int plume_ark_sign_batch(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                         const uint8_t* pk, const uint8_t* sk, const uint8_t* r, uint8_t* nullifier, uint8_t* digest_private,
                         uint8_t* s_out, uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status) {
    return sign_host(ctx, PLUME_FLAVOUR_ARKWORKS, version, n, msgs, msg_offsets, msg_len, pk, sk, r, nullptr, nullifier, digest_private,
                     s_out, r_point, hashed_to_curve_r, status);
}
int plume_verify_batch(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets,
                       size_t msg_len, const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s_in,
                       const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok) {
    return verify_host(ctx, PLUME_FLAVOUR_K256, version, n, msgs, msg_offsets, msg_len, pk, nullifier, c, s_in, r_point,
                       hashed_to_curve_r, ok);
}
int plume_ark_verify_batch(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                           const uint8_t* pk, const uint8_t* nullifier, const uint8_t* digest_private, const uint8_t* s_in,
                           const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok) {
    return verify_host(ctx, PLUME_FLAVOUR_ARKWORKS, version, n, msgs, msg_offsets, msg_len, pk, nullifier, digest_private, s_in,
                       r_point, hashed_to_curve_r, ok);
}

}  // extern "C" (reopened below)
namespace {
int h2c_host(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len, const uint8_t* pk33, uint8_t* out) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return h2c_host(sub, k, msgs_at(msgs, msg_offsets, msg_len, f), offs_at(msg_offsets, f), msg_len, at(pk33, f, 33), at(out, f, 64));
        });
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, true)) return rc;
    if (n == 0) return PLUME_OK;
    if (!out) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 33) + 4096)) return rc;
        h2c_args a{};
        a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        if (pk33) {
            uint8_t* d_pk;
            if (int rc = lane_input(ctx, L, pk33 + i0 * 33, cn * 33, &d_pk)) return rc;
            a.pk33 = d_pk;
        }
        size_t o_out;
        a.out = lane_output(L, cn * 64, &o_out);
        a.ws = L.ws;
        if (int rc = enqueue_h2c(ctx, a, L.stream)) return rc;
        return lane_fetch(ctx, L, out + i0 * 64, o_out, cn * 64);
    });
}
}  // namespace
extern "C" {

int plume_hash_to_curve_batch(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                              uint8_t* out) {
    return h2c_host(ctx, n, msgs, msg_offsets, msg_len, nullptr, out);
}
int plume_hash_to_curve_pk_batch(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                 const uint8_t* pk33, uint8_t* out) {
    if (ctx && !pk33 && n) return fail(ctx, PLUME_E_ARG, "pk33 is null");
    return h2c_host(ctx, n, msgs, msg_offsets, msg_len, pk33, out);
}

int plume_hash_to_curve_witness_batch(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                      uint8_t* u, uint8_t* q, uint8_t* gx1_square, uint8_t* h, uint8_t* hints) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return plume_hash_to_curve_witness_batch(sub, k, msgs_at(msgs, msg_offsets, msg_len, f), offs_at(msg_offsets, f), msg_len,
                                                     at(u, f, 64), at(q, f, 128), at(gx1_square, f, 2), at(h, f, 64), at(hints, f, 192));
        });
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, true)) return rc;
    if (n == 0) return PLUME_OK;
    if (!u || !q || !gx1_square || !h) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 128 + 2 + 64 + 192) + 8192)) return rc;
        h2cw_args a{};
        a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        size_t o_u, o_q, o_f, o_h, o_hint = 0;
        a.u = lane_output(L, cn * 64, &o_u);
        a.q = lane_output(L, cn * 128, &o_q);
        a.gx1_square = lane_output(L, cn * 2, &o_f);
        a.h = lane_output(L, cn * 64, &o_h);
        a.hints = hints ? lane_output(L, cn * 192, &o_hint) : nullptr;
        a.ws = L.ws;
        if (int rc = enqueue_h2cw(ctx, a, L.stream)) return rc;
        if (int rc = lane_fetch(ctx, L, u + i0 * 64, o_u, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, q + i0 * 128, o_q, cn * 128)) return rc;
        if (int rc = lane_fetch(ctx, L, gx1_square + i0 * 2, o_f, cn * 2)) return rc;
        if (hints) if (int rc = lane_fetch(ctx, L, hints + i0 * 192, o_hint, cn * 192)) return rc;
        return lane_fetch(ctx, L, h + i0 * 64, o_h, cn * 64);
    });
}

int plume_fixed_base_mul_batch(plume_ctx* ctx, size_t n, const uint8_t* scalars, uint8_t* out) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) { return plume_fixed_base_mul_batch(sub, k, at(scalars, f, 32), at(out, f, 64)); });
    if (n == 0) return PLUME_OK;
    if (!scalars || !out) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, cn * 96 + 4096)) return rc;
        fbmul_args a{};
        a.n = (uint32_t)cn;
        uint8_t* d_k;
        if (int rc = lane_input(ctx, L, scalars + i0 * 32, cn * 32, &d_k, true)) return rc;   // may be secret keys
        a.k = d_k;
        size_t o_out;
        a.out = lane_output(L, cn * 64, &o_out);
        a.ws = L.ws; a.gtab = ctx->gtab; a.gw = ctx->gw;
        cudaStream_t s = L.stream;
        RUN(ST_FIXED_MUL, launch_fbmul(0, a, s));
        if (int rc = binv(ctx, a.ws, a.n, a.n, s)) return rc;
        RUN(ST_FIXED_MUL, launch_fbmul(1, a, s));
        CU(cudaMemsetAsync(d_k, 0, cn * 32, L.stream));
        return lane_fetch(ctx, L, out + i0 * 64, o_out, cn * 64);
    });
}

int plume_registers_batch(plume_ctx* ctx, size_t n, const uint8_t* in32, uint64_t* out4) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) { return plume_registers_batch(sub, k, at(in32, f, 32), at(out4, f, 4)); });
    if (n == 0) return PLUME_OK;
    if (!in32 || !out4) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, ctx->host_chunk * 4, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_reserve(ctx, L, cn * 64 + 4096)) return rc;
        uint8_t* d_in;
        if (int rc = lane_input(ctx, L, in32 + i0 * 32, cn * 32, &d_in)) return rc;
        size_t o_out;
        uint8_t* d_out = lane_output(L, cn * 32, &o_out);
        cudaStream_t s = L.stream;
        RUN(ST_REGISTERS, launch_registers((uint32_t)cn, d_in, reinterpret_cast<uint64_t*>(d_out), s));
        return lane_fetch(ctx, L, reinterpret_cast<uint8_t*>(out4) + i0 * 32, o_out, cn * 32);
    });
}

int plume_points_compress_batch_device(plume_ctx* ctx, size_t n, const uint8_t* in64, uint8_t* out33, void* stream) {
    if (!ctx || !in64 || !out33) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return fail(ctx, PLUME_E_ARG, kMultiDevice);
    if (n == 0) return PLUME_OK;
    ScopedDevice sd(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    RUN(ST_SEC1_COMPRESS, launch_sec1_compress((uint32_t)n, in64, out33, s));
    return PLUME_OK;
}
int plume_points_decompress_batch_device(plume_ctx* ctx, size_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok, void* stream) {
    if (!ctx || !in33 || !out64 || !ok) return PLUME_E_ARG;
    if (ctx_is_multi(ctx)) return fail(ctx, PLUME_E_ARG, kMultiDevice);
    if (n == 0) return PLUME_OK;
    ScopedDevice sd(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    RUN(ST_SEC1_DECOMPRESS, launch_sec1_decompress((uint32_t)n, in33, out64, ok, s));
    return PLUME_OK;
}

// ---- SEC1-compressed wire form (33-byte slots) -------------------------------------------------------------------
int plume_points_compress_batch(plume_ctx* ctx, size_t n, const uint8_t* in64, uint8_t* out33) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) { return plume_points_compress_batch(sub, k, at(in64, f, 64), at(out33, f, 33)); });
    if (n == 0) return PLUME_OK;
    if (!in64 || !out33) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_reserve(ctx, L, cn * 97 + 4096)) return rc;
        uint8_t* d_in;
        if (int rc = lane_input(ctx, L, in64 + i0 * 64, cn * 64, &d_in)) return rc;
        size_t o_out;
        uint8_t* d_out = lane_output(L, cn * 33, &o_out);
        cudaStream_t s = L.stream;
        RUN(ST_SEC1_COMPRESS, launch_sec1_compress((uint32_t)cn, d_in, d_out, s));
        return lane_fetch(ctx, L, out33 + i0 * 33, o_out, cn * 33);
    });
}

int plume_points_decompress_batch(plume_ctx* ctx, size_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return plume_points_decompress_batch(sub, k, at(in33, f, 33), at(out64, f, 64), at(ok, f, 1));
        });
    if (n == 0) return PLUME_OK;
    if (!in33 || !out64 || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_reserve(ctx, L, cn * 98 + 4096)) return rc;
        uint8_t* d_in;
        if (int rc = lane_input(ctx, L, in33 + i0 * 33, cn * 33, &d_in)) return rc;
        size_t o_out, o_ok;
        uint8_t* d_out = lane_output(L, cn * 64, &o_out);
        uint8_t* d_ok = lane_output(L, cn, &o_ok);
        cudaStream_t s = L.stream;
        RUN(ST_SEC1_DECOMPRESS, launch_sec1_decompress((uint32_t)cn, d_in, d_out, d_ok, s));
        if (int rc = lane_fetch(ctx, L, out64 + i0 * 64, o_out, cn * 64)) return rc;
        return lane_fetch(ctx, L, ok + i0, o_ok, cn);
    });
}

int plume_sign_batch_sec1(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                          const uint8_t* sk, const uint8_t* r, uint8_t* pk33, uint8_t* nullifier33, uint8_t* c, uint8_t* s_out,
                          uint8_t* r_point33, uint8_t* hashed_to_curve_r33, uint8_t* status) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return plume_sign_batch_sec1(sub, version, k, msgs_at(msgs, msg_offsets, msg_len, f), offs_at(msg_offsets, f), msg_len,
                                         at(sk, f, 32), at(r, f, 32), at(pk33, f, 33), at(nullifier33, f, 33), at(c, f, 32), at(s_out, f, 32),
                                         at(r_point33, f, 33), at(hashed_to_curve_r33, f, 33), at(status, f, 1));
        });
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, true)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!sk || !r || !pk33 || !nullifier33 || !c || !s_out || !status) return fail(ctx, PLUME_E_ARG, "null array");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 64 * 4 + 33 * 4 + 64 + 1) + 8192)) return rc;
        sign_args a{};
        a.version = version; a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        uint8_t *d_sk, *d_r;
        if (int rc = lane_input(ctx, L, sk + i0 * 32, cn * 32, &d_sk, true)) return rc;
        if (int rc = lane_input(ctx, L, r + i0 * 32, cn * 32, &d_r, true)) return rc;
        a.sk = d_sk; a.r = d_r;
        size_t o_tmp, o_c, o_s, o_st, o33[4] = {0, 0, 0, 0};
        a.pk = lane_output(L, cn * 64, &o_tmp);
        a.nullifier = lane_output(L, cn * 64, &o_tmp);
        a.c = lane_output(L, cn * 32, &o_c);
        a.s = lane_output(L, cn * 32, &o_s);
        a.r_point = lane_output(L, cn * 64, &o_tmp);
        a.hashed_to_curve_r = lane_output(L, cn * 64, &o_tmp);
        a.status = lane_output(L, cn, &o_st);
        a.ws = L.ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = L.vbtab;
        if (int rc = enqueue_sign(ctx, a, L.stream)) return rc;
        CU(cudaMemsetAsync(d_sk, 0, cn * 32, L.stream));
        CU(cudaMemsetAsync(d_r, 0, cn * 32, L.stream));
        cudaStream_t s = L.stream;
        uint8_t* src[4] = {a.pk, a.nullifier, a.r_point, a.hashed_to_curve_r};
        uint8_t* dst[4] = {pk33, nullifier33, r_point33, hashed_to_curve_r33};
        for (int q = 0; q < 4; q++) {
            if (!dst[q]) continue;
            uint8_t* d33 = lane_output(L, cn * 33, &o33[q]);
            RUN(ST_SEC1_COMPRESS, launch_sec1_compress((uint32_t)cn, src[q], d33, s));
            if (int rc = lane_fetch(ctx, L, dst[q] + i0 * 33, o33[q], cn * 33)) return rc;
        }
        if (int rc = lane_fetch(ctx, L, c + i0 * 32, o_c, cn * 32)) return rc;
        if (int rc = lane_fetch(ctx, L, s_out + i0 * 32, o_s, cn * 32)) return rc;
        return lane_fetch(ctx, L, status + i0, o_st, cn);
    });
}

int plume_verify_batch_sec1(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                            const uint8_t* pk33, const uint8_t* nullifier33, const uint8_t* c, const uint8_t* s_in,
                            const uint8_t* r_point33, const uint8_t* hashed_to_curve_r33, uint8_t* ok) {
    if (!ctx) return PLUME_E_ARG;
    if (ctx_is_multi(ctx))
        return multi_split(ctx, n, [&](plume_ctx* sub, size_t f, size_t k) {
            return plume_verify_batch_sec1(sub, version, k, msgs_at(msgs, msg_offsets, msg_len, f), offs_at(msg_offsets, f), msg_len,
                                           at(pk33, f, 33), at(nullifier33, f, 33), at(c, f, 32), at(s_in, f, 32), at(r_point33, f, 33),
                                           at(hashed_to_curve_r33, f, 33), at(ok, f, 1));
        });
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len, true)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!pk33 || !nullifier33 || !c || !s_in || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    if (version == 1 && (!r_point33 || !hashed_to_curve_r33)) return fail(ctx, PLUME_E_ARG, "V1 needs r_point and hashed_to_curve_r");
    return run_chunks(ctx, n, 0, [&](Lane& L, size_t i0, size_t cn) -> int {
        if (int rc = lane_workspace(ctx, L, cn)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (33 * 4 + 64 * 4 + 64 + 5) + 8192)) return rc;
        verify_args a{};
        a.version = version; a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        cudaStream_t s = L.stream;
        const uint8_t* src[4] = {pk33, nullifier33, version == 1 ? r_point33 : nullptr, version == 1 ? hashed_to_curve_r33 : nullptr};
        uint8_t* pts[4] = {nullptr, nullptr, nullptr, nullptr};
        uint8_t* flg[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int q = 0; q < 4; q++) {
            if (!src[q]) continue;
            uint8_t* d33;
            if (int rc = lane_input(ctx, L, src[q] + i0 * 33, cn * 33, &d33)) return rc;
            size_t o_tmp;
            pts[q] = lane_output(L, cn * 64, &o_tmp);
            flg[q] = lane_output(L, cn, &o_tmp);
            RUN(ST_SEC1_DECOMPRESS, launch_sec1_decompress((uint32_t)cn, d33, pts[q], flg[q], s));
        }
        uint8_t *d_c, *d_s;
        if (int rc = lane_input(ctx, L, c + i0 * 32, cn * 32, &d_c)) return rc;
        if (int rc = lane_input(ctx, L, s_in + i0 * 32, cn * 32, &d_s)) return rc;
        a.pk = pts[0]; a.nullifier = pts[1]; a.c = d_c; a.s = d_s; a.r_point = pts[2]; a.hashed_to_curve_r = pts[3];
        size_t o_ok;
        a.ok = lane_output(L, cn, &o_ok);
        a.ws = L.ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = L.vbtab;
        if (int rc = enqueue_verify(ctx, a, L.stream)) return rc;
        CU(launch_and_flags((uint32_t)cn, a.ok, flg[0], flg[1], flg[2], flg[3], s));
        ctx->launches++;
        return lane_fetch(ctx, L, ok + i0, o_ok, cn);
    });
}

}  // extern "C"

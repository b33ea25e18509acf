// api.cu -- the C ABI of include/plume_b200.h: context, chunked double-buffered execution of the
// stage pipelines, host staging, per-stage event timing.  No CPU fallback lives here: without a
// CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/plume_b200.h"
#include "launch.h"

namespace {

const char* const kStageNames[] = {"sign_fixed", "sign_h2c", "sign_varbase", "sign_final", "verify_h2c",
                                   "verify_muls", "verify_final", "h2c_map", "h2c_out", "binv", "sec1_compress", "sec1_decompress",
                                   "verify_mul_a", "verify_mul_b", "h2c_witness", "registers", "verify_tab_b"};
enum Stage { ST_SIGN_FIXED, ST_SIGN_H2C, ST_SIGN_VARBASE, ST_SIGN_FINAL, ST_VERIFY_H2C, ST_VERIFY_MULS,
             ST_VERIFY_FINAL, ST_H2C_MAP, ST_H2C_OUT, ST_BINV, ST_SEC1_COMPRESS, ST_SEC1_DECOMPRESS, ST_VERIFY_MUL_A,
             ST_VERIFY_MUL_B, ST_H2C_WITNESS, ST_REGISTERS, ST_VERIFY_TAB_B, ST_COUNT };

struct PendingCopy { void* dst; const void* src; size_t bytes; };

struct Lane {
    cudaStream_t stream = nullptr;
    uint32_t* ws = nullptr;          // WS_SLOTS * chunk * 32 bytes
    uint32_t* vbtab = nullptr;       // chunk * 1 KiB of window-table scratch in HBM/L2 (verify: two tables per item)
    uint8_t* d_io = nullptr;         // device arena for inputs and outputs of one chunk
    size_t d_io_cap = 0, d_io_used = 0;
    uint8_t* h_stage = nullptr;      // pinned staging arena (same layout as d_io)
    size_t h_cap = 0;
    std::vector<PendingCopy> pending;  // staged outputs to hand to the caller after the stream drains
    bool busy = false;
};

std::string g_create_error;

}  // namespace

struct plume_ctx {
    int device = 0;
    int gw = 0;
    uint32_t* gtab = nullptr;
    size_t chunk = 0;        // capacity of the workspaces: largest n of one pass
    size_t host_chunk = 0;   // pipelining granularity of the host-pointer entry points
    uint32_t binv_k = 16;
    Lane lanes[2];
    std::string err;
    uint64_t launches = 0;
    bool profiling = false;
    struct Ev { int stage; cudaEvent_t a, b; };
    std::vector<Ev> events;
    double stage_ms[ST_COUNT] = {0};
    uint64_t stage_n[ST_COUNT] = {0};
};

namespace {

int fail(plume_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, PLUME_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
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

// one stage launch, counted and (optionally) bracketed by an event pair on the launching stream
template <class F>
int run_stage(plume_ctx* ctx, int stage, cudaStream_t s, F&& launch) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (ctx->profiling) {
        CU(cudaEventCreate(&a));
        CU(cudaEventCreate(&b));
        CU(cudaEventRecord(a, s));
    }
    CU(launch());
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

int binv(plume_ctx* ctx, uint32_t* ws, uint32_t n, uint32_t m, cudaStream_t s) {
    RUN(ST_BINV, launch_binv(ws + (size_t)WS_Z0 * n * 8, ws + (size_t)WS_P0 * n * 8, m, ctx->binv_k, s));
    return PLUME_OK;
}

int enqueue_sign(plume_ctx* ctx, sign_args a, cudaStream_t s) {
    RUN(ST_SIGN_FIXED, launch_sign_fixed(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_SIGN_H2C, launch_sign_h2c(a, s));
    if (int rc = binv(ctx, a.ws, a.n, a.n, s)) return rc;
    RUN(ST_SIGN_VARBASE, launch_sign_varbase(a, s));
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_SIGN_FINAL, launch_sign_final(a, s));
    return PLUME_OK;
}
int enqueue_verify(plume_ctx* ctx, verify_args a, cudaStream_t s) {
    RUN(ST_VERIFY_H2C, launch_verify_h2c(a, s));
    if (int rc = binv(ctx, a.ws, a.n, a.n, s)) return rc;
#ifdef PLUME_VERIFY_FUSED
    RUN(ST_VERIFY_MULS, launch_verify_muls(a, s));
#else
    // separate kernels, each with its own register budget: the fused one needs 168 registers (12 warps/SM), the two
    // ladders alone run at 128 (16 warps/SM); 18 % faster in total
#ifdef PLUME_VERIFY_B_ONE
    RUN(ST_VERIFY_MUL_B, launch_verify_mul_b(a, s));
#else
    RUN(ST_VERIFY_TAB_B, launch_verify_tab_b(a, s));
    RUN(ST_VERIFY_MUL_B, launch_verify_lad_b(a, s));
#endif
    RUN(ST_VERIFY_MUL_A, launch_verify_mul_a(a, s));
#endif
    if (int rc = binv(ctx, a.ws, a.n, 2 * a.n, s)) return rc;
    RUN(ST_VERIFY_FINAL, launch_verify_final(a, s));
    return PLUME_OK;
}
int enqueue_h2c(plume_ctx* ctx, h2c_args a, cudaStream_t s) {
    RUN(ST_H2C_MAP, launch_h2c_map(a, s));
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

// ---- lane arenas ---------------------------------------------------------------------------------------
int lane_reserve(plume_ctx* ctx, Lane& L, size_t bytes) {
    if (bytes <= L.d_io_cap) return PLUME_OK;
    size_t cap = bytes + bytes / 4;
    if (L.d_io) { cudaFree(L.d_io); L.d_io = nullptr; }
    if (L.h_stage) { cudaFreeHost(L.h_stage); L.h_stage = nullptr; }
    L.d_io_cap = L.h_cap = 0;
    CU(cudaMalloc(&L.d_io, cap));
    CU(cudaHostAlloc(&L.h_stage, cap, cudaHostAllocDefault));
    L.d_io_cap = L.h_cap = cap;
    return PLUME_OK;
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// carve `bytes` out of the lane arena; returns the offset
size_t lane_take(Lane& L, size_t bytes) {
    size_t off = L.d_io_used;
    L.d_io_used = align256(off + bytes);
    return off;
}
// device copy of a host input array
int lane_input(plume_ctx* ctx, Lane& L, const void* host, size_t bytes, uint8_t** dev) {
    size_t off = lane_take(L, bytes);
    *dev = L.d_io + off;
    if (bytes == 0) return PLUME_OK;
    const void* src = host;
    if (!is_pinned(host)) {
        memcpy(L.h_stage + off, host, bytes);
        src = L.h_stage + off;
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
        CU(cudaMemcpyAsync(L.h_stage + off, L.d_io + off, bytes, cudaMemcpyDeviceToHost, L.stream));
        L.pending.push_back({host, L.h_stage + off, bytes});
    }
    return PLUME_OK;
}
int lane_finish(plume_ctx* ctx, Lane& L) {
    if (!L.busy) return PLUME_OK;
    CU(cudaStreamSynchronize(L.stream));
    for (const PendingCopy& p : L.pending) memcpy(p.dst, p.src, p.bytes);
    L.pending.clear();
    L.busy = false;
    return PLUME_OK;
}

struct MsgChunk { const uint8_t* base; size_t bytes; const uint64_t* offs; uint64_t first; };

// stage the messages of items [i0, i0+cn): returns device views
int lane_msgs(plume_ctx* ctx, Lane& L, const uint8_t* msgs, const uint64_t* offs, size_t msg_len, size_t i0, size_t cn,
              msg_view* view) {
    uint8_t* d_msgs = nullptr;
    if (offs) {
        uint64_t b0 = offs[i0], b1 = offs[i0 + cn];
        if (int rc = lane_input(ctx, L, msgs + b0, (size_t)(b1 - b0), &d_msgs)) return rc;
        // rebased offsets
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

int check_common(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* offs, size_t msg_len) {
    if (!ctx) return PLUME_E_ARG;
    if (n > 0 && !msgs && (offs ? offs[n] != offs[0] : msg_len != 0)) return fail(ctx, PLUME_E_ARG, "msgs is null");
    if (!offs && msg_len > 0xFFFFFFFFull) return fail(ctx, PLUME_E_ARG, "msg_len too large");
    return PLUME_OK;
}

}  // namespace

extern "C" {

int plume_version(void) { return PLUME_ABI_VERSION; }

const char* plume_last_error(const plume_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

size_t plume_ctx_chunk_items(const plume_ctx* ctx) { return ctx ? ctx->chunk : 0; }
uint64_t plume_ctx_launch_count(const plume_ctx* ctx) { return ctx ? ctx->launches : 0; }

void plume_ctx_destroy(plume_ctx* ctx) {
    if (!ctx) return;
    ScopedDevice sd(ctx->device);
    cudaDeviceSynchronize();
    for (auto& e : ctx->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (Lane& L : ctx->lanes) {
        if (L.ws) cudaFree(L.ws);
        if (L.vbtab) cudaFree(L.vbtab);
        if (L.d_io) cudaFree(L.d_io);
        if (L.h_stage) cudaFreeHost(L.h_stage);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    if (ctx->gtab) cudaFree(ctx->gtab);
    delete ctx;
}

int plume_ctx_create(plume_ctx** out, int device, int fixed_window_bits) {
    plume_ctx* ctx = nullptr;  // CU() reports into g_create_error while ctx is null
    if (!out) return fail(nullptr, PLUME_E_ARG, "out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PLUME_E_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, PLUME_E_NO_DEVICE, "device ordinal out of range");
    // default 20 bits: 13 windows x 2^20 affine points = 872 MB of HBM, 13 additions per fixed-base multiplication
    // (16 bits: 64 MiB, 16 additions; measured sign_fixed 3.44 -> 2.79 ms and verify_mul_a 19.46 -> 19.03 ms per 2^20 items)
    int w = fixed_window_bits ? fixed_window_bits : (int)env_size("PLUME_FIXED_WINDOW", 20);
    if (w < 4 || w > 22) return fail(nullptr, PLUME_E_ARG, "fixed_window_bits must be in 4..22");
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
    // = 4 full waves of the one that holds 3 (k_verify_mul_b), i.e. no tail, and short enough that the first upload and
    // the last download of a call -- the part the two lanes cannot overlap -- are a small fraction of it
    c->host_chunk = env_size("PLUME_HOST_CHUNK_ITEMS", (size_t)prop.multiProcessorCount * 128 * 12);
    if (c->host_chunk > c->chunk) c->host_chunk = c->chunk;
    c->binv_k = (uint32_t)env_size("PLUME_BINV_K", 16);
    struct Guard { plume_ctx* c; ~Guard() { if (c) plume_ctx_destroy(c); } } guard{c};
    for (Lane& L : c->lanes) {
        CU(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&L.ws, (size_t)WS_SLOTS * c->chunk * 32));
        CU(cudaMalloc(&L.vbtab, c->chunk * (size_t)VB_TAB_WORDS * 4 * 2));  // two window tables per item (verify)
    }
    // generator table: entries -> batched inversion -> affine
    const int nwin = (256 + w - 1) / w;
    const size_t ne = (size_t)nwin << w;
    uint32_t *bases = nullptr, *zs = nullptr, *scratch = nullptr;
    CU(cudaMalloc(&c->gtab, ne * 64));
    CU(cudaMalloc(&bases, (size_t)nwin * 64));
    CU(cudaMalloc(&zs, ne * 32));
    CU(cudaMalloc(&scratch, ne * 32));
    cudaStream_t s = c->lanes[0].stream;
    CU(launch_gtab_bases(bases, w, s));
    CU(launch_gtab_entries((uint32_t)ne, c->gtab, zs, bases, w, s));
    CU(launch_binv(zs, scratch, (uint32_t)ne, 16, s));
    CU(launch_gtab_norm((uint32_t)ne, c->gtab, zs, s));
    CU(cudaStreamSynchronize(s));
    cudaFree(bases); cudaFree(zs); cudaFree(scratch);
    guard.c = nullptr;
    *out = c;
    return PLUME_OK;
}

int plume_ctx_set_profiling(plume_ctx* ctx, int on) {
    if (!ctx) return PLUME_E_ARG;
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
    ScopedDevice sd(ctx->device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    uint32_t* sink = nullptr;
    CU(cudaMalloc(&sink, 64));
    cudaStream_t s = ctx->lanes[0].stream;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double best[2] = {0, 0};
    for (int form = 0; form < 2; form++) {
        CU(launch_imad_peak(sink, iters, blocks, threads, form, s));  // warm-up
        for (int rep = 0; rep < 5; rep++) {
            CU(cudaEventRecord(a, s));
            CU(launch_imad_peak(sink, iters, blocks, threads, form, s));
            CU(cudaEventRecord(b, s));
            CU(cudaEventSynchronize(b));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, a, b));
            double rate = (double)blocks * threads * (double)iters * 64.0 / (ms * 1e-3);
            if (rate > best[form]) best[form] = rate;
        }
    }
    ctx->launches += 12;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
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
    if (n == 0) return PLUME_OK;
    ScopedDevice sd(ctx->device);
    uint32_t *da = nullptr, *db = nullptr, *dout = nullptr;
    cudaStream_t s = ctx->lanes[0].stream;
    CU(cudaMalloc(&da, n * 32));
    CU(cudaMalloc(&db, n * 32));
    CU(cudaMalloc(&dout, n * 32));
    CU(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, s));
    CU(launch_debug_fe_op(op, (uint32_t)n, da, db, dout, s));
    ctx->launches++;
    CU(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return PLUME_OK;
}

// ---- device-pointer variants ---------------------------------------------------------------------------------
}  // extern "C" (reopened below)

namespace {
int sign_device(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                const uint8_t* pk_in, const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nullifier, uint8_t* c, uint8_t* s_out,
                uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status, void* stream) {
    if (!ctx) return PLUME_E_ARG;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (n > ctx->chunk) return fail(ctx, PLUME_E_ARG, "n exceeds plume_ctx_chunk_items()");
    if (!sk || !r || !nullifier || !c || !s_out || !status) return fail(ctx, PLUME_E_ARG, "null array");
    if (flavour == PLUME_FLAVOUR_ARKWORKS ? !pk_in : !pk) return fail(ctx, PLUME_E_ARG, "null pk array");
    ScopedDevice sd(ctx->device);
    sign_args a{};
    a.version = version; a.flavour = flavour; a.n = (uint32_t)n;
    a.msgs.base = msgs; a.msgs.offs = msg_offsets; a.msgs.fixed_len = (uint32_t)msg_len;
    a.sk = sk; a.r = r; a.pk = pk; a.pk_in = pk_in; a.nullifier = nullifier; a.c = c; a.s = s_out; a.r_point = r_point;
    a.hashed_to_curve_r = hashed_to_curve_r; a.status = status;
    a.ws = ctx->lanes[0].ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = ctx->lanes[0].vbtab;
    return enqueue_sign(ctx, a, (cudaStream_t)stream);
}
int verify_device(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                  const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s_in, const uint8_t* r_point,
                  const uint8_t* hashed_to_curve_r, uint8_t* ok, void* stream) {
    if (!ctx) return PLUME_E_ARG;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (n > ctx->chunk) return fail(ctx, PLUME_E_ARG, "n exceeds plume_ctx_chunk_items()");
    if (!pk || !nullifier || !c || !s_in || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    if ((version == 1 || flavour == PLUME_FLAVOUR_ARKWORKS) && (!r_point || !hashed_to_curve_r))
        return fail(ctx, PLUME_E_ARG, "r_point and hashed_to_curve_r are required");
    ScopedDevice sd(ctx->device);
    verify_args a{};
    a.version = version; a.flavour = flavour; a.n = (uint32_t)n;
    a.msgs.base = msgs; a.msgs.offs = msg_offsets; a.msgs.fixed_len = (uint32_t)msg_len;
    a.pk = pk; a.nullifier = nullifier; a.c = c; a.s = s_in; a.r_point = r_point; a.hashed_to_curve_r = hashed_to_curve_r;
    a.ok = ok; a.ws = ctx->lanes[0].ws; a.gtab = ctx->gtab; a.gw = ctx->gw; a.vbtab = ctx->lanes[0].vbtab;
    return enqueue_verify(ctx, a, (cudaStream_t)stream);
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
    if (n == 0) return PLUME_OK;
    if (n > ctx->chunk) return fail(ctx, PLUME_E_ARG, "n exceeds plume_ctx_chunk_items()");
    if (!out) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    h2c_args a;
    a.n = (uint32_t)n;
    a.msgs.base = msgs; a.msgs.offs = msg_offsets; a.msgs.fixed_len = (uint32_t)msg_len;
    a.out = out; a.ws = ctx->lanes[0].ws;
    return enqueue_h2c(ctx, a, (cudaStream_t)stream);
}

// ---- host-pointer variants: chunked, two lanes in flight --------------------------------------------------------
}  // extern "C" (reopened below)

namespace {
int sign_host(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
              const uint8_t* pk_in, const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nullifier, uint8_t* c, uint8_t* s_out,
              uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status) {
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!sk || !r || !nullifier || !c || !s_out || !status) return fail(ctx, PLUME_E_ARG, "null array");
    if (flavour == PLUME_FLAVOUR_ARKWORKS ? !pk_in : !pk) return fail(ctx, PLUME_E_ARG, "null pk array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 64 * 4 + 64 + 1) + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        sign_args a{};
        a.version = version; a.flavour = flavour; a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        uint8_t *d_sk, *d_r, *d_pk = nullptr;
        if (int rc = lane_input(ctx, L, sk + i0 * 32, cn * 32, &d_sk)) return rc;
        if (int rc = lane_input(ctx, L, r + i0 * 32, cn * 32, &d_r)) return rc;
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
        if (pk) if (int rc = lane_fetch(ctx, L, pk + i0 * 64, o_pk, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, nullifier + i0 * 64, o_nul, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, c + i0 * 32, o_c, cn * 32)) return rc;
        if (int rc = lane_fetch(ctx, L, s_out + i0 * 32, o_s, cn * 32)) return rc;
        if (r_point) if (int rc = lane_fetch(ctx, L, r_point + i0 * 64, o_rp, cn * 64)) return rc;
        if (hashed_to_curve_r) if (int rc = lane_fetch(ctx, L, hashed_to_curve_r + i0 * 64, o_hr, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, status + i0, o_st, cn)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int verify_host(plume_ctx* ctx, int flavour, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s_in, const uint8_t* r_point,
                const uint8_t* hashed_to_curve_r, uint8_t* ok) {
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!pk || !nullifier || !c || !s_in || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    const bool need_points = version == 1 || flavour == PLUME_FLAVOUR_ARKWORKS;
    if (need_points && (!r_point || !hashed_to_curve_r)) return fail(ctx, PLUME_E_ARG, "r_point and hashed_to_curve_r are required");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 * 4 + 64 + 1) + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
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
        if (int rc = lane_fetch(ctx, L, ok + i0, o_ok, cn)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
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

int plume_hash_to_curve_batch(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                              uint8_t* out) {
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len)) return rc;
    if (n == 0) return PLUME_OK;
    if (!out) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * 64 + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        h2c_args a;
        a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        size_t o_out;
        a.out = lane_output(L, cn * 64, &o_out);
        a.ws = L.ws;
        if (int rc = enqueue_h2c(ctx, a, L.stream)) return rc;
        if (int rc = lane_fetch(ctx, L, out + i0 * 64, o_out, cn * 64)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int plume_hash_to_curve_witness_batch(plume_ctx* ctx, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                      uint8_t* u, uint8_t* q, uint8_t* gx1_square, uint8_t* h) {
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len)) return rc;
    if (n == 0) return PLUME_OK;
    if (!u || !q || !gx1_square || !h) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 128 + 2 + 64) + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        h2cw_args a{};
        a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        size_t o_u, o_q, o_f, o_h;
        a.u = lane_output(L, cn * 64, &o_u);
        a.q = lane_output(L, cn * 128, &o_q);
        a.gx1_square = lane_output(L, cn * 2, &o_f);
        a.h = lane_output(L, cn * 64, &o_h);
        a.ws = L.ws;
        if (int rc = enqueue_h2cw(ctx, a, L.stream)) return rc;
        if (int rc = lane_fetch(ctx, L, u + i0 * 64, o_u, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, q + i0 * 128, o_q, cn * 128)) return rc;
        if (int rc = lane_fetch(ctx, L, gx1_square + i0 * 2, o_f, cn * 2)) return rc;
        if (int rc = lane_fetch(ctx, L, h + i0 * 64, o_h, cn * 64)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int plume_registers_batch(plume_ctx* ctx, size_t n, const uint8_t* in32, uint64_t* out4) {
    if (!ctx) return PLUME_E_ARG;
    if (n == 0) return PLUME_OK;
    if (!in32 || !out4) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    const size_t step = ctx->host_chunk * 4;
    for (size_t i0 = 0; i0 < n; i0 += step, k++) {
        const size_t cn = (n - i0 < step) ? n - i0 : step;
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, cn * 64 + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        uint8_t* d_in;
        if (int rc = lane_input(ctx, L, in32 + i0 * 32, cn * 32, &d_in)) return rc;
        size_t o_out;
        uint8_t* d_out = lane_output(L, cn * 32, &o_out);
        cudaStream_t s = L.stream;
        RUN(ST_REGISTERS, launch_registers((uint32_t)cn, d_in, reinterpret_cast<uint64_t*>(d_out), s));
        if (int rc = lane_fetch(ctx, L, reinterpret_cast<uint8_t*>(out4) + i0 * 32, o_out, cn * 32)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int plume_points_compress_batch_device(plume_ctx* ctx, size_t n, const uint8_t* in64, uint8_t* out33, void* stream) {
    if (!ctx || !in64 || !out33) return PLUME_E_ARG;
    if (n == 0) return PLUME_OK;
    ScopedDevice sd(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    RUN(ST_SEC1_COMPRESS, launch_sec1_compress((uint32_t)n, in64, out33, s));
    return PLUME_OK;
}
int plume_points_decompress_batch_device(plume_ctx* ctx, size_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok, void* stream) {
    if (!ctx || !in33 || !out64 || !ok) return PLUME_E_ARG;
    if (n == 0) return PLUME_OK;
    ScopedDevice sd(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    RUN(ST_SEC1_DECOMPRESS, launch_sec1_decompress((uint32_t)n, in33, out64, ok, s));
    return PLUME_OK;
}

// ---- SEC1-compressed wire form (33-byte slots) -------------------------------------------------------------------
int plume_points_compress_batch(plume_ctx* ctx, size_t n, const uint8_t* in64, uint8_t* out33) {
    if (!ctx) return PLUME_E_ARG;
    if (n == 0) return PLUME_OK;
    if (!in64 || !out33) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, cn * 97 + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        uint8_t* d_in;
        if (int rc = lane_input(ctx, L, in64 + i0 * 64, cn * 64, &d_in)) return rc;
        size_t o_out;
        uint8_t* d_out = lane_output(L, cn * 33, &o_out);
        cudaStream_t s = L.stream;
        RUN(ST_SEC1_COMPRESS, launch_sec1_compress((uint32_t)cn, d_in, d_out, s));
        if (int rc = lane_fetch(ctx, L, out33 + i0 * 33, o_out, cn * 33)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int plume_points_decompress_batch(plume_ctx* ctx, size_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok) {
    if (!ctx) return PLUME_E_ARG;
    if (n == 0) return PLUME_OK;
    if (!in33 || !out64 || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, cn * 98 + 4096)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        uint8_t* d_in;
        if (int rc = lane_input(ctx, L, in33 + i0 * 33, cn * 33, &d_in)) return rc;
        size_t o_out, o_ok;
        uint8_t* d_out = lane_output(L, cn * 64, &o_out);
        uint8_t* d_ok = lane_output(L, cn, &o_ok);
        cudaStream_t s = L.stream;
        RUN(ST_SEC1_DECOMPRESS, launch_sec1_decompress((uint32_t)cn, d_in, d_out, d_ok, s));
        if (int rc = lane_fetch(ctx, L, out64 + i0 * 64, o_out, cn * 64)) return rc;
        if (int rc = lane_fetch(ctx, L, ok + i0, o_ok, cn)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int plume_sign_batch_sec1(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                          const uint8_t* sk, const uint8_t* r, uint8_t* pk33, uint8_t* nullifier33, uint8_t* c, uint8_t* s_out,
                          uint8_t* r_point33, uint8_t* hashed_to_curve_r33, uint8_t* status) {
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!sk || !r || !pk33 || !nullifier33 || !c || !s_out || !status) return fail(ctx, PLUME_E_ARG, "null array");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (64 + 64 * 4 + 33 * 4 + 64 + 1) + 8192)) return rc;
        L.d_io_used = 0;
        L.busy = true;
        sign_args a{};
        a.version = version; a.n = (uint32_t)cn;
        if (int rc = lane_msgs(ctx, L, msgs, msg_offsets, msg_len, i0, cn, &a.msgs)) return rc;
        uint8_t *d_sk, *d_r;
        if (int rc = lane_input(ctx, L, sk + i0 * 32, cn * 32, &d_sk)) return rc;
        if (int rc = lane_input(ctx, L, r + i0 * 32, cn * 32, &d_r)) return rc;
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
        if (int rc = lane_fetch(ctx, L, status + i0, o_st, cn)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

int plume_verify_batch_sec1(plume_ctx* ctx, int version, size_t n, const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                            const uint8_t* pk33, const uint8_t* nullifier33, const uint8_t* c, const uint8_t* s_in,
                            const uint8_t* r_point33, const uint8_t* hashed_to_curve_r33, uint8_t* ok) {
    if (int rc = check_common(ctx, n, msgs, msg_offsets, msg_len)) return rc;
    if (version != 1 && version != 2) return fail(ctx, PLUME_E_ARG, "version must be 1 or 2");
    if (n == 0) return PLUME_OK;
    if (!pk33 || !nullifier33 || !c || !s_in || !ok) return fail(ctx, PLUME_E_ARG, "null array");
    if (version == 1 && (!r_point33 || !hashed_to_curve_r33)) return fail(ctx, PLUME_E_ARG, "V1 needs r_point and hashed_to_curve_r");
    ScopedDevice sd(ctx->device);
    size_t k = 0;
    for (size_t i0 = 0, cn = 0; i0 < n; i0 += cn, k++) {
        cn = host_chunk_len(ctx, n, i0);
        Lane& L = ctx->lanes[k & 1];
        if (int rc = lane_finish(ctx, L)) return rc;
        if (int rc = lane_reserve(ctx, L, msgs_bytes(msg_offsets, msg_len, i0, cn) + cn * (33 * 4 + 64 * 4 + 64 + 5) + 8192)) return rc;
        L.d_io_used = 0;
        L.busy = true;
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
        if (int rc = lane_fetch(ctx, L, ok + i0, o_ok, cn)) return rc;
    }
    for (Lane& L : ctx->lanes) if (int rc = lane_finish(ctx, L)) return rc;
    return PLUME_OK;
}

}  // extern "C"

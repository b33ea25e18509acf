// api_multi.cu -- the multi-device context of include/plume_b200.h: ONE caller, ONE batch in host memory, G GPUs.
//
// PLUME items are independent, so a multi-device context is G ordinary single-device contexts plus one
// worker thread per device: a host-pointer call range-splits its batch, device g taking items [g n / G, (g+1) n / G),
// and every worker runs the single-device two-lane pipeline on its slice, writing straight into the caller's arrays.
// There is no exchange step on the data path.  The only thing the devices could share is the fixed-base table of the
// generator; by default each device builds its own copy concurrently (0.1 s), and PLUME_GTAB_BCAST selects the two
// broadcast forms north_star mentions so that they can be measured against that:
//     PLUME_GTAB_BCAST=p2p    device 0 builds, the others receive a peer copy over NVLink (cudaMemcpyPeerAsync)
//     PLUME_GTAB_BCAST=nccl   device 0 builds, ncclBroadcast inside one ncclCommInitAll communicator
//                             (libnccl.so.2 is looked up at run time -- PLUME_NCCL_LIB or the loader's search path; the
//                             library itself does not link against NCCL)
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include "ctx.h"

// one worker thread per device: runs the tasks posted to it in order
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = false, quit = false;
    int rc = 0;

    Worker() : th([this] { loop(); }) {}
    ~Worker() {
        { std::lock_guard<std::mutex> g(mu); quit = true; }
        cv.notify_all();
        th.join();
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [this] { return has_task || quit; });
            if (quit) return;
            std::function<int()> t = std::move(task);
            has_task = false;
            lk.unlock();
            int r = t();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
    void post(std::function<int()> t) {
        { std::lock_guard<std::mutex> g(mu); task = std::move(t); has_task = true; done = false; }
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return done; });
        return rc;
    }
};

// the range split of SURVEY.md 8e: part g of G owns items [g n / G, (g+1) n / G) -- sizes differ by at most one and the
// parts tile [0, n) in order.  Exported so that the rule can be checked without a GPU (tests/test_sharding_gloo.py) and
// reused by callers that shard across processes (bench.py's torchrun ranks use the same formula).
extern "C" int plume_shard_range(size_t n, int part, int parts, size_t* first, size_t* count) {
    if (parts < 1 || part < 0 || part >= parts || !first || !count) return PLUME_E_ARG;
    // n * part can exceed 64 bits only for n > 2^64 / parts; split the product to stay exact for any size_t n
    const size_t q = n / (size_t)parts, r = n % (size_t)parts;
    const size_t lo = q * (size_t)part + (r * (size_t)part) / (size_t)parts;
    const size_t hi = q * (size_t)(part + 1) + (r * (size_t)(part + 1)) / (size_t)parts;
    *first = lo;
    *count = hi - lo;
    return PLUME_OK;
}

int multi_split(plume_ctx* ctx, size_t n, const std::function<int(plume_ctx*, size_t, size_t)>& f) {
    const size_t G = ctx->subs.size();
    for (size_t g = 0; g < G; g++) {
        size_t first = 0, count = 0;
        plume_shard_range(n, (int)g, (int)G, &first, &count);
        const size_t last = first + count;
        plume_ctx* sub = ctx->subs[g];
        ctx->workers[g]->post([&f, sub, first, last]() -> int { return last > first ? f(sub, first, last - first) : PLUME_OK; });
    }
    int rc = PLUME_OK;
    for (size_t g = 0; g < G; g++) {
        int r = ctx->workers[g]->wait();
        if (r != PLUME_OK && rc == PLUME_OK) {
            rc = r;
            ctx->err = "device " + std::to_string(ctx->subs[g]->device) + ": " + ctx->subs[g]->err;
        }
    }
    return rc;
}

void multi_destroy(plume_ctx* ctx) {
    for (Worker* w : ctx->workers) delete w;
    for (plume_ctx* s : ctx->subs) plume_ctx_destroy(s);
    ctx->workers.clear();
    ctx->subs.clear();
    delete ctx;
}

namespace {

// ---- NCCL, resolved at run time ------------------------------------------------------------------------------
struct Nccl {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& why) {
        const char* names[] = {getenv("PLUME_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { why = "libnccl.so.2 not found (set PLUME_NCCL_LIB)"; return false; }
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Broadcast) { why = "libnccl lacks an expected symbol"; return false; }
        return true;
    }
};

// ncclBroadcast of device 0's finished table into the (allocated) tables of the other sub-contexts
int nccl_broadcast_gtab(std::vector<plume_ctx*>& subs, std::string& why) {
    Nccl nc;
    if (!nc.load(why)) return PLUME_E_CUDA;
    const int G = (int)subs.size();
    std::vector<void*> comms(G, nullptr);
    std::vector<int> devs(G);
    for (int g = 0; g < G; g++) devs[g] = subs[g]->device;
    int r = nc.CommInitAll(comms.data(), G, devs.data());
    if (r != 0) { why = std::string("ncclCommInitAll: ") + (nc.GetErrorString ? nc.GetErrorString(r) : "error"); return PLUME_E_CUDA; }
    const size_t bytes = ctx_gtab_bytes(subs[0]->gw);
    int rc = PLUME_OK;
    nc.GroupStart();
    for (int g = 0; g < G; g++) {
        cudaSetDevice(subs[g]->device);
        r = nc.Broadcast(subs[g]->gtab, subs[g]->gtab, bytes, /*ncclUint8*/ 1, 0, comms[g], subs[g]->lanes[0].stream);
        if (r != 0 && rc == PLUME_OK) { why = std::string("ncclBroadcast: ") + (nc.GetErrorString ? nc.GetErrorString(r) : "error"); rc = PLUME_E_CUDA; }
    }
    r = nc.GroupEnd();
    if (r != 0 && rc == PLUME_OK) { why = "ncclGroupEnd failed"; rc = PLUME_E_CUDA; }
    for (int g = 0; g < G; g++) {
        cudaSetDevice(subs[g]->device);
        if (cudaStreamSynchronize(subs[g]->lanes[0].stream) != cudaSuccess && rc == PLUME_OK) { why = "broadcast stream failed"; rc = PLUME_E_CUDA; }
    }
    for (int g = 0; g < G; g++) nc.CommDestroy(comms[g]);
    return rc;
}

}  // namespace

extern "C" int plume_ctx_create_multi(plume_ctx** out, const int* devices, int n_devices, int fixed_window_bits) {
    if (!out) return ctx_fail(nullptr, PLUME_E_ARG, "out is null");
    *out = nullptr;
    if (!devices || n_devices < 1) return ctx_fail(nullptr, PLUME_E_ARG, "devices[] is empty");
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return ctx_fail(nullptr, PLUME_E_ARG, "devices[] names a device twice");
    const char* mode = getenv("PLUME_GTAB_BCAST");
    const bool p2p = mode && strcmp(mode, "p2p") == 0, nccl = mode && strcmp(mode, "nccl") == 0;
    int prev = 0;
    cudaGetDevice(&prev);
    plume_ctx* c = new plume_ctx();
    c->device = -1;
    c->subs.assign(n_devices, nullptr);
    std::vector<int> rcs(n_devices, PLUME_OK);
    std::vector<std::string> errs(n_devices);
    auto make = [&](int g, const uint32_t* from, int from_dev) {
        rcs[g] = ctx_create_single(&c->subs[g], devices[g], fixed_window_bits, from, from_dev);
        if (rcs[g] != PLUME_OK) errs[g] = plume_last_error(nullptr);   // thread-local text of this worker
    };
    if (p2p || nccl) {
        // device 0 builds the table; the others allocate theirs and receive a copy
        make(0, nullptr, 0);
        if (rcs[0] == PLUME_OK) {
            if (p2p) {
                for (int g = 1; g < n_devices; g++) {   // peer access makes the copy a direct NVLink transfer
                    cudaSetDevice(devices[g]);
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, devices[g], devices[0]);
                    if (can && cudaDeviceEnablePeerAccess(devices[0], 0) != cudaSuccess) cudaGetLastError();
                }
                std::vector<std::thread> th;
                for (int g = 1; g < n_devices; g++) th.emplace_back(make, g, (const uint32_t*)c->subs[0]->gtab, devices[0]);
                for (auto& t : th) t.join();
            } else {
                // allocate without building: a peer copy from device 0 fills the table, NCCL then overwrites it with the
                // same bytes -- the broadcast is what gets timed by the caller (tests/gpu_multi.py)
                std::vector<std::thread> th;
                for (int g = 1; g < n_devices; g++) th.emplace_back(make, g, (const uint32_t*)c->subs[0]->gtab, devices[0]);
                for (auto& t : th) t.join();
                bool all = true;
                for (int g = 0; g < n_devices; g++) all = all && rcs[g] == PLUME_OK;
                if (all) {
                    std::string why;
                    int r = nccl_broadcast_gtab(c->subs, why);
                    if (r != PLUME_OK) { rcs[0] = r; errs[0] = why; }
                }
            }
        }
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < n_devices; g++) th.emplace_back(make, g, nullptr, 0);
        for (auto& t : th) t.join();
    }
    cudaSetDevice(prev);
    for (int g = 0; g < n_devices; g++) {
        if (rcs[g] != PLUME_OK) {
            int rc = rcs[g];
            std::string msg = "device " + std::to_string(devices[g]) + ": " + errs[g];
            for (plume_ctx* s : c->subs) if (s) plume_ctx_destroy(s);
            delete c;
            return ctx_fail(nullptr, rc, msg);
        }
    }
    for (int g = 0; g < n_devices; g++) c->workers.push_back(new Worker());
    c->chunk = c->subs[0]->chunk;
    c->host_chunk = c->subs[0]->host_chunk;
    c->gw = c->subs[0]->gw;
    *out = c;
    return PLUME_OK;
}

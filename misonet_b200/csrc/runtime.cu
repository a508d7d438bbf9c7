// Library-level plumbing of the C ABI: error string, launch counter, device check.
#include <mutex>
#include <vector>

#include <cstdlib>

#include "common.cuh"

namespace miso {

int pdl_level() {
    static const int lvl = getenv("MISO_PDL") ? atoi(getenv("MISO_PDL")) : 0;  // 1: whole chain, 2: operand-preparation kernels only
    return lvl;
}

bool pdl_enabled() {
    static const bool on = getenv("MISO_PDL") && atoi(getenv("MISO_PDL")) != 0;  // measured on B200: 11.84 ms/step with, 11.54 without -> off by default
    return on;
}


static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}


// ---- per-launch event profiler ---------------------------------------------------------
namespace {

std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_pool;
std::atomic<bool> g_prof_on{false};
thread_local cudaEvent_t g_open = nullptr;
thread_local std::vector<ProfRec> *g_sink = nullptr;  // set while a forward is being captured into a CUDA graph

cudaEvent_t get_event() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
// While capturing, the events become event-record NODES of the graph (cudaEventRecordExternal): every replay
// re-records them, and prof_replayed() queues them for the next miso_prof_collect.
void prof_capture(std::vector<ProfRec> *sink) { g_sink = sink; }
void prof_replayed(const std::vector<ProfRec> &recs) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto r : recs) {
        r.persistent = true;
        g_prof.push_back(r);
    }
}
void prof_begin(cudaStream_t st) {
    if (!prof_enabled()) return;
    g_open = get_event();
    if (g_sink)
        cudaEventRecordWithFlags(g_open, st, cudaEventRecordExternal);
    else
        cudaEventRecord(g_open, st);
}
void prof_end(cudaStream_t st, double flops, double bytes, int family, double exec_flops) {
    if (!g_open) return;
    ProfRec r;
    r.a = g_open;
    g_open = nullptr;
    r.b = get_event();
    r.flops = flops;
    r.bytes = bytes;
    r.exec_flops = exec_flops;
    r.family = family;
    r.persistent = false;
    if (g_sink) {
        cudaEventRecordWithFlags(r.b, st, cudaEventRecordExternal);
        g_sink->push_back(r);
        return;
    }
    cudaEventRecord(r.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
}

}  // namespace miso

extern "C" {

int miso_prof_enable(int on) {
    miso::g_prof_on.store(on != 0);
    return MISO_OK;
}

static int prof_collect_impl(int family, double *total_ms, double *total_flops, double *total_bytes, double *total_exec, uint64_t *launches);
int miso_prof_collect(int family, double *total_ms, double *total_flops, double *total_bytes, uint64_t *launches) {
    return prof_collect_impl(family, total_ms, total_flops, total_bytes, nullptr, launches);
}
int miso_prof_collect2(int family, double *total_ms, double *total_flops, double *total_bytes, double *total_exec_flops, uint64_t *launches) {
    return prof_collect_impl(family, total_ms, total_flops, total_bytes, total_exec_flops, launches);
}
static int prof_collect_impl(int family, double *total_ms, double *total_flops, double *total_bytes, double *total_exec, uint64_t *launches) {
    std::vector<miso::ProfRec> recs;
    {
        std::lock_guard<std::mutex> lk(miso::g_prof_mu);
        std::vector<miso::ProfRec> keep;
        for (auto &r : miso::g_prof) (family < 0 || r.family == family ? recs : keep).push_back(r);
        miso::g_prof.swap(keep);
    }
    double ms = 0.0, fl = 0.0, by = 0.0, ex = 0.0;
    for (auto &r : recs) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return miso::cuda_fail(e, "cudaEventSynchronize");
        float t = 0.f;
        e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e != cudaSuccess) return miso::cuda_fail(e, "cudaEventElapsedTime");
        ms += t;
        fl += r.flops;
        by += r.bytes;
        ex += r.exec_flops;
    }
    {
        std::lock_guard<std::mutex> lk(miso::g_prof_mu);
        for (auto &r : recs) {
            if (r.persistent) continue;  // owned by a captured graph
            miso::g_pool.push_back(r.a);
            miso::g_pool.push_back(r.b);
        }
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (total_bytes) *total_bytes = by;
    if (total_exec) *total_exec = ex;
    if (launches) *launches = recs.size();
    return MISO_OK;
}


int miso_prof_dump(double *ms, double *flops, int *family, int capacity) {
    // per-record times of everything queued since the last collect, in launch order (records stay queued)
    std::vector<miso::ProfRec> recs;
    {
        std::lock_guard<std::mutex> lk(miso::g_prof_mu);
        recs = miso::g_prof;
    }
    int n = 0;
    for (auto &r : recs) {
        if (n >= capacity) break;
        if (cudaEventSynchronize(r.b) != cudaSuccess) return MISO_E_CUDA;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return MISO_E_CUDA;
        if (ms) ms[n] = t;
        if (flops) flops[n] = r.flops;
        if (family) family[n] = r.family;
        ++n;
    }
    return n;
}

int miso_abi_version(void) { return MISO_ABI_VERSION; }
const char *miso_last_error(void) { return miso::g_err; }
uint64_t miso_launch_count(void) { return miso::g_launch_count.load(); }

int miso_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return miso::cuda_fail(e, "cudaGetDevice");
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return miso::cuda_fail(e, "cudaDeviceGetAttribute");
    if (major != 10) {
        miso::set_error("misonet_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
        return MISO_E_ARCH;
    }
    return MISO_OK;
}

}  // extern "C"

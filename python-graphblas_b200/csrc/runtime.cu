// runtime.cu -- context, stream, stream-ordered memory, errors, options, launch accounting
#include <chrono>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "grb_internal.h"

cudaStream_t g_stream = nullptr;
static cudaStream_t g_own_stream = nullptr;
bool g_initialized = false;
int g_num_sms = 148;
static int g_device = -1;
static std::mutex g_mu;
static std::map<std::string, std::string> g_opts;
static uint64_t g_launches = 0;
static thread_local std::string g_last_error;
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
static size_t g_bytes_in_use = 0;
static std::map<void *, size_t> g_alloc_sizes;

struct KStat { double ms = 0; uint64_t n = 0; };
static std::map<std::string, KStat> g_kstats;
static bool g_profile = false;

void set_last_error(const char *msg) { g_last_error = msg ? msg : ""; }

GrB_Info set_error(std::string *slot, GrB_Info info, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (slot) *slot = buf;
    return info;
}

GrB_Info cuda_fail(std::string *slot, cudaError_t e, const char *what) {
    cudaGetLastError();  // clear sticky-less errors
    GrB_Info info = (e == cudaErrorMemoryAllocation) ? GrB_OUT_OF_MEMORY : GrB_PANIC;
    return set_error(slot, info, "CUDA error %s (%s) in %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
}

extern "C" const char *GrB_cuda_last_error(void) { return g_last_error.c_str(); }

void note_launch(const char *) { g_launches++; }

KernelTimer::KernelTimer(const char *n) : name(n), a(nullptr), b(nullptr), on(g_profile) {
    if (on) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, g_stream);
    }
}
KernelTimer::~KernelTimer() {
    if (on) {
        cudaEventRecord(b, g_stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        KStat &s = g_kstats[name];
        s.ms += ms;
        s.n++;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
}

cudaStream_t aux_stream(int i) {
    static cudaStream_t aux[2] = {nullptr, nullptr};
    if (i < 0 || i > 1) return nullptr;
    if (!aux[i] && cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError();
        aux[i] = nullptr;
    }
    return aux[i];
}
bool profiling() { return g_profile; }

void phase_mark(const char *name) {
    static std::chrono::steady_clock::time_point last;
    if (opt_get_int("trace_host", 0) == 0) return;
    const auto now = std::chrono::steady_clock::now();
    if (name) {
        KStat &s = g_kstats[std::string("host:") + name];
        s.ms += std::chrono::duration<double, std::milli>(now - last).count();
        s.n++;
    }
    last = std::chrono::steady_clock::now();
}

extern "C" uint64_t GrB_cuda_launch_count(void) { return g_launches; }

extern "C" GrB_Info GrB_cuda_kernel_time(const char *name, double *total_ms, uint64_t *launches) {
    if (!name) {  // reset
        g_kstats.clear();
        return GrB_SUCCESS;
    }
    auto it = g_kstats.find(name);
    if (it == g_kstats.end()) {
        if (total_ms) *total_ms = 0;
        if (launches) *launches = 0;
        return GrB_NO_VALUE;
    }
    if (total_ms) *total_ms = it->second.ms;
    if (launches) *launches = it->second.n;
    return GrB_SUCCESS;
}

// names of profiled kernels, NUL separated (for reports)
extern "C" size_t GrB_cuda_kernel_names(char *buf, size_t buflen) {
    size_t need = 0;
    for (auto &kv : g_kstats) need += kv.first.size() + 1;
    if (buf && buflen >= need) {
        char *p = buf;
        for (auto &kv : g_kstats) {
            memcpy(p, kv.first.c_str(), kv.first.size() + 1);
            p += kv.first.size() + 1;
        }
    }
    return need;
}

// Large blocks (>= 1 MiB) released by the library are parked in a size-keyed cache and handed out again to the next request
// of (nearly) the same size, so an iterative workload -- the same mxm / mxv sizes every step -- does no allocator work at all
// in steady state, not even cudaMallocAsync's pool bookkeeping (which showed up as multi-millisecond, erratic gaps between
// steps when 10-20 GB result arrays were freed and re-allocated).  Everything the library enqueues runs on ONE stream, so
// reuse is stream-ordered by construction; GrB_cuda_set_stream synchronises before switching.  The cache gives its blocks
// back on allocation failure, on option "trim" and in GrB_finalize, and is bounded by option "cache_fraction" (default 0.6 of
// device memory).
static std::multimap<size_t, void *> g_block_cache;
static size_t g_cache_bytes = 0, g_cache_limit = 0;
constexpr size_t CACHE_MIN_BYTES = (size_t)1 << 20;

static void cache_flush() {
    std::multimap<size_t, void *> old;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        old.swap(g_block_cache);
        g_cache_bytes = 0;
    }
    for (auto &kv : old) cudaFreeAsync(kv.second, g_stream);
}

void *dev_alloc(size_t bytes) {
    bytes += 64;   // slack: TMA bulk copies read 16-byte granules that may run a few elements past an array's logical end (spgemm_tile.cuh)
    void *p = nullptr;
    if (bytes >= CACHE_MIN_BYTES) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_block_cache.lower_bound(bytes);
        if (it != g_block_cache.end() && it->first <= bytes + bytes / 8) {
            p = it->second;
            g_cache_bytes -= it->first;
            g_alloc_sizes[p] = it->first;
            g_bytes_in_use += it->first;
            g_block_cache.erase(it);
            return p;
        }
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, g_stream);
    if (e != cudaSuccess) {
        // give cached blocks back and retry once
        cudaGetLastError();
        cache_flush();
        cudaStreamSynchronize(g_stream);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, g_device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocAsync(&p, bytes, g_stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            char buf[128];
            snprintf(buf, sizeof buf, "device allocation of %zu bytes failed", bytes);
            g_last_error = buf;
            return nullptr;
        }
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_alloc_sizes[p] = bytes;
        g_bytes_in_use += bytes;
    }
    return p;
}

void dev_free(void *p) {
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_alloc_sizes.find(p);
        if (it != g_alloc_sizes.end()) {
            bytes = it->second;
            g_bytes_in_use -= bytes;
            g_alloc_sizes.erase(it);
        }
        if (bytes >= CACHE_MIN_BYTES && g_initialized) {
            if (g_cache_limit == 0) {
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                const long pct = opt_get_int("cache_percent", 60);
                g_cache_limit = (size_t)((double)total_b * (double)(pct < 0 ? 0 : pct) / 100.0) + 1;
            }
            if (g_cache_bytes + bytes <= g_cache_limit) {
                g_block_cache.emplace(bytes, p);
                g_cache_bytes += bytes;
                return;
            }
        }
    }
    cudaFreeAsync(p, g_stream);
}

extern "C" size_t GrB_cuda_memory_in_use(void) { return g_bytes_in_use; }

// Workspace slots: large scratch buffers that recur with the same size on every call of an iterative workload
// (SpGEMM staging).  Kept across calls so the steady state performs no allocator work at all; dropped by
// GrB_cuda_set_option("trim", "1") / GrB_finalize or when a request no longer fits the cached block.
static struct { void *ptr; size_t bytes; bool busy; } g_ws[8];
void *ws_acquire(int slot, size_t bytes) {
    auto &w = g_ws[slot];
    if (w.ptr && !w.busy && w.bytes >= bytes && w.bytes <= bytes + bytes / 2 + (1 << 20)) { w.busy = true; return w.ptr; }
    if (w.ptr && !w.busy) { dev_free(w.ptr); w.ptr = nullptr; w.bytes = 0; }
    if (w.busy) return dev_alloc(bytes);   // nested use: plain allocation (released by ws_release via pointer check)
    void *p = dev_alloc(bytes);
    if (p) { w.ptr = p; w.bytes = bytes; w.busy = true; }
    return p;
}
void ws_release(int slot, void *p) {
    auto &w = g_ws[slot];
    if (p && p == w.ptr) { w.busy = false; return; }
    dev_free(p);
}
void ws_detach(int slot, void *p) {
    auto &w = g_ws[slot];
    if (p && p == w.ptr) { w.ptr = nullptr; w.bytes = 0; w.busy = false; }
}
void ws_trim() {
    for (auto &w : g_ws)
        if (w.ptr && !w.busy) { dev_free(w.ptr); w.ptr = nullptr; w.bytes = 0; }
    cache_flush();
}

const char *opt_get(const char *key, const char *dflt) {
    auto it = g_opts.find(key);
    if (it != g_opts.end()) return it->second.c_str();
    std::string env = std::string("GRB_CUDA_") + key;
    for (auto &c : env) c = (char)toupper(c);
    const char *e = getenv(env.c_str());
    return e ? e : dflt;
}
long opt_get_int(const char *key, long dflt) {
    const char *v = opt_get(key, nullptr);
    return v ? strtol(v, nullptr, 10) : dflt;
}

extern "C" GrB_Info GrB_cuda_set_option(const char *key, const char *value) {
    if (!key) return GrB_NULL_POINTER;
    if (!value) g_opts.erase(key);
    else g_opts[key] = value;
    if (!strcmp(key, "profile")) g_profile = value && atoi(value) != 0;
    if (!strcmp(key, "trim")) ws_trim();
    return GrB_SUCCESS;
}
extern "C" const char *GrB_cuda_get_option(const char *key) { return key ? opt_get(key, "") : ""; }

extern "C" GrB_Info GrB_cuda_set_device(int device) {
    if (g_initialized && device != g_device)
        return set_error(nullptr, GrB_INVALID_VALUE, "device already bound to %d", g_device);
    g_device = device;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_init(GrB_Mode mode) {
    (void)mode;
    if (g_initialized) return GrB_SUCCESS;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(nullptr, GrB_PANIC, "libgrb_cuda: no CUDA device available (%s); this backend has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (g_device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        g_device = lr ? atoi(lr) % count : 0;
        const char *dv = getenv("GRB_CUDA_DEVICE");
        if (dv) g_device = atoi(dv);
    }
    CUDA_TRY(nullptr, cudaSetDevice(g_device));
    cudaDeviceProp prop;
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, g_device));
    g_num_sms = prop.multiProcessorCount;
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    g_stream = g_own_stream;
    cudaMemPool_t pool;
    CUDA_TRY(nullptr, cudaDeviceGetDefaultMemPool(&pool, g_device));
    uint64_t thresh = UINT64_MAX;  // keep freed blocks cached: no cudaMalloc on the steady-state path
    CUDA_TRY(nullptr, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    CUDA_TRY(nullptr, cudaEventCreate(&g_t0));
    CUDA_TRY(nullptr, cudaEventCreate(&g_t1));
    const char *pf = getenv("GRB_CUDA_PROFILE");
    g_profile = pf && atoi(pf) != 0;
    builtins_init();
    g_initialized = true;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_finalize(void) {
    if (!g_initialized) return GrB_SUCCESS;
    ws_trim();
    cudaStreamSynchronize(g_stream);
    g_initialized = false;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_getVersion(unsigned int *version, unsigned int *subversion) {
    if (version) *version = 2;
    if (subversion) *subversion = 0;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_set_stream(void *s) {
    CHECK_INIT();
    cudaStreamSynchronize(g_stream);
    g_stream = s ? (cudaStream_t)s : g_own_stream;
    return GrB_SUCCESS;
}
extern "C" void *GrB_cuda_get_stream(void) { return (void *)g_stream; }

// Stream ordering without a host synchronisation.  direction 0: the library stream waits for everything enqueued so
// far on `other` (call BEFORE handing device buffers produced on `other` to the library); direction 1: `other` waits
// for the library stream (call AFTER, before `other` frees / reuses those buffers or reads library-owned arrays).
// `other` == NULL or (void*)1 names the legacy NULL stream.
extern "C" GrB_Info GrB_cuda_stream_order(void *other, int direction) {
    CHECK_INIT();
    cudaStream_t o = (other == nullptr || other == (void *)1) ? cudaStreamLegacy : (cudaStream_t)other;
    if (o == g_stream) return GrB_SUCCESS;
    static cudaEvent_t ev = nullptr;
    if (!ev) CUDA_TRY(nullptr, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaStream_t first = direction == 0 ? o : g_stream, second = direction == 0 ? g_stream : o;
    CUDA_TRY(nullptr, cudaEventRecord(ev, first));
    CUDA_TRY(nullptr, cudaStreamWaitEvent(second, ev, 0));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_sync(void) {
    CHECK_INIT();
    CUDA_TRY(nullptr, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_timer_start(void) {
    CHECK_INIT();
    CUDA_TRY(nullptr, cudaEventRecord(g_t0, g_stream));
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_timer_stop(float *ms) {
    CHECK_INIT();
    CUDA_TRY(nullptr, cudaEventRecord(g_t1, g_stream));
    CUDA_TRY(nullptr, cudaEventSynchronize(g_t1));
    float t = 0;
    CUDA_TRY(nullptr, cudaEventElapsedTime(&t, g_t0, g_t1));
    if (ms) *ms = t;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ peer-visible buffers (fused multiply + exchange)
// cudaMallocAsync pool memory cannot be exported with CUDA IPC; buffers other ranks write into come from plain cudaMalloc.
extern "C" GrB_Info GrB_cuda_peer_alloc(void **ptr, size_t bytes) {
    CHECK_INIT();
    if (!ptr) return GrB_NULL_POINTER;
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaMalloc of a peer buffer");
    CUDA_TRY(nullptr, cudaMemsetAsync(*ptr, 0, bytes ? bytes : 16, g_stream));
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_peer_free(void *ptr) {
    if (ptr) { cudaStreamSynchronize(g_stream); cudaFree(ptr); }
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_ipc_get(void *ptr, unsigned char handle[64]) {
    if (!ptr || !handle) return GrB_NULL_POINTER;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(nullptr, cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_ipc_open(const unsigned char handle[64], void **ptr) {
    CHECK_INIT();
    if (!ptr || !handle) return GrB_NULL_POINTER;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(nullptr, cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_ipc_close(void *ptr) {
    if (ptr) { cudaStreamSynchronize(g_stream); CUDA_TRY(nullptr, cudaIpcCloseMemHandle(ptr)); }
    return GrB_SUCCESS;
}

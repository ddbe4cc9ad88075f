// Library plumbing: version, thread-local error text, device queries.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {
thread_local char g_error[512] = "";
}

void qs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// Device properties are cached PER DEVICE: a process may drive several GPUs (cudaSetDevice between calls).
int qs_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

int qs_sm_count() {
    static int sms[kMaxDevices] = {0};
    const int dev = qs_current_device();
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

// ---------------------------------------------------------------------------------------------
// device-resident cache of shape-dependent index tables (see common.cuh)
// ---------------------------------------------------------------------------------------------
namespace {
struct CachedTable {
    std::string key;
    void* dev;
};
std::mutex g_table_mutex;
std::unordered_multimap<uint64_t, CachedTable>* g_tables = nullptr;
size_t g_table_bytes = 0;
constexpr size_t kTableCacheLimit = 512u << 20;  // bytes of device memory over all devices
}  // namespace

uint64_t qs_hash_bytes(const void* data, size_t bytes, uint64_t seed) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint64_t h = 1469598103934665603ull ^ seed;  // FNV-1a
    for (size_t i = 0; i < bytes; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

static std::string full_key(const void* key, size_t key_bytes) {
    const int dev = qs_current_device();
    std::string k(reinterpret_cast<const char*>(&dev), sizeof(dev));
    k.append(static_cast<const char*>(key), key_bytes);
    return k;
}

const void* qs_table_cache_get(const void* key, size_t key_bytes) {
    const std::string k = full_key(key, key_bytes);
    const uint64_t h = qs_hash_bytes(k.data(), k.size(), 0);
    std::lock_guard<std::mutex> lock(g_table_mutex);
    if (!g_tables) return nullptr;
    auto range = g_tables->equal_range(h);
    for (auto it = range.first; it != range.second; ++it)
        if (it->second.key == k) return it->second.dev;
    return nullptr;
}

const void* qs_table_cache_put(const void* key, size_t key_bytes, const void* host_data, size_t bytes) {
    const std::string k = full_key(key, key_bytes);
    const uint64_t h = qs_hash_bytes(k.data(), k.size(), 0);
    std::lock_guard<std::mutex> lock(g_table_mutex);
    if (!g_tables) g_tables = new std::unordered_multimap<uint64_t, CachedTable>();
    if (g_table_bytes + bytes > kTableCacheLimit) return nullptr;
    void* dev = nullptr;
    if (cudaMalloc(&dev, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    // blocking copy: the table is complete in device memory before any stream can be handed the pointer
    if (bytes && cudaMemcpy(dev, host_data, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(dev);
        return nullptr;
    }
    g_table_bytes += bytes;
    g_tables->emplace(h, CachedTable{k, dev});
    return dev;
}

// ---------------------------------------------------------------------------------------------
// instrumentation: launch counter and per-kernel-family device timers (CUDA events on the
// launching stream), read by bench.py for "gpu_launches" and the roofline of the dominant kernel
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxTimed = 8192;
struct TimedSpan {
    cudaEvent_t e0, e1;
    double work;
    int family;
};
std::atomic<long long> g_launches{0};
bool g_timing = false;
std::vector<TimedSpan>* g_spans = nullptr;
}  // namespace

void qs_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool qs_timing_begin(int family, double work, void* stream, int* slot) {
    *slot = -1;
    if (!g_timing || !g_spans || (int)g_spans->size() >= kMaxTimed) return false;
    TimedSpan sp;
    if (cudaEventCreate(&sp.e0) != cudaSuccess) return false;
    if (cudaEventCreate(&sp.e1) != cudaSuccess) {
        cudaEventDestroy(sp.e0);
        return false;
    }
    sp.work = work;
    sp.family = family;
    cudaEventRecord(sp.e0, static_cast<cudaStream_t>(stream));
    g_spans->push_back(sp);
    *slot = (int)g_spans->size() - 1;
    return true;
}

void qs_timing_end(int slot, void* stream) {
    if (slot >= 0 && g_spans && slot < (int)g_spans->size())
        cudaEventRecord((*g_spans)[slot].e1, static_cast<cudaStream_t>(stream));
}

extern "C" int64_t qs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int qs_kernel_timing_enable(int enable) {
    if (!g_spans) g_spans = new std::vector<TimedSpan>();
    for (auto& sp : *g_spans) {
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    g_spans->clear();
    g_timing = enable != 0;
    return QS_OK;
}

extern "C" int qs_kernel_timing_read(int family, double* ms_total, double* work_total, int64_t* spans) {
    QS_REQUIRE(ms_total && work_total && spans, "qs_kernel_timing_read: null pointer");
    *ms_total = 0.0;
    *work_total = 0.0;
    *spans = 0;
    if (!g_spans) return QS_OK;
    for (auto& sp : *g_spans) {
        if (sp.family != family) continue;
        QS_CUDA(cudaEventSynchronize(sp.e1));
        float ms = 0.f;
        QS_CUDA(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
        *ms_total += ms;
        *work_total += sp.work;
        *spans += 1;
    }
    return QS_OK;
}

extern "C" int qs_version(void) { return 100; }  // 0.1.0

extern "C" const char* qs_last_error(void) { return g_error; }

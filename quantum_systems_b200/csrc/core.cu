// Library plumbing: version, thread-local error text, device queries.
#include <stdarg.h>

#include <atomic>
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

int qs_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
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

// Library plumbing: version, thread-local error text, device queries.
#include <stdarg.h>

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

extern "C" int qs_version(void) { return 100; }  // 0.1.0

extern "C" const char* qs_last_error(void) { return g_error; }

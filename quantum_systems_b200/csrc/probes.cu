// Roofline denominators measured in place (MEASURED_PEAKS.json has no FP64 figure):
//   qs_probe_dmma_tflops : register-resident DMMA.8x8x4 loop on every SM sub-partition
//   qs_probe_copy_gbs    : streaming 16-byte copy through HBM
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters) {
    double d0[8], d1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d0[i] = d1[i] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma_8x8x4(d0[i], d1[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d0[i] + d1[i];
    if (s == 12345.678) out[0] = s;  // keep the loop alive without a store per thread
}

__global__ void __launch_bounds__(512) copy_probe_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                                         long long n2) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = in[i];
}

}  // namespace

extern "C" int qs_probe_dmma_tflops(double* host_tflops, void* stream) {
    QS_REQUIRE(host_tflops, "qs_probe_dmma_tflops: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* sink = nullptr;
    QS_CUDA(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    QS_CUDA(cudaEventCreate(&e0));
    QS_CUDA(cudaEventCreate(&e1));
    const int blocks = qs_sm_count() * 2, threads = 256, iters = 20000;
    dmma_probe_kernel<<<blocks, threads, 0, st>>>(sink, 1000);  // warm-up
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        QS_CUDA(cudaEventRecord(e0, st));
        dmma_probe_kernel<<<blocks, threads, 0, st>>>(sink, iters);
        QS_CUDA(cudaEventRecord(e1, st));
        QS_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        QS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    QS_CUDA(cudaEventDestroy(e0));
    QS_CUDA(cudaEventDestroy(e1));
    QS_CUDA(cudaFree(sink));
    const double flops = (double)blocks * (threads / 32) * (double)iters * 8 * 512.0;
    *host_tflops = flops / (best * 1e-3) * 1e-12;
    return QS_OK;
}

extern "C" int qs_probe_copy_gbs(double* host_gbs, void* scratch, int64_t scratch_bytes, void* stream) {
    QS_REQUIRE(host_gbs && scratch && scratch_bytes >= (1 << 20), "qs_probe_copy_gbs: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long half = (scratch_bytes / 2) & ~1023LL;
    const double2* in = static_cast<const double2*>(scratch);
    double2* out = reinterpret_cast<double2*>(static_cast<char*>(scratch) + half);
    cudaEvent_t e0, e1;
    QS_CUDA(cudaEventCreate(&e0));
    QS_CUDA(cudaEventCreate(&e1));
    const int blocks = qs_sm_count() * 8;
    copy_probe_kernel<<<blocks, 512, 0, st>>>(in, out, half / 16);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        QS_CUDA(cudaEventRecord(e0, st));
        copy_probe_kernel<<<blocks, 512, 0, st>>>(in, out, half / 16);
        QS_CUDA(cudaEventRecord(e1, st));
        QS_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        QS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    QS_CUDA(cudaEventDestroy(e0));
    QS_CUDA(cudaEventDestroy(e1));
    *host_gbs = 2.0 * half / (best * 1e-3) * 1e-9;
    return QS_OK;
}

// FP64 peak probes for the B200 roofline denominators (DMMA.8x8x4 tensor pipe, DFMA pipe, HBM copy).
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
// MEASURED_PEAKS.json (driver-written) has no FP64 figure, so this tool measures one on the box.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_loop(double* out, int iters, long long* cyc) {
    double acc0[NACC], acc1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc0[i] = 0.0; acc1[i] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(acc0[i], acc1[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc0[i] + acc1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NACC>
__global__ void dfma_loop(double* out, int iters, long long* cyc) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = in[i];
}

template <int NACC>
void run_dmma(int blocks, int threads, int iters, double* out, long long* cyc, const char* tag) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dmma_loop<NACC><<<blocks, threads>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        dmma_loop<NACC><<<blocks, threads>>>(out, iters, cyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    long long c; CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    double warps = (double)blocks * threads / 32.0;
    double flops = warps * (double)iters * NACC * 512.0;
    printf("{\"probe\":\"%s\",\"nacc\":%d,\"blocks\":%d,\"threads\":%d,\"ms\":%.4f,\"tflops\":%.3f,\"cyc_per_dmma_per_warp\":%.2f}\n",
           tag, NACC, blocks, threads, best, flops / best * 1e-9, (double)c / ((double)iters * NACC));
}

template <int NACC>
void run_dfma(int blocks, int threads, int iters, double* out, long long* cyc) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dfma_loop<NACC><<<blocks, threads>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        dfma_loop<NACC><<<blocks, threads>>>(out, iters, cyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    long long c; CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    double flops = (double)blocks * threads * (double)iters * NACC * 2.0;
    printf("{\"probe\":\"dfma\",\"nacc\":%d,\"blocks\":%d,\"threads\":%d,\"ms\":%.4f,\"tflops\":%.3f,\"cyc_per_dfma_per_warp\":%.2f}\n",
           NACC, blocks, threads, best, flops / best * 1e-9, (double)c / ((double)iters * NACC));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);
    double* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024)); CK(cudaMalloc(&cyc, 8));
    // latency: one warp, one accumulator chain
    run_dmma<1>(1, 32, 20000, out, cyc, "dmma_latency_1warp");
    run_dmma<2>(1, 32, 20000, out, cyc, "dmma_1warp");
    run_dmma<4>(1, 32, 20000, out, cyc, "dmma_1warp");
    run_dmma<8>(1, 32, 20000, out, cyc, "dmma_1warp");
    run_dmma<16>(1, 32, 20000, out, cyc, "dmma_1warp");
    // one SM, 4 warps (one per SMSP)
    run_dmma<8>(1, 128, 20000, out, cyc, "dmma_1sm_4warps");
    run_dmma<8>(1, 256, 20000, out, cyc, "dmma_1sm_8warps");
    run_dmma<8>(1, 512, 20000, out, cyc, "dmma_1sm_16warps");
    // full chip
    run_dmma<8>(sms, 128, 40000, out, cyc, "dmma_chip_4warps");
    run_dmma<8>(sms, 256, 40000, out, cyc, "dmma_chip_8warps");
    run_dmma<16>(sms, 256, 20000, out, cyc, "dmma_chip_8warps");
    run_dmma<8>(sms, 512, 20000, out, cyc, "dmma_chip_16warps");
    run_dmma<16>(sms * 2, 256, 20000, out, cyc, "dmma_chip_2x8warps");
    run_dmma<8>(sms, 1024, 10000, out, cyc, "dmma_chip_32warps");
    // sustained: ~1-2 s loop
    run_dmma<16>(sms * 2, 256, 2000000, out, cyc, "dmma_chip_sustained");
    run_dfma<1>(1, 32, 20000, out, cyc);
    run_dfma<8>(1, 32, 20000, out, cyc);
    run_dfma<8>(sms, 256, 40000, out, cyc);
    run_dfma<8>(sms, 512, 40000, out, cyc);
    run_dfma<8>(sms, 1024, 40000, out, cyc);
    run_dfma<16>(sms * 2, 512, 40000, out, cyc);
    // HBM copy
    size_t n = (size_t)1 << 28;  // 2 GiB of doubles each way
    double *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
    CK(cudaMemset(a, 1, n * 8)); CK(cudaMemset(b, 0, n * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int bl = 1; bl <= 16; bl *= 2) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0));
            copy_kernel<<<sms * bl, 512>>>((const double2*)a, (double2*)b, n / 2);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("{\"probe\":\"hbm_copy\",\"blocks_per_sm\":%d,\"ms\":%.4f,\"gbs\":%.1f}\n", bl, best, 2.0 * n * 8 / best * 1e-6);
    }
    {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0));
            CK(cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("{\"probe\":\"hbm_memcpy_d2d\",\"ms\":%.4f,\"gbs\":%.1f}\n", best, 2.0 * n * 8 / best * 1e-6);
    }
    return 0;
}

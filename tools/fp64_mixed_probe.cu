// Do the FP64 tensor pipe (DMMA.8x8x4) and the FP64 FMA pipe (DFMA) of a B200 SM run concurrently?
// Half of the warps of every CTA issue a register-resident DMMA loop, the other half a DFMA loop.  If the two
// instruction kinds execute on separate hardware the mixed kernel takes max(T_dmma, T_dfma); if they share the
// FP64 datapath it takes T_dmma + T_dfma.  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mixed_probe tools/fp64_mixed_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// warps with ((warp >> 2) & 1) == 0 run `iters_d` rounds of 8 DMMAs, the others `iters_f` rounds of 16 DFMAs
__global__ void mixed_loop(double* out, int iters_d, int iters_f) {
    const int warp = threadIdx.x >> 5;
    double s = 0.0;
    if (((warp >> 2) & 1) == 0) {  // warps 0-3 (one per SM sub-partition) DMMA, warps 4-7 DFMA, ...
        double acc0[8], acc1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc0[i] = 0.0; acc1[i] = 0.0; }
        double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
        for (int it = 0; it < iters_d; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(acc0[i], acc1[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += acc0[i] + acc1[i];
    } else {
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = i;
        double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
        for (int it = 0; it < iters_f; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_it(double* out, int blocks, int threads, int id, int jf) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mixed_loop<<<blocks, threads>>>(out, id, jf); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        mixed_loop<<<blocks, threads>>>(out, id, jf);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    for (int threads : {256, 512}) {
        const int warps_each = threads / 64;                 // DMMA warps = DFMA warps per CTA
        const int id = 40000, jf = 40000 * 2;                // rounds; calibrated below by the solo timings
        const float td = time_it(out, sms, threads, id, 0);
        const float tf = time_it(out, sms, threads, 0, jf);
        const float tb = time_it(out, sms, threads, id, jf);
        const double fd = (double)sms * warps_each * id * 8 * 512.0;        // DMMA flops
        const double ff = (double)sms * warps_each * 32 * (double)jf * 16 * 2.0;  // DFMA flops
        printf("{\"probe\":\"mixed_dmma_dfma\",\"threads\":%d,\"dmma_only_ms\":%.4f,\"dfma_only_ms\":%.4f,\"both_ms\":%.4f,"
               "\"dmma_only_tflops\":%.2f,\"dfma_only_tflops\":%.2f,\"both_tflops\":%.2f,\"sum_ms\":%.4f,\"max_ms\":%.4f}\n",
               threads, td, tf, tb, fd / td * 1e-9, ff / tf * 1e-9, (fd + ff) / tb * 1e-9, td + tf, td > tf ? td : tf);
    }
    return 0;
}

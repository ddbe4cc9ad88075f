// NVLink store-pattern probe (2 GPUs, one process): how fast can SMs of GPU 0 write into GPU 1's memory, and both
// directions at once, for the store shapes the scattering epilogue of the quarter GEMM can produce?
//
//   lines   : every warp store instruction writes 4 separate 128-byte lines (8 lanes x 16 B each), 4 "columns" that
//             are `pitch` bytes apart -- the current epilogue (quarter_gemm.cu, vec2 path)
//   run512  : every warp store instruction writes 512 contiguous bytes (32 lanes x 16 B)
//   bulk    : one elected lane issues cp.async.bulk shared -> global copies of `run` bytes (512 / 1024 / 2048)
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peer_store_probe tools/peer_store_probe.cu
// Run  :  tools/peer_store_probe            (prints one JSON line per pattern: GB/s one-way and both-ways)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e = (x);                                                                      \
        if (e != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__);    \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

// bytes per CTA tile: 64 columns x 1 KiB (128 rows x 8 B), like a 128 x 64 real accumulator tile
constexpr int kCols = 64;
constexpr int kColBytes = 1024;
constexpr int kTileBytes = kCols * kColBytes;

// tile t, column c lives at  base + c * pitch + t * 1024   (pitch >= tiles * 1024): columns far apart, rows adjacent
__global__ void __launch_bounds__(256) lines_kernel(char* dst, long long pitch, long long tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        // warp w owns rows [16 w, 16 w + 16) of the 128: one 128-byte line per column
        for (int cg = 0; cg < kCols / 4; ++cg) {
            const int c = 4 * cg + t;
            double2 v = make_double2((double)tile, (double)c);
            *reinterpret_cast<double2*>(dst + c * pitch + tile * kColBytes + warp * 128 + g * 16) = v;
        }
    }
}

__global__ void __launch_bounds__(256) run512_kernel(char* dst, long long pitch, long long tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        // warp w writes columns w, w + 8, ...: two 512-byte halves of each column's 1 KiB run
        for (int c = warp; c < kCols; c += 8)
            for (int half = 0; half < 2; ++half) {
                double2 v = make_double2((double)tile, (double)c);
                *reinterpret_cast<double2*>(dst + c * pitch + tile * kColBytes + half * 512 + lane * 16) = v;
            }
    }
}

__global__ void __launch_bounds__(256) bulk_kernel(char* dst, long long pitch, long long tiles, int run) {
    extern __shared__ __align__(128) unsigned char stage[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kTileBytes / 16; i += blockDim.x)
        reinterpret_cast<double2*>(stage)[i] = make_double2(1.0, 2.0);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
    const int per_col = kColBytes / run;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        // 64 * per_col copies per tile, spread over the 256 threads
        for (int i = threadIdx.x; i < kCols * per_col; i += blockDim.x) {
            const int c = i / per_col, part = i % per_col;
            char* g = dst + c * pitch + tile * kColBytes + part * run;
            const uint32_t s = sbase + c * kColBytes + part * run;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(run)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging tile may be rewritten
        (void)warp; (void)lane;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Side {
    int dev;
    char* remote;   // memory on the OTHER device
    cudaStream_t st;
    cudaEvent_t e0, e1;
};

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        printf("{\"error\": \"needs 2 GPUs, found %d\"}\n", ndev);
        return 0;
    }
    const long long tiles = 16384;                     // 1 GiB per direction
    const long long pitch = tiles * kColBytes;         // column c at c * 16 MiB
    const size_t bytes = (size_t)kCols * pitch;
    char* mem[2];
    Side side[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&mem[d], bytes));
        CK(cudaMemset(mem[d], 0, bytes));
    }
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        side[d].dev = d;
        side[d].remote = mem[1 - d];
        CK(cudaStreamCreate(&side[d].st));
        CK(cudaEventCreate(&side[d].e0));
        CK(cudaEventCreate(&side[d].e1));
        CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes));
        CK(cudaDeviceSynchronize());
    }
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    auto launch = [&](int pattern, Side& s, char* dst, int ctas_per_sm) {
        CK(cudaSetDevice(s.dev));
        const int grid = sms * ctas_per_sm;
        CK(cudaEventRecord(s.e0, s.st));
        if (pattern == 0) lines_kernel<<<grid, 256, 0, s.st>>>(dst, pitch, tiles);
        else if (pattern == 1) run512_kernel<<<grid, 256, 0, s.st>>>(dst, pitch, tiles);
        else bulk_kernel<<<grid, 256, kTileBytes, s.st>>>(dst, pitch, tiles, pattern);
        CK(cudaGetLastError());
        CK(cudaEventRecord(s.e1, s.st));
    };
    auto elapsed = [&](Side& s) {
        CK(cudaSetDevice(s.dev));
        CK(cudaEventSynchronize(s.e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.e0, s.e1));
        return (double)ms;
    };
    const int patterns[] = {0, 1, 512, 1024, 256};
    const char* names[] = {"lines_4x128B_per_warp_store", "run_512B_per_warp_store", "bulk_s2g_512B", "bulk_s2g_1024B",
                           "bulk_s2g_256B"};
    for (int pi = 0; pi < 5; ++pi) {
        for (int ctas = 1; ctas <= (patterns[pi] <= 1 ? 4 : 2); ctas *= 2) {
            double local = 0, one = 0, both = 0;
            for (int rep = 0; rep < 3; ++rep) {
                // local HBM for reference, then one direction, then both directions at once
                launch(patterns[pi], side[0], mem[0], ctas);
                local = elapsed(side[0]);
                launch(patterns[pi], side[0], side[0].remote, ctas);
                one = elapsed(side[0]);
                launch(patterns[pi], side[0], side[0].remote, ctas);
                launch(patterns[pi], side[1], side[1].remote, ctas);
                const double a = elapsed(side[0]), b = elapsed(side[1]);
                both = a > b ? a : b;
            }
            const double gb = (double)tiles * kTileBytes * 1e-9;
            printf("{\"pattern\": \"%s\", \"ctas_per_sm\": %d, \"local_gbs\": %.1f, \"one_way_gbs\": %.1f, "
                   "\"both_ways_gbs_per_direction\": %.1f}\n",
                   names[pi], ctas, gb / (local * 1e-3), gb / (one * 1e-3), gb / (both * 1e-3));
            fflush(stdout);
        }
    }
    return 0;
}

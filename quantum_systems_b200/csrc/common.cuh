// Shared device/host helpers for libqsb200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/qsb200.h"

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void qs_set_error(const char* fmt, ...);

#define QS_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            qs_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__,     \
                         __LINE__);                                                            \
            return QS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define QS_REQUIRE(cond, ...)                                                                  \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            qs_set_error(__VA_ARGS__);                                                         \
            return QS_ERR_INVALID;                                                             \
        }                                                                                      \
    } while (0)

#define QS_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        qs_count_launch();                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            qs_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                            \
            return QS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static inline int64_t qs_round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
static inline int64_t qs_ceil_div(int64_t v, int64_t m) { return (v + m - 1) / m; }
static inline int qs_elem_doubles(int dtype) { return dtype == QS_C128 ? 2 : 1; }

// Instrumentation (core.cu): every kernel launch is counted; when timing is enabled the launch
// sites of a kernel family bracket themselves with CUDA events on the launching stream.
void qs_count_launch();
bool qs_timing_begin(int family, double work, void* stream, int* slot);
void qs_timing_end(int slot, void* stream);

// Number of SMs of the current device, and the device ordinal itself (clamped to [0, kMaxDevices)); both
// caches are per device so that one process may drive several GPUs.
constexpr int kMaxDevices = 64;
int qs_current_device();
int qs_sm_count();

// Shape-dependent index tables (tile lists of masked launches, pair tables of the packed layouts) are the same for
// every call with the same extents.  They are built on the host ONCE per (device, key), uploaded with a blocking
// copy and kept in device memory for the life of the process, so that a steady-state transform issues no
// host-to-device copy at all (a cudaMemcpyAsync from pageable memory synchronises the stream first and drains the
// launch queue).  `key` is any byte string that determines the contents; returns nullptr when the cache is full
// (the caller then stages the table through its workspace as before).
const void* qs_table_cache_get(const void* key, size_t key_bytes);
const void* qs_table_cache_put(const void* key, size_t key_bytes, const void* host_data, size_t bytes);
uint64_t qs_hash_bytes(const void* data, size_t bytes, uint64_t seed);

// Internal (not part of the C ABI): symmetry mask of a quarter transform (see quarter_gemm.cu, tile_wanted) and
// the masked launch used by the symmetry-aware four-index transform (transform.cu).
// Which of the two ordered pairs (r, s), (s, r) the SHARDED symmetry-aware transform computes: the one whose
// cyclic distance d = (s - r) mod m is the shorter (ties, 2 d = m, go to r < s).  Every r then has the same number
// of partners s, so any partition of r over the ranks stays balanced.
#ifdef __CUDACC__
__host__ __device__
#endif
static inline bool qs_cyclic_wanted(long long r, long long s, long long m) {
    const long long d = s >= r ? s - r : s - r + m;
    return d > 0 && (2 * d < m || (2 * d == m && r < s));
}

struct QsTileMask {
    int kind;             // 0 none; 1 column < row_lo; 2 row_hi < row_lo; 3 cyclic pair rule on (column, row_lo):
                          //   wanted iff qs_cyclic_wanted(column, row_lo, ml), or (strict == 0) the same with the
                          //   other member row_lo ^ 1 of the aligned couple (the padded pair lists of the real case)
    int strict;           // 1: "<", 0: "<="
    int64_t dh, mh, dl, ml;  // row_hi(x) = (x / dh) % mh, row_lo(x) = (x / dl) % ml
};
int64_t qs_tile_list_bytes(int64_t X, int64_t K, int64_t W, int a_dtype, int m_dtype);
int qs_quarter_transform_masked(const void* A, int a_dtype, int64_t X, int64_t K, int64_t lda, const void* image,
                                int m_dtype, int64_t W, void* out, int64_t x_inner, int64_t sx0, int64_t sx1,
                                int64_t w_inner, int64_t sw0, int64_t sw1, const QsTileMask* mask, void* list_ws,
                                const long long* xq_table, const long long* xr_table, int xq_even, int xr_paired,
                                void* stream);

int qs_mirror_fill(void* out, int dtype, int64_t m, int mode, void* stream);

// Internal (not part of the C ABI): coefficient image of the ODQD shielded-Coulomb matrix.
int qs_build_coulomb_image(const double* grid, double alpha, double a, int64_t Gp, void* image, void* stream);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA (tensor and 1-D bulk), FP64 tensor-core MMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QS_DONE_%=;\n"
        "bra QS_WAIT_%=;\n"
        "QS_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// 2-D tiled TMA load: box at (c0 = fastest coordinate, c1) -> shared, completes on `bar`.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// 1-D bulk copy global -> shared (size multiple of 16 B), completes on `bar`.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col) in FP64 on the tensor pipe; SASS: DMMA.8x8x4.
// Lane (g = lane >> 2, t = lane & 3) supplies A[g][t], B[t][g] and owns D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void sts_128(uint32_t addr, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}

// Asynchronous bulk copy shared -> global (16-byte aligned on both sides, size a multiple of 16 B), tracked by the
// issuing thread's bulk async-groups (cp.async.bulk.commit_group / wait_group[.read]).
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ double2 lds_128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
#endif  // __CUDACC__

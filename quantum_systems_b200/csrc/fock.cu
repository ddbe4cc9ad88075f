// Fock-matrix construction: strided gathers over the occupied index, reduced with warp shuffles.
//
//   qs_fock_general : f = h + sum_i u[p,i,q,i]                 (reference general_orbital_system.py:119-159)
//   qs_fock_spatial : f = h + 2 sum_i u[p,i,q,i] - sum_i u[p,i,i,q]   (reference spatial_orbital_system.py:150-190)
//
// One warp per output element (p, q): lanes stride over the occupied index i, each lane gathers
// u[p,i,q,i] (one 8/16-byte element per 32-byte sector: the pass is sector/latency-bound, not
// bandwidth-bound), then a butterfly __shfl_xor reduction; lane 0 adds h and writes f.
#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <bool H_COMPLEX, bool U_COMPLEX, bool SPATIAL>
__global__ void __launch_bounds__(256) fock_kernel(const double* __restrict__ h, const double* __restrict__ u,
                                                   double* __restrict__ f, int n, int n_occ, long long p_begin,
                                                   long long p_end) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long rows = p_end - p_begin;
    for (long long e = warp_global; e < rows * n; e += warps_total) {
        const long long pl = e / n;  // local row (u points at plane p_begin)
        const int q = (int)(e - pl * n);
        const double* up = u + pl * (long long)n * n * n * (U_COMPLEX ? 2 : 1);
        double sr = 0.0, si = 0.0;
        for (int i = lane; i < n_occ; i += 32) {
            const long long direct = ((long long)i * n + q) * n + i;  // u[p,i,q,i]
            if (U_COMPLEX) {
                const double2 v = reinterpret_cast<const double2*>(up)[direct];
                sr += SPATIAL ? 2.0 * v.x : v.x;
                si += SPATIAL ? 2.0 * v.y : v.y;
            } else {
                sr += SPATIAL ? 2.0 * up[direct] : up[direct];
            }
            if (SPATIAL) {
                const long long exch = ((long long)i * n + i) * n + q;  // u[p,i,i,q]
                if (U_COMPLEX) {
                    const double2 v = reinterpret_cast<const double2*>(up)[exch];
                    sr -= v.x;
                    si -= v.y;
                } else {
                    sr -= up[exch];
                }
            }
        }
        sr = warp_sum(sr);
        if (U_COMPLEX) si = warp_sum(si);
        if (lane == 0) {
            const long long o = (p_begin + pl) * n + q;
            if (H_COMPLEX) {
                const double2 hv = reinterpret_cast<const double2*>(h)[o];
                reinterpret_cast<double2*>(f)[o] = make_double2(hv.x + sr, hv.y + si);
            } else {
                f[o] = h[o] + sr;
            }
        }
    }
}

template <bool SPATIAL>
int launch_fock(const void* h, int h_dtype, const void* u, int u_dtype, int64_t n, int64_t n_occ, void* f,
                int64_t p_begin, int64_t p_end, void* stream) {
    QS_REQUIRE(h && u && f && n > 0, "qs_fock: bad arguments");
    QS_REQUIRE(0 <= n_occ && n_occ <= n, "qs_fock: n_occ out of range");
    QS_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= n, "qs_fock: bad row range");
    QS_REQUIRE(!(h_dtype == QS_F64 && u_dtype == QS_C128),
               "qs_fock: complex u cannot be accumulated into a real Fock matrix");
    if (p_begin == p_end) return QS_OK;
    const long long warps = (p_end - p_begin) * n;
    long long blocks = qs_ceil_div(warps, 8);
    const long long cap = (long long)qs_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double* hp = static_cast<const double*>(h);
    const double* up = static_cast<const double*>(u);
    double* fp = static_cast<double*>(f);
    if (h_dtype == QS_C128 && u_dtype == QS_C128)
        fock_kernel<true, true, SPATIAL><<<(unsigned)blocks, 256, 0, st>>>(hp, up, fp, (int)n, (int)n_occ, p_begin, p_end);
    else if (h_dtype == QS_C128)
        fock_kernel<true, false, SPATIAL><<<(unsigned)blocks, 256, 0, st>>>(hp, up, fp, (int)n, (int)n_occ, p_begin, p_end);
    else
        fock_kernel<false, false, SPATIAL><<<(unsigned)blocks, 256, 0, st>>>(hp, up, fp, (int)n, (int)n_occ, p_begin, p_end);
    QS_LAUNCH_CHECK();
    return QS_OK;
}

}  // namespace

extern "C" int qs_fock_general(const void* h, int h_dtype, const void* u, int u_dtype, int64_t n, int64_t n_occ, void* f,
                               int64_t p_begin, int64_t p_end, void* stream) {
    return launch_fock<false>(h, h_dtype, u, u_dtype, n, n_occ, f, p_begin, p_end, stream);
}

extern "C" int qs_fock_spatial(const void* h, int h_dtype, const void* u, int u_dtype, int64_t n, int64_t n_occ, void* f,
                               int64_t p_begin, int64_t p_end, void* stream) {
    return launch_fock<true>(h, h_dtype, u, u_dtype, n, n_occ, f, p_begin, p_end, stream);
}

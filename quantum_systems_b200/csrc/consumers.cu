// Access patterns of the solvers that consume the two-body tensor, as memory-bound
// passes that work on one GPU's leading-index shard:
//
//   qs_extract_block  : dense copy of u[a0:a1, b0:b1, c0:c1, d0:d1]  -- the u[o,o,v,v]-style slicing of
//                       coupled-cluster / HF codes on top of QuantumSystem.o / .v (reference system.py:47-51)
//   qs_scale_add      : out = alpha x + beta y -- AdiabaticSwitching.u_t = f(t) u and the sums of
//                       QuantumSystem.h_t / u_t (reference system.py:189-215, time_evolution_operators/operator.py:182-196)
//   qs_occupied_traces: tr h[o,o], sum_ij u[i,j,i,j], sum_ij u[i,j,j,i] -- the three numbers behind
//                       compute_reference_energy (reference general_orbital_system.py:75-117,
//                       spatial_orbital_system.py:106-148)
#include "common.cuh"

namespace {

struct BlockShape {
    long long n;               // extent of the three trailing axes of the source
    long long a0, b0, c0, d0;  // origin of the block (a0 relative to the first plane behind the pointer)
    long long na, nb, nc, nd;  // extents of the block
};

template <typename T>
__global__ void __launch_bounds__(256) extract_block_kernel(const T* __restrict__ u, T* __restrict__ out, BlockShape s) {
    const long long total = s.na * s.nb * s.nc * s.nd;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long d = e % s.nd;
        long long rest = e / s.nd;
        const long long c = rest % s.nc;
        rest /= s.nc;
        const long long b = rest % s.nb;
        const long long a = rest / s.nb;
        out[e] = u[(((s.a0 + a) * s.n + s.b0 + b) * s.n + s.c0 + c) * s.n + s.d0 + d];
    }
}

template <bool COMPLEX, bool HAS_Y>
__global__ void __launch_bounds__(256) scale_add_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                        double* __restrict__ out, long long count, double ar,
                                                        double ai, double br, double bi) {
    // `count` elements; two doubles (one complex element, or two real ones) per 16-byte access
    const long long pairs = COMPLEX ? count : count / 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const double2* y2 = reinterpret_cast<const double2*>(y);
    double2* o2 = reinterpret_cast<double2*>(out);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < pairs; e += stride) {
        const double2 xv = x2[e];
        double2 r;
        if (COMPLEX) {
            r.x = ar * xv.x - ai * xv.y;
            r.y = ar * xv.y + ai * xv.x;
        } else {
            r.x = ar * xv.x;
            r.y = ar * xv.y;
        }
        if (HAS_Y) {
            const double2 yv = y2[e];
            if (COMPLEX) {
                r.x += br * yv.x - bi * yv.y;
                r.y += br * yv.y + bi * yv.x;
            } else {
                r.x += br * yv.x;
                r.y += br * yv.y;
            }
        }
        o2[e] = r;
    }
    if (!COMPLEX && (count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double r = ar * x[count - 1];
        if (HAS_Y) r += br * y[count - 1];
        out[count - 1] = r;
    }
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += scratch[w];
    return total;  // valid in thread 0
}

// One CTA: the occupied corner holds n_occ^2 elements per trace, far too few to need more.
template <bool H_COMPLEX, bool U_COMPLEX>
__global__ void __launch_bounds__(256) occupied_traces_kernel(const double* __restrict__ h, const double* __restrict__ u,
                                                              long long n, long long n_occ, long long p_begin,
                                                              long long p_end, double* __restrict__ out) {
    __shared__ double scratch[8];
    double acc[6] = {0, 0, 0, 0, 0, 0};  // tr h (re, im), direct (re, im), exchange (re, im)
    const long long i_end = p_end < n_occ ? p_end : n_occ;
    const long long rows = i_end > p_begin ? i_end - p_begin : 0;
    for (long long e = threadIdx.x; e < rows * n_occ; e += blockDim.x) {
        const long long i = p_begin + e / n_occ, j = e % n_occ;
        const long long il = i - p_begin;  // u points at plane p_begin
        const long long direct = ((il * n + j) * n + i) * n + j;
        const long long exchange = ((il * n + j) * n + j) * n + i;
        if (U_COMPLEX) {
            const double2 dv = reinterpret_cast<const double2*>(u)[direct];
            const double2 ev = reinterpret_cast<const double2*>(u)[exchange];
            acc[2] += dv.x, acc[3] += dv.y, acc[4] += ev.x, acc[5] += ev.y;
        } else {
            acc[2] += u[direct], acc[4] += u[exchange];
        }
        if (j == 0) {
            if (H_COMPLEX) {
                const double2 hv = reinterpret_cast<const double2*>(h)[i * n + i];
                acc[0] += hv.x, acc[1] += hv.y;
            } else {
                acc[0] += h[i * n + i];
            }
        }
    }
    for (int k = 0; k < 6; ++k) {
        const double total = block_sum(acc[k], scratch);
        if (threadIdx.x == 0) out[k] = total;
    }
}

int grid_for(long long work_items) {
    const long long want = qs_ceil_div(work_items, 256);
    const long long cap = (long long)qs_sm_count() * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" int qs_extract_block(const void* u, int dtype, int64_t n, int64_t planes, int64_t a0, int64_t a1,
                                int64_t b0, int64_t b1, int64_t c0, int64_t c1, int64_t d0, int64_t d1, void* out,
                                void* stream) {
    QS_REQUIRE(dtype == QS_F64 || dtype == QS_C128, "qs_extract_block: bad dtype");
    QS_REQUIRE(n > 0 && 0 <= a0 && a0 <= a1 && a1 <= planes && 0 <= b0 && b0 <= b1 && b1 <= n && 0 <= c0 &&
                   c0 <= c1 && c1 <= n && 0 <= d0 && d0 <= d1 && d1 <= n,
               "qs_extract_block: block [%lld:%lld, %lld:%lld, %lld:%lld, %lld:%lld] outside (%lld, %lld, %lld, %lld)",
               (long long)a0, (long long)a1, (long long)b0, (long long)b1, (long long)c0, (long long)c1, (long long)d0,
               (long long)d1, (long long)planes, (long long)n, (long long)n, (long long)n);
    BlockShape s{n, a0, b0, c0, d0, a1 - a0, b1 - b0, c1 - c0, d1 - d0};
    const long long total = s.na * s.nb * s.nc * s.nd;
    if (total == 0) return QS_OK;
    QS_REQUIRE(u && out, "qs_extract_block: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int slot = -1;
    qs_timing_begin(QS_FAMILY_SPIN_PASS, 2.0 * total * 8 * qs_elem_doubles(dtype), stream, &slot);
    if (dtype == QS_C128)
        extract_block_kernel<double2><<<grid_for(total), 256, 0, st>>>(static_cast<const double2*>(u),
                                                                      static_cast<double2*>(out), s);
    else
        extract_block_kernel<double><<<grid_for(total), 256, 0, st>>>(static_cast<const double*>(u),
                                                                     static_cast<double*>(out), s);
    QS_LAUNCH_CHECK();
    qs_timing_end(slot, stream);
    return QS_OK;
}

extern "C" int qs_scale_add(const void* x, const void* y, int dtype, int64_t count, double alpha_re, double alpha_im,
                            double beta_re, double beta_im, void* out, void* stream) {
    QS_REQUIRE(dtype == QS_F64 || dtype == QS_C128, "qs_scale_add: bad dtype");
    QS_REQUIRE(count >= 0, "qs_scale_add: negative count");
    QS_REQUIRE(dtype == QS_C128 || (alpha_im == 0.0 && beta_im == 0.0),
               "qs_scale_add: complex factor on a real tensor");
    if (count == 0) return QS_OK;
    QS_REQUIRE(x && out, "qs_scale_add: null pointer");
    QS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0,
               "qs_scale_add: operands must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool cplx = dtype == QS_C128;
    const long long pairs = cplx ? count : (count + 1) / 2;
    const int grid = grid_for(pairs);
    const double* xp = static_cast<const double*>(x);
    const double* yp = static_cast<const double*>(y);
    double* op = static_cast<double*>(out);
    int slot = -1;
    qs_timing_begin(QS_FAMILY_SPIN_PASS, (y ? 3.0 : 2.0) * count * 8 * qs_elem_doubles(dtype), stream, &slot);
    if (cplx && y)
        scale_add_kernel<true, true><<<grid, 256, 0, st>>>(xp, yp, op, count, alpha_re, alpha_im, beta_re, beta_im);
    else if (cplx)
        scale_add_kernel<true, false><<<grid, 256, 0, st>>>(xp, yp, op, count, alpha_re, alpha_im, beta_re, beta_im);
    else if (y)
        scale_add_kernel<false, true><<<grid, 256, 0, st>>>(xp, yp, op, count, alpha_re, alpha_im, beta_re, beta_im);
    else
        scale_add_kernel<false, false><<<grid, 256, 0, st>>>(xp, yp, op, count, alpha_re, alpha_im, beta_re, beta_im);
    QS_LAUNCH_CHECK();
    qs_timing_end(slot, stream);
    return QS_OK;
}

extern "C" int qs_occupied_traces(const void* h, int h_dtype, const void* u, int u_dtype, int64_t n, int64_t n_occ,
                                  int64_t p_begin, int64_t p_end, double* out6, void* stream) {
    QS_REQUIRE(h && out6 && n > 0, "qs_occupied_traces: bad arguments");
    QS_REQUIRE(0 <= n_occ && n_occ <= n, "qs_occupied_traces: n_occ out of range");
    QS_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= n, "qs_occupied_traces: bad plane range");
    QS_REQUIRE(u || p_begin == p_end || p_begin >= n_occ, "qs_occupied_traces: null tensor");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double* hp = static_cast<const double*>(h);
    const double* up = static_cast<const double*>(u);
    const bool hc = h_dtype == QS_C128, uc = u_dtype == QS_C128;
    if (hc && uc)
        occupied_traces_kernel<true, true><<<1, 256, 0, st>>>(hp, up, n, n_occ, p_begin, p_end, out6);
    else if (hc)
        occupied_traces_kernel<true, false><<<1, 256, 0, st>>>(hp, up, n, n_occ, p_begin, p_end, out6);
    else if (uc)
        occupied_traces_kernel<false, true><<<1, 256, 0, st>>>(hp, up, n, n_occ, p_begin, p_end, out6);
    else
        occupied_traces_kernel<false, false><<<1, 256, 0, st>>>(hp, up, n, n_occ, p_begin, p_end, out6);
    QS_LAUNCH_CHECK();
    return QS_OK;
}

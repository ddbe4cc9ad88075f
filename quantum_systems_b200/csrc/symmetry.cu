// Symmetries of the two-body tensor that the four-index transform can exploit (csrc/transform.cu):
//
//   antisymmetry in the last pair   u[p,q,r,s] = -u[p,q,s,r]   (every anti-symmetrised tensor, basis_set.py:776-778)
//   particle-exchange symmetry      u[p,q,r,s] =  u[q,p,s,r]   (every physical interaction; RandomBasisSet builds it
//                                                               in, random_basis.py:40-42)
//
// Both survive the basis change u' = (C~ x C~) u (C x C) for ANY C, C~, already half way through: after the two
// ket contractions T2[r,s,a,b] is antisymmetric in (r,s), resp. T2[r,s,a,b] = T2[s,r,b,a].  So the second, third
// and fourth quarter steps only need the pairs r < s (r <= s), and the other half of the result is its mirror image.
//
//   qs_two_body_symmetry : EXACT test of both properties on the device (every element read once, blocks stop early
//                          once a counter-example is known) -- a promise such as BasisSet's anti_symmetrized_u flag
//                          is never trusted for skipping work.
//   qs_mirror_fill       : completes a result of which only r < s (r <= s) was computed.
#include "common.cuh"

namespace {

constexpr int kTile = 32;

template <bool COMPLEX>
struct Val {
    double re, im;
    static __device__ __forceinline__ Val load(const double* u, long long idx) {
        if (COMPLEX) {
            const double2 v = reinterpret_cast<const double2*>(u)[idx];
            return {v.x, v.y};
        }
        return {u[idx], 0.0};
    }
};

// linear index over tile pairs (tr <= ts) of a tiles x tiles grid
__device__ __forceinline__ void tile_pair(int linear, int tiles, int& tr, int& ts) {
    tr = 0;
    int rem = linear;
    while (rem >= tiles - tr) {
        rem -= tiles - tr;
        ++tr;
    }
    ts = tr + rem;
}

// MODE 1: u[p,q,r,s] == -u[p,q,s,r] for all p,q,r,s: work items (p, q, tile pair tr <= ts), every element read once.
// MODE 2: u[p,q,r,s] ==  u[q,p,s,r]: work items (p <= q, tr, ts), every element read once.
// Persistent blocks walk the work items and leave as soon as ANY block has found a counter-example, so a tensor
// without the symmetry costs a few microseconds and only a symmetric one pays the full read of u.
template <bool COMPLEX, int MODE>
__global__ void __launch_bounds__(256) symmetry_check_kernel(const double* __restrict__ u, int n, int tiles,
                                                             long long items, int* __restrict__ ok) {
    __shared__ double pre[kTile][kTile + 1], pim[COMPLEX ? kTile : 1][kTile + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int per_plane = MODE == 1 ? tiles * (tiles + 1) / 2 : tiles * tiles;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        // block-uniform early exit (one thread polls the flag); also fences the shared tile of the previous item
        if (__syncthreads_or(tx == 0 && ty == 0 && *reinterpret_cast<volatile int*>(ok) == 0)) return;
        const int t = (int)(item % per_plane);
        const long long pq = item / per_plane;
        const int p = (int)(pq / n), q = (int)(pq % n);
        int tr, ts;
        if (MODE == 1) {
            tile_pair(t, tiles, tr, ts);
        } else {
            if (p > q) continue;
            tr = t / tiles;
            ts = t % tiles;
        }
        const int r0 = tr * kTile, s0 = ts * kTile;
        const long long plane = ((long long)p * n + q) * n * n;
        const long long partner = MODE == 1 ? plane : ((long long)q * n + p) * n * n;
        // both tiles go to registers first (8 independent loads in flight per thread), then the partner tile
        // E[i][j] = u[partner plane, s0 + i, r0 + j] is exchanged through shared memory
        Val<COMPLEX> mine[kTile / 8], theirs[kTile / 8];
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            const int i = ty + 8 * k;
            theirs[k] = (s0 + i < n && r0 + tx < n) ? Val<COMPLEX>::load(u, partner + (long long)(s0 + i) * n + r0 + tx)
                                                     : Val<COMPLEX>{0.0, 0.0};
            mine[k] = (r0 + i < n && s0 + tx < n) ? Val<COMPLEX>::load(u, plane + (long long)(r0 + i) * n + s0 + tx)
                                                   : Val<COMPLEX>{0.0, 0.0};
        }
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            pre[ty + 8 * k][tx] = theirs[k].re;
            if (COMPLEX) pim[ty + 8 * k][tx] = theirs[k].im;
        }
        __syncthreads();
        bool bad = false;
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            const int i = ty + 8 * k;
            if (r0 + i < n && s0 + tx < n) {
                const double er = pre[tx][i], ei = COMPLEX ? pim[tx][i] : 0.0;
                if (MODE == 1) bad |= (mine[k].re != -er) || (COMPLEX && mine[k].im != -ei);
                else bad |= (mine[k].re != er) || (COMPLEX && mine[k].im != ei);
            }
        }
        if (__syncthreads_or(bad)) {
            if (tx == 0 && ty == 0) *ok = 0;
            return;
        }
    }
}

// Complete out[p,q,r,s] of which only r < s (MODE 1) or r <= s (MODE 2) holds valid data:
//   MODE 1: out[p,q,s,r] = -out[p,q,r,s] (r < s), out[p,q,r,r] = 0
//   MODE 2: out[q,p,s,r] =  out[p,q,r,s] (r < s), and out[q,p,r,r] = out[p,q,r,r] for p < q
// One block reads the valid tile (tr, ts) and writes the mirrored tile (ts, tr) of kFillPlanes consecutive planes p,
// keeping the loads of the next plane in flight while it stores the current one (a block that turns over after one
// 8 KB tile has no loads in flight during its stores: 4.2 TB/s).  grid (tile pairs tr <= ts, q, ceil(m / kFillPlanes)).
constexpr int kFillPlanes = 4;

template <bool COMPLEX, int MODE>
__global__ void __launch_bounds__(256) mirror_fill_pipelined_kernel(double* __restrict__ out, int n, int tiles) {
    int tr, ts;
    tile_pair(blockIdx.x, tiles, tr, ts);
    const int q = blockIdx.y;
    const int p_first = blockIdx.z * kFillPlanes;
    const int p_last = min(p_first + kFillPlanes, n);
    __shared__ double vre[kTile][kTile + 1], vim[COMPLEX ? kTile : 1][kTile + 1];
    const int r0 = tr * kTile, s0 = ts * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const double sign = MODE == 1 ? -1.0 : 1.0;
    Val<COMPLEX> cur[kTile / 8], nxt[kTile / 8];
    auto load = [&](Val<COMPLEX>(&v)[kTile / 8], int p) {
        const long long src = ((long long)p * n + q) * n * n;
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            const int i = ty + 8 * k;
            v[k] = (r0 + i < n && s0 + tx < n) ? Val<COMPLEX>::load(out, src + (long long)(r0 + i) * n + s0 + tx)
                                               : Val<COMPLEX>{0.0, 0.0};
        }
    };
    load(cur, p_first);
    for (int p = p_first; p < p_last; ++p) {
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            vre[ty + 8 * k][tx] = cur[k].re;
            if (COMPLEX) vim[ty + 8 * k][tx] = cur[k].im;
        }
        __syncthreads();
        if (p + 1 < p_last) load(nxt, p + 1);  // in flight during the stores below
        const long long dst = MODE == 1 ? ((long long)p * n + q) * n * n : ((long long)q * n + p) * n * n;
        // target element (s, r) = (s0 + i, r0 + tx) takes source (r, s) = (r0 + tx, s0 + i), valid iff r < s
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            const int i = ty + 8 * k;
            const int sI = s0 + i, r = r0 + tx;
            if (sI >= n || r >= n) continue;
            const long long at = dst + (long long)sI * n + r;
            if (r < sI) {
                if (COMPLEX) reinterpret_cast<double2*>(out)[at] = make_double2(sign * vre[tx][i], sign * vim[tx][i]);
                else out[at] = sign * vre[tx][i];
            } else if (MODE == 1 && r == sI) {
                if (COMPLEX) reinterpret_cast<double2*>(out)[at] = make_double2(0.0, 0.0);
                else out[at] = 0.0;
            } else if (MODE == 2 && r == sI && p < q) {
                // both out[p,q,r,r] and out[q,p,r,r] were computed; keep one so that the result is EXACTLY symmetric
                if (COMPLEX) reinterpret_cast<double2*>(out)[at] = make_double2(vre[tx][i], vim[tx][i]);
                else out[at] = vre[tx][i];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) cur[k] = nxt[k];
    }
}

__host__ __device__ __forceinline__ bool cyclic_wanted(int r, int s, int m) { return qs_cyclic_wanted(r, s, m); }

// out[p,q,r,s] = -out[p,q,s,r] for every pair the cyclic rule did not compute, out[p,q,r,r] = 0.
// grid (tiles * tiles, m, ceil(planes / kFillPlanes)): the block owns target tile (tr, ts) of kFillPlanes consecutive
// planes p, reads source tile (ts, tr) and keeps the next plane's loads in flight while it stores the current one.
template <bool COMPLEX>
__global__ void __launch_bounds__(256) cyclic_fill_kernel(double* __restrict__ out, int n, int tiles, int planes) {
    const int tr = blockIdx.x / tiles, ts = blockIdx.x % tiles;
    __shared__ double vre[kTile][kTile + 1], vim[COMPLEX ? kTile : 1][kTile + 1];
    const int r0 = tr * kTile, s0 = ts * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int p_first = blockIdx.z * kFillPlanes;
    const int p_last = min(p_first + kFillPlanes, planes);
    Val<COMPLEX> cur[kTile / 8], nxt[kTile / 8];
    auto load = [&](Val<COMPLEX>(&v)[kTile / 8], int p) {
        const long long plane = ((long long)p * n + blockIdx.y) * n * n;
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            const int i = ty + 8 * k;
            v[k] = (s0 + i < n && r0 + tx < n) ? Val<COMPLEX>::load(out, plane + (long long)(s0 + i) * n + r0 + tx)
                                               : Val<COMPLEX>{0.0, 0.0};
        }
    };
    load(cur, p_first);
    for (int p = p_first; p < p_last; ++p) {
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            vre[ty + 8 * k][tx] = cur[k].re;
            if (COMPLEX) vim[ty + 8 * k][tx] = cur[k].im;
        }
        __syncthreads();
        if (p + 1 < p_last) load(nxt, p + 1);
        const long long plane = ((long long)p * n + blockIdx.y) * n * n;
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) {
            const int i = ty + 8 * k;
            const int r = r0 + i, sI = s0 + tx;
            if (r >= n || sI >= n || cyclic_wanted(r, sI, n)) continue;
            const long long at = plane + (long long)r * n + sI;
            const double re = r == sI ? 0.0 : -vre[tx][i], im = (COMPLEX && r != sI) ? -vim[tx][i] : 0.0;
            if (COMPLEX) reinterpret_cast<double2*>(out)[at] = make_double2(re, im);
            else out[at] = re;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k) cur[k] = nxt[k];
    }
}

}  // namespace

// Exact test of u[p,q,r,s] == -u[p,q,s,r] on a leading-index slab of `planes` planes (a rank's shard).
extern "C" int qs_is_antisymmetric_last_pair(const void* u, int dtype, int64_t n, int64_t planes, int* host_flag,
                                             void* device_scratch, void* stream) {
    QS_REQUIRE(host_flag && device_scratch && n > 0 && n <= 65535 && planes >= 0, "qs_is_antisymmetric_last_pair: bad arguments");
    QS_REQUIRE(dtype == QS_F64 || dtype == QS_C128, "qs_is_antisymmetric_last_pair: bad dtype");
    *host_flag = 1;
    if (planes == 0) return QS_OK;
    QS_REQUIRE(u, "qs_is_antisymmetric_last_pair: null tensor");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* ok = static_cast<int*>(device_scratch);
    const int tiles = (int)qs_ceil_div(n, kTile);
    int one = 1;
    QS_CUDA(cudaMemcpyAsync(ok, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    const long long items = planes * n * (tiles * (tiles + 1) / 2);
    const long long resident = (long long)qs_sm_count() * 8;
    const unsigned grid = (unsigned)(items < resident ? items : resident);
    const dim3 block(32, 8);
    if (dtype == QS_C128)
        symmetry_check_kernel<true, 1><<<grid, block, 0, st>>>(static_cast<const double*>(u), (int)n, tiles, items, ok);
    else
        symmetry_check_kernel<false, 1><<<grid, block, 0, st>>>(static_cast<const double*>(u), (int)n, tiles, items, ok);
    QS_LAUNCH_CHECK();
    QS_CUDA(cudaMemcpyAsync(&one, ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    QS_CUDA(cudaStreamSynchronize(st));
    *host_flag = one ? 1 : 0;
    return QS_OK;
}

// Complete `planes` planes of an (.., m, m, m) result of which only the pairs chosen by the cyclic rule were computed.
extern "C" int qs_cyclic_antisymmetric_fill(void* out, int dtype, int64_t m, int64_t planes, void* stream) {
    QS_REQUIRE(m > 0 && m <= 32767 && planes >= 0 && planes <= 65535, "qs_cyclic_antisymmetric_fill: bad arguments");
    QS_REQUIRE(dtype == QS_F64 || dtype == QS_C128, "qs_cyclic_antisymmetric_fill: bad dtype");
    if (planes == 0) return QS_OK;
    QS_REQUIRE(out, "qs_cyclic_antisymmetric_fill: null tensor");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles = (int)qs_ceil_div(m, kTile);
    const dim3 block(32, 8);
    const dim3 grid((unsigned)(tiles * tiles), (unsigned)m, (unsigned)qs_ceil_div(planes, kFillPlanes));
    int span = -1;
    qs_timing_begin(QS_FAMILY_SPIN_PASS, 1.5 * (double)planes * m * m * m * 8.0 * qs_elem_doubles(dtype), stream, &span);
    if (dtype == QS_C128)
        cyclic_fill_kernel<true><<<grid, block, 0, st>>>(static_cast<double*>(out), (int)m, tiles, (int)planes);
    else
        cyclic_fill_kernel<false><<<grid, block, 0, st>>>(static_cast<double*>(out), (int)m, tiles, (int)planes);
    QS_LAUNCH_CHECK();
    qs_timing_end(span, stream);
    return QS_OK;
}

// Host helper: 1 if the cyclic rule computes the ordered pair (r, s) of an extent-m index pair.
extern "C" int qs_cyclic_pair_wanted(int64_t r, int64_t s, int64_t m) { return cyclic_wanted((int)r, (int)s, (int)m) ? 1 : 0; }

extern "C" int qs_two_body_symmetry(const void* u, int dtype, int64_t n, int first_match, int* host_flags,
                                    void* device_scratch, void* stream) {
    QS_REQUIRE(u && host_flags && device_scratch && n > 0 && n <= 65535, "qs_two_body_symmetry: bad arguments");
    QS_REQUIRE(dtype == QS_F64 || dtype == QS_C128, "qs_two_body_symmetry: bad dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* ok = static_cast<int*>(device_scratch);
    const int tiles = (int)qs_ceil_div(n, kTile);
    const dim3 block(32, 8);
    const double* up = static_cast<const double*>(u);
    int result[2] = {1, 1};
    QS_CUDA(cudaMemcpyAsync(ok, result, sizeof(result), cudaMemcpyHostToDevice, st));
    const long long items1 = (long long)n * n * (tiles * (tiles + 1) / 2), items2 = (long long)n * n * tiles * tiles;
    const long long resident = (long long)qs_sm_count() * 8;  // 8 CTAs of 256 threads per SM
    const unsigned grid1 = (unsigned)(items1 < resident ? items1 : resident);
    const unsigned grid2 = (unsigned)(items2 < resident ? items2 : resident);
    if (dtype == QS_C128) symmetry_check_kernel<true, 1><<<grid1, block, 0, st>>>(up, (int)n, tiles, items1, ok);
    else symmetry_check_kernel<false, 1><<<grid1, block, 0, st>>>(up, (int)n, tiles, items1, ok);
    QS_LAUNCH_CHECK();
    if (first_match) {
        // the caller will use the first symmetry that holds: do not pay for the second test if the first one passed
        QS_CUDA(cudaMemcpyAsync(result, ok, sizeof(int), cudaMemcpyDeviceToHost, st));
        QS_CUDA(cudaStreamSynchronize(st));
        if (result[0]) {
            *host_flags = 1;
            return QS_OK;
        }
    }
    if (dtype == QS_C128) symmetry_check_kernel<true, 2><<<grid2, block, 0, st>>>(up, (int)n, tiles, items2, ok + 1);
    else symmetry_check_kernel<false, 2><<<grid2, block, 0, st>>>(up, (int)n, tiles, items2, ok + 1);
    QS_LAUNCH_CHECK();
    QS_CUDA(cudaMemcpyAsync(result, ok, sizeof(result), cudaMemcpyDeviceToHost, st));
    QS_CUDA(cudaStreamSynchronize(st));
    *host_flags = (result[0] ? 1 : 0) | (result[1] ? 2 : 0);
    return QS_OK;
}

// Internal (common.cuh): mode 1 antisymmetric fill, mode 2 particle-exchange fill of an (m,m,m,m) result.
int qs_mirror_fill(void* out, int dtype, int64_t m, int mode, void* stream) {
    QS_REQUIRE(out && m > 0 && m <= 65535 && (mode == 1 || mode == 2), "qs_mirror_fill: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles = (int)qs_ceil_div(m, kTile);
    const dim3 block(32, 8);
    const dim3 grid((unsigned)(tiles * (tiles + 1) / 2), (unsigned)m, (unsigned)qs_ceil_div(m, kFillPlanes));
    double* o = static_cast<double*>(out);
    int span = -1;
    qs_timing_begin(QS_FAMILY_SPIN_PASS, (double)m * m * m * m * 8.0 * qs_elem_doubles(dtype), stream, &span);
    if (dtype == QS_C128) {
        if (mode == 1) mirror_fill_pipelined_kernel<true, 1><<<grid, block, 0, st>>>(o, (int)m, tiles);
        else mirror_fill_pipelined_kernel<true, 2><<<grid, block, 0, st>>>(o, (int)m, tiles);
    } else {
        if (mode == 1) mirror_fill_pipelined_kernel<false, 1><<<grid, block, 0, st>>>(o, (int)m, tiles);
        else mirror_fill_pipelined_kernel<false, 2><<<grid, block, 0, st>>>(o, (int)m, tiles);
    }
    QS_LAUNCH_CHECK();
    qs_timing_end(span, stream);
    return QS_OK;
}

// Quarter transform: one index contraction of the four-index transform as an FP64 tensor-core GEMM.
//
//   out[w-major rotated store] = sum_k A[x, k] * M[k, w]
//
// Replaces one np.tensordot (+ transpose) of BasisSet.transform_two_body_elements
// (reference quantum_systems/basis_set.py:342-348).  Design ("quarter GEMM"):
//   * every dtype combination is lowered to ONE real GEMM:  a complex A row is 2K real numbers
//     (interleaved storage IS the [re, im] K-doubling), and the coefficient matrix is expanded to a
//     real (K', W') "image" [[Mr, Mi], [-Mi, Mr]] (4M complex product, no extra passes);
//   * A tiles (128 rows x 16 k') arrive by TMA with the 128-byte swizzle; the coefficient image
//     is pre-arranged in MMA fragment order, so it arrives by one 1-D bulk copy per stage;
//   * 4 warps (one per SM sub-partition) own a 32 x (8*NT) accumulator tile each in registers
//     and issue mma.sync m8n8k4 f64 (SASS DMMA.8x8x4; measured 16 clk issue interval per
//     sub-partition, 26 clk latency, so one warp with >= 2 accumulators saturates its pipe);
//     thread 0 doubles as the TMA producer, 3 chunks ahead, over a 4-stage full/empty mbarrier
//     ring (a 5th warp would cap the kernel at 168 registers/thread and spill); 2 CTAs per SM
//     so one CTA's epilogue overlaps the other's main loop;
//   * the epilogue writes the new index as the slowest axis (or any 2-level strided address).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBlockX = 128;    // rows of A per CTA
constexpr int kChunkK = 16;     // k' per pipeline stage (= one 128-byte swizzled row)
constexpr int kStages = 4;
constexpr int kConsumerWarps = 4;
constexpr int kThreads = kConsumerWarps * 32;
constexpr int kATileBytes = kBlockX * kChunkK * 8;  // 16 KiB

constexpr int kMaxDest = 16;    // destination buffers of a scattering store (ranks of one NVSwitch domain)

struct QuarterParams {
    const double* image;      // [tiles_w][nchunks][2][NT][32][2] doubles (this launch's tile group)
    double* out;
    // Scattering store (fused re-partition): when ndest > 0 the new index w selects the destination
    // buffer outs[w / w_inner] -- a peer GPU's memory mapped over NVLink -- and sw1 is unused.
    double* outs[kMaxDest];
    int ndest;
    uint32_t X;               // rows of A
    uint32_t Wp;              // real columns actually valid (W or 2W)
    uint32_t w_first;         // first real column of this launch's tile group
    uint32_t tiles_x;         // ceil(X / 128)
    int nchunks;              // ceil(K'/16)
    int last_halves;          // 8-k' halves of the last chunk that hold data (1 or 2)
    int tiles_w;              // column tiles in this group, each 8*NT wide
    long long stagger_clocks; // start delay of the second CTA wave (0 = none)
    // store address = (w/w_inner)*sw1 + (w%w_inner)*sw0 + x2*sx2 + x1*sx1 + x0*sx0 with
    // x = (x2 * x_mid + x1) * x_inner + x0, in OUTPUT ELEMENTS (w = w' for real output, w'/2 for complex)
    uint32_t x_inner, x_mid, w_inner;
    long long sx0, sx1, sx2, sw0, sw1;
};

__device__ __forceinline__ int row_permutation(int g) {
    // MMA row g of an 8-row group reads tile row perm(g): rows of lanes g, g+1 (same quarter-warp)
    // differ in bit 2, which makes the swizzled LDS.128 conflict-free.
    return (g >> 1) | ((g & 1) << 2);
}

template <int NT, bool COMPLEX_OUT>
__global__ void __launch_bounds__(kThreads, 2)
quarter_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ QuarterParams p) {
    constexpr int kBTileBytes = NT * 1024;  // 16 k' x 8*NT w' doubles
    constexpr int kStageBytes = kATileBytes + kBTileBytes;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * kStageBytes;  // full[s] at +8s, empty[s] at +8(S+s)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t total_tiles = p.tiles_x * (uint32_t)p.tiles_w;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (kStages + s), kConsumerWarps);
        }
        mbar_fence_init();
        prefetch_tensormap(&map_a);
    }
    __syncthreads();

    // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The chunk ring runs straight
    // across tile boundaries, so the loads of the next tile are in flight during this tile's epilogue.
    // ===== TMA producer role (thread 0): the pj-th chunk of this CTA goes to stage pj % kStages =====
    uint32_t pj = 0, ptile = blockIdx.x;
    int pc = 0;
    auto produce_next = [&]() {
        if (ptile >= total_tiles) return;
        const uint32_t s = pj % kStages;
        const uint32_t full = bar_base + 8 * s;
        if (pj >= kStages) mbar_wait(bar_base + 8 * (kStages + s), ((pj / kStages) - 1) & 1);
        const uint32_t tw = ptile % (uint32_t)p.tiles_w;
        const uint32_t px0 = (ptile / (uint32_t)p.tiles_w) * kBlockX;
        mbar_expect_tx(full, kStageBytes);
        const uint32_t dst = smem_base + s * kStageBytes;
        tma_load_2d(dst, &map_a, pc * kChunkK, (int)px0, full);
        bulk_load_1d(dst + kATileBytes, p.image + ((size_t)tw * p.nchunks + pc) * (kBTileBytes / 8), kBTileBytes, full);
        ++pj;
        if (++pc == p.nchunks) {
            pc = 0;
            ptile += gridDim.x;
        }
    };
    if (threadIdx.x == 0) {
        for (int c = 0; c < kStages - 1; ++c) produce_next();
    }

    // Co-resident CTAs (block b and b + #SMs share an SM) would otherwise run their epilogues at
    // the same time and leave the tensor pipe idle; start the second wave half a tile late.
    if (p.stagger_clocks > 0 && blockIdx.x >= (gridDim.x + 1) / 2) {
        const long long t0 = clock64();
        while (clock64() - t0 < p.stagger_clocks) __nanosleep(256);
    }

    // ===== MMA: warp w owns rows [32w, 32w+32) x all 8*NT columns of the CTA tile =====
    const int g = lane >> 2;
    const int t = lane & 3;
    const int prow = row_permutation(g);

    // A fragment address inside a stage: row r = 32*warp + 8*mt + prow (128 B per row), logical
    // 16-byte chunk (4h + t) stored at chunk ^ (r & 7) by the TMA 128-byte swizzle; r & 7 == prow.
    const uint32_t a_row_off = (uint32_t)(32 * warp + prow) * 128u;
    const uint32_t a_chunk0 = (uint32_t)((t ^ prow) << 4);        // h = 0
    const uint32_t a_chunk1 = (uint32_t)(((4 + t) ^ prow) << 4);  // h = 1
    const uint32_t b_lane_off = kATileBytes + (uint32_t)lane * 16u;

    uint32_t j = 0;  // chunks consumed so far by this CTA
    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        double acc[4][NT][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        for (int c = 0; c < p.nchunks; ++c, ++j) {
            const uint32_t s = j % kStages;
            if (threadIdx.x == 0) produce_next();
            __syncwarp();
            mbar_wait(bar_base + 8 * s, (j / kStages) & 1);
            const uint32_t stage = smem_base + s * kStageBytes;
            // the last chunk may hold <= 8 valid k': skip its all-zero second half
            const int halves = (c == p.nchunks - 1) ? p.last_halves : 2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h < halves) {
                    double2 a[4];
                    double2 b[NT];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
                        a[mt] = lds_128(stage + a_row_off + mt * 1024u + (h ? a_chunk1 : a_chunk0));
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        b[nt] = lds_128(stage + b_lane_off + (uint32_t)(h * NT + nt) * 512u);
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8 * (kStages + s));
        }

        // ===== epilogue: rotated / strided store straight from the accumulators =====
        const uint32_t tile_w = tile % (uint32_t)p.tiles_w;
        const uint32_t x0 = (tile / (uint32_t)p.tiles_w) * kBlockX;
        long long xoff[4];
        bool xok[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const uint32_t x = x0 + 32 * warp + 8 * mt + prow;
            xok[mt] = x < p.X;
            const uint32_t xq = x / p.x_inner;
            const uint32_t xr = x - xq * p.x_inner;
            const uint32_t x2 = xq / p.x_mid;
            const uint32_t x1 = xq - x2 * p.x_mid;
            xoff[mt] = (long long)x2 * p.sx2 + (long long)x1 * p.sx1 + (long long)xr * p.sx0;
        }
        const uint32_t wbase = p.w_first + tile_w * (8 * NT);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint32_t wp = wbase + 8 * nt + 2 * t;  // real column of acc[..][nt][0]; wp + 1 for [1]
            if (COMPLEX_OUT) {
                if (wp < p.Wp) {
                    const uint32_t w = wp >> 1;
                    const uint32_t wq = w / p.w_inner;
                    const uint32_t wr = w - wq * p.w_inner;
                    const long long woff = (long long)wq * p.sw1 + (long long)wr * p.sw0;
                    double* const base = p.ndest ? p.outs[wq] : p.out;
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
                        if (xok[mt])
                            *reinterpret_cast<double2*>(base + 2 * (woff + xoff[mt])) =
                                make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t w = wp + e;
                    if (w < p.Wp) {
                        const uint32_t wq = w / p.w_inner;
                        const uint32_t wr = w - wq * p.w_inner;
                        const long long woff = (long long)wq * p.sw1 + (long long)wr * p.sw0;
                        double* const base = p.ndest ? p.outs[wq] : p.out;
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt)
                            if (xok[mt]) base[woff + xoff[mt]] = acc[mt][nt][e];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// coefficient image: real (K', W') expansion of M laid out in MMA fragment order, per tile group
//   image[tile][chunk][h][nt][lane][e] = M'[16*chunk + 8h + 2t + e][w_first + 8*NT*tile + 8nt + g]
// ---------------------------------------------------------------------------------------------
struct ImageParams {
    const double* m;
    long long sk, sw;  // element strides of M[k, w]
    int m_complex, a_complex, conj;
    int Kp, Wp;        // real extents: K * (a_complex ? 2 : 1), W * (out_complex ? 2 : 1)
    int NT, nchunks, tiles_w, w_first;
    int coulomb;       // 1: M[k, w] = alpha / sqrt((m[k] - m[w])^2 + a^2), m = grid (ODQD interaction)
    double alpha, a2;
};

__device__ __forceinline__ double image_value(const ImageParams& q, int kp, int wp) {
    if (kp >= q.Kp || wp >= q.Wp) return 0.0;
    if (q.coulomb) {
        // reference quantum_dots/one_dim/one_dim_qd.py:29-32 (_shielded_coulomb)
        const double dx = q.m[kp] - q.m[wp];
        return q.alpha / sqrt(dx * dx + q.a2);
    }
    const bool out_complex = q.m_complex || q.a_complex;
    const int k = q.a_complex ? (kp >> 1) : kp;
    const int ka = q.a_complex ? (kp & 1) : 0;  // 0: real part of A column, 1: imaginary part
    const int w = out_complex ? (wp >> 1) : wp;
    const int wb = out_complex ? (wp & 1) : 0;  // 0: real part of output, 1: imaginary part
    const long long idx = (long long)k * q.sk + (long long)w * q.sw;
    double mr, mi = 0.0;
    if (q.m_complex) {
        mr = q.m[2 * idx];
        mi = q.m[2 * idx + 1];
        if (q.conj) mi = -mi;
    } else {
        mr = q.m[idx];
    }
    // (ar + i ai)(mr + i mi): re = ar mr - ai mi ; im = ar mi + ai mr
    if (ka == 0) return wb == 0 ? mr : mi;
    return wb == 0 ? -mi : mr;
}

__global__ void build_image_kernel(ImageParams q, double* __restrict__ image) {
    const long long per_chunk = 128LL * q.NT;
    const long long total = (long long)q.tiles_w * q.nchunks * per_chunk;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int e = r & 1; r >>= 1;
        const int lane = r & 31; r >>= 5;
        const int nt = r % q.NT; r /= q.NT;
        const int h = r & 1; r >>= 1;
        const int chunk = r % q.nchunks;
        const int tw = r / q.nchunks;
        const int g = lane >> 2, t = lane & 3;
        const int kp = 16 * chunk + 8 * h + 2 * t + e;
        const int wp = q.w_first + 8 * q.NT * tw + 8 * nt + g;
        image[i] = image_value(q, kp, wp);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// The W' real columns are cut into 8-wide MMA column tiles; those are dealt to CTA tiles of NT <= 8
// column tiles each, in at most two groups whose NT differ by one (e.g. W' = 400 -> 6 x NT=7 + 1 x NT=8),
// so no CTA tile carries more than 7 padding columns.  Each group is one persistent launch.
struct TileGroup {
    int NT, tiles_w, w_first;
    int64_t image_offset;  // doubles
};

struct Tiling {
    int Kp, Wp, nchunks, last_halves, ngroups;
    TileGroup group[2];
    int64_t image_doubles;
};

Tiling make_tiling(int64_t K, int64_t W, int a_dtype, int m_dtype) {
    Tiling tl;
    const bool out_complex = a_dtype == QS_C128 || m_dtype == QS_C128;
    tl.Kp = (int)(K * (a_dtype == QS_C128 ? 2 : 1));
    tl.Wp = (int)(W * (out_complex ? 2 : 1));
    tl.nchunks = (int)qs_ceil_div(tl.Kp, kChunkK);
    tl.last_halves = (tl.Kp - (tl.nchunks - 1) * kChunkK) <= 8 ? 1 : 2;
    const int col_tiles = (int)qs_ceil_div(tl.Wp, 8);
    const int cta_tiles = (int)qs_ceil_div(col_tiles, 8);
    const int base = col_tiles / cta_tiles, extra = col_tiles % cta_tiles;
    tl.ngroups = 0;
    int64_t off = 0;
    int w = 0;
    if (extra > 0) {
        tl.group[tl.ngroups++] = {base + 1, extra, w, off};
        off += (int64_t)extra * tl.nchunks * 128 * (base + 1);
        w += extra * (base + 1) * 8;
    }
    tl.group[tl.ngroups++] = {base, cta_tiles - extra, w, off};
    off += (int64_t)(cta_tiles - extra) * tl.nchunks * 128 * base;
    tl.image_doubles = off;
    return tl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int g_stagger = 1;  // QS_STAGGER=0 disables the second-wave start delay (tuning aid)

template <int NT, bool CO>
int launch_variant(const CUtensorMap& map, QuarterParams p, cudaStream_t st) {
    constexpr int smem = kStages * (kATileBytes + NT * 1024) + 2 * kStages * 8 + 1024;
    static bool configured = false;
    if (!configured) {
        QS_CUDA(cudaFuncSetAttribute(quarter_gemm_kernel<NT, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        QS_CUDA(cudaFuncSetAttribute(quarter_gemm_kernel<NT, CO>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        const char* env = getenv("QS_STAGGER");
        if (env) g_stagger = atoi(env);
        configured = true;
    }
    const int64_t total = (int64_t)p.tiles_x * p.tiles_w;
    const int64_t resident = 2LL * qs_sm_count();
    const int64_t grid = total < resident ? total : resident;
    // half a tile of tensor-pipe time when two CTAs share an SM: nchunks * (4 * NT * 4 DMMA) * 16 clk
    p.stagger_clocks = (g_stagger && grid > resident / 2) ? (long long)p.nchunks * NT * 256 : 0;
    quarter_gemm_kernel<NT, CO><<<(unsigned)grid, kThreads, smem, st>>>(map, p);
    QS_LAUNCH_CHECK();
    return QS_OK;
}

template <bool CO>
int launch_nt(int NT, const CUtensorMap& map, const QuarterParams& p, cudaStream_t st) {
    switch (NT) {
        case 1: return launch_variant<1, CO>(map, p, st);
        case 2: return launch_variant<2, CO>(map, p, st);
        case 3: return launch_variant<3, CO>(map, p, st);
        case 4: return launch_variant<4, CO>(map, p, st);
        case 5: return launch_variant<5, CO>(map, p, st);
        case 6: return launch_variant<6, CO>(map, p, st);
        case 7: return launch_variant<7, CO>(map, p, st);
        case 8: return launch_variant<8, CO>(map, p, st);
    }
    qs_set_error("internal: NT=%d out of range", NT);
    return QS_ERR_INVALID;
}

int build_image(ImageParams q, const Tiling& tl, void* image, void* stream) {
    q.Kp = tl.Kp;
    q.Wp = tl.Wp;
    q.nchunks = tl.nchunks;
    for (int gi = 0; gi < tl.ngroups; ++gi) {
        const TileGroup& gr = tl.group[gi];
        q.NT = gr.NT;
        q.tiles_w = gr.tiles_w;
        q.w_first = gr.w_first;
        const long long total = (long long)gr.tiles_w * tl.nchunks * 128 * gr.NT;
        long long blocks = qs_ceil_div(total, 256);
        const long long cap = (long long)qs_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        build_image_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            q, static_cast<double*>(image) + gr.image_offset);
        QS_LAUNCH_CHECK();
    }
    return QS_OK;
}

}  // namespace

extern "C" int qs_coeff_image_bytes(int64_t K, int64_t W, int a_dtype, int m_dtype, int64_t* bytes) {
    QS_REQUIRE(K > 0 && W > 0 && bytes, "qs_coeff_image_bytes: bad arguments");
    QS_REQUIRE(K < (1 << 24) && W < (1 << 24), "qs_coeff_image_bytes: K or W too large");
    const Tiling tl = make_tiling(K, W, a_dtype, m_dtype);
    *bytes = tl.image_doubles * 8;
    return QS_OK;
}

extern "C" int qs_build_coeff_image(const void* m, int m_dtype, int64_t m_sk, int64_t m_sw, int m_conj, int64_t K,
                                    int64_t W, int a_dtype, void* image, void* stream) {
    QS_REQUIRE(m && image && K > 0 && W > 0, "qs_build_coeff_image: bad arguments");
    const Tiling tl = make_tiling(K, W, a_dtype, m_dtype);
    ImageParams q;
    memset(&q, 0, sizeof(q));
    q.m = static_cast<const double*>(m);
    q.sk = m_sk;
    q.sw = m_sw;
    q.m_complex = m_dtype == QS_C128;
    q.a_complex = a_dtype == QS_C128;
    q.conj = m_conj;
    return build_image(q, tl, image, stream);
}

// Coefficient image of the shielded-Coulomb matrix W[p, q] on `grid` (real, Gp x Gp), never
// materialised as a dense matrix.  Internal entry used by qs_odqd_coulomb.
int qs_build_coulomb_image(const double* grid, double alpha, double a, int64_t Gp, void* image, void* stream) {
    QS_REQUIRE(grid && image && Gp > 0, "qs_build_coulomb_image: bad arguments");
    const Tiling tl = make_tiling(Gp, Gp, QS_F64, QS_F64);
    ImageParams q;
    memset(&q, 0, sizeof(q));
    q.m = grid;
    q.coulomb = 1;
    q.alpha = alpha;
    q.a2 = a * a;
    return build_image(q, tl, image, stream);
}

namespace {

int quarter_launch(const void* A, int a_dtype, int64_t X, int64_t K, int64_t lda, const void* image, int m_dtype,
                   int64_t W, void* out, void* const* out_table, int64_t n_dest, int64_t x_inner, int64_t x_mid,
                   int64_t sx0, int64_t sx1, int64_t sx2, int64_t w_inner, int64_t sw0, int64_t sw1, void* stream) {
    QS_REQUIRE(A && image && (out || out_table), "qs_quarter_transform: null pointer");
    QS_REQUIRE(X > 0 && K > 0 && W > 0 && lda >= K, "qs_quarter_transform: bad extents");
    QS_REQUIRE(X < (1LL << 31) - kBlockX, "qs_quarter_transform: X=%lld exceeds 2^31", (long long)X);
    QS_REQUIRE(x_inner > 0 && w_inner > 0 && x_inner < (1LL << 32) && w_inner < (1LL << 31) && x_mid > 0 &&
                   x_mid < (1LL << 32),
               "qs_quarter_transform: bad inner extents");
    QS_REQUIRE(n_dest >= 0 && n_dest <= kMaxDest, "qs_quarter_transform_scatter: at most %d destinations", kMaxDest);
    QS_REQUIRE(n_dest == 0 || qs_ceil_div(W, w_inner) <= n_dest,
               "qs_quarter_transform_scatter: W=%lld needs more than %lld destinations of %lld", (long long)W,
               (long long)n_dest, (long long)w_inner);
    const Tiling tl = make_tiling(K, W, a_dtype, m_dtype);
    const bool out_complex = a_dtype == QS_C128 || m_dtype == QS_C128;
    const int64_t pitch_bytes = lda * 8 * qs_elem_doubles(a_dtype);
    QS_REQUIRE(pitch_bytes % 16 == 0,
               "qs_quarter_transform: row pitch of A must be a multiple of 16 bytes (pad odd real K)");
    // A and the image are TMA / bulk-copy sources (16 bytes); the epilogue stores whole elements
    const uintptr_t out_mask = out_complex ? 15 : 7;
    QS_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & out_mask) == 0,
               "qs_quarter_transform: A and image must be 16-byte aligned, out element-aligned");
    for (int64_t d = 0; d < n_dest; ++d)
        QS_REQUIRE(out_table[d] && (reinterpret_cast<uintptr_t>(out_table[d]) & out_mask) == 0,
                   "qs_quarter_transform_scatter: destination %lld is null or misaligned", (long long)d);

    EncodeTiledFn encode = get_encode_fn();
    QS_REQUIRE(encode, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)tl.Kp, (cuuint64_t)X};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {kChunkK, kBlockX};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(A), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        qs_set_error("cuTensorMapEncodeTiled failed with CUresult %d (X=%lld K'=%d pitch=%lld)", (int)cr, (long long)X,
                     tl.Kp, (long long)pitch_bytes);
        return QS_ERR_CUDA;
    }

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int span = -1;
    qs_timing_begin(QS_FAMILY_QUARTER_GEMM, 2.0 * (double)X * tl.Kp * tl.Wp, stream, &span);
    for (int gi = 0; gi < tl.ngroups; ++gi) {
        const TileGroup& gr = tl.group[gi];
        QuarterParams p;
        memset(&p, 0, sizeof(p));
        p.image = static_cast<const double*>(image) + gr.image_offset;
        p.out = static_cast<double*>(out);
        p.ndest = (int)n_dest;
        for (int64_t d = 0; d < n_dest; ++d) p.outs[d] = static_cast<double*>(out_table[d]);
        p.X = (uint32_t)X;
        p.Wp = (uint32_t)tl.Wp;
        p.w_first = (uint32_t)gr.w_first;
        p.tiles_x = (uint32_t)qs_ceil_div(X, kBlockX);
        p.nchunks = tl.nchunks;
        p.last_halves = tl.last_halves;
        p.tiles_w = gr.tiles_w;
        p.stagger_clocks = 0;
        p.x_inner = (uint32_t)x_inner;
        p.x_mid = (uint32_t)x_mid;
        p.w_inner = (uint32_t)w_inner;
        p.sx0 = sx0;
        p.sx1 = sx1;
        p.sx2 = sx2;
        p.sw0 = sw0;
        p.sw1 = n_dest ? 0 : sw1;
        QS_REQUIRE((int64_t)p.tiles_x * p.tiles_w < (1LL << 31), "qs_quarter_transform: too many tiles");
        const int rc = out_complex ? launch_nt<true>(gr.NT, map, p, st) : launch_nt<false>(gr.NT, map, p, st);
        if (rc) return rc;
    }
    qs_timing_end(span, stream);
    return QS_OK;
}

}  // namespace

extern "C" int qs_quarter_transform(const void* A, int a_dtype, int64_t X, int64_t K, int64_t lda, const void* image,
                                    int m_dtype, int64_t W, void* out, int64_t x_inner, int64_t sx0, int64_t sx1,
                                    int64_t w_inner, int64_t sw0, int64_t sw1, void* stream) {
    return quarter_launch(A, a_dtype, X, K, lda, image, m_dtype, W, out, nullptr, 0, x_inner, 0xFFFFFFFFLL, sx0, sx1, 0,
                          w_inner, sw0, sw1, stream);
}

extern "C" int qs_quarter_transform_scatter(const void* A, int a_dtype, int64_t X, int64_t K, int64_t lda,
                                            const void* image, int m_dtype, int64_t W, void* const* host_out_table,
                                            int64_t n_dest, int64_t x_inner, int64_t x_mid, int64_t sx0, int64_t sx1,
                                            int64_t sx2, int64_t w_inner, int64_t sw0, void* stream) {
    QS_REQUIRE(host_out_table && n_dest > 0, "qs_quarter_transform_scatter: empty destination table");
    return quarter_launch(A, a_dtype, X, K, lda, image, m_dtype, W, nullptr, host_out_table, n_dest, x_inner, x_mid, sx0,
                          sx1, sx2, w_inner, sw0, 0, stream);
}

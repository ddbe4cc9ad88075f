// Two-dimensional harmonic-oscillator Coulomb elements (Anisimovas & Matulis, J. Phys.: Condens. Matter
// 10, 601 (1998)) -- replaces coulomb_ho / _get_coulomb_elements of the reference
// (quantum_dots/two_dim/coulomb_elements.py:6-92, two_dim_helper.py:250-268, :283-300).
//
// The reference evaluates, for each of the l^4 index tuples, eight nested loops of
// exp(log-factorial sums + lgamma).  Two exact identities collapse that nest:
//   * the inner four-fold sum over (l1..l4) depends on the outer indices only through
//     s14 = j1 + j4 and s23 = j2 + j3, and its constraint l1 + l2 = l3 + l4 =: L makes it a single sum
//       inner(g) = (-1)^(g2+g3) sum_L  L! Gamma(S - L + 1/2)  a_L(g1, g2)  a_L(g4, g3),   S = g1 + g2,
//     with a_L(gp, gm) the integer coefficients of (1 + x)^gp (1 - x)^gm;
//   * the outer four-fold sum factorises into two convolutions of per-orbital weights
//       w(n, m, j) = (-1)^j binom(n + |m|, n - j) / j!.
// Work per element drops from O(n^4 g^3) special-function calls to O(n^2 g) multiply-adds on tables.
// These are alternating sums with heavy cancellation (the reference, in plain FP64, is 2e-10..4e-9 away
// from the exact rational value at l = 36), so the tables and both sums are carried in double-double
// arithmetic (FMA-based error-free products); the result is correctly rounded to ~1 ulp.
//
// Parallelisation: only tuples with m_p + m_q = m_r + m_s are non-zero (about 1 in the width of the m
// range).  The output is zero-filled with a memset and one thread is launched per (p, q, r, slot), slot
// indexing the orbitals whose m equals m_p + m_q - m_r in an m-sorted orbital list.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// double-double arithmetic (host and device)
// ---------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};

#define QS_HD __host__ __device__ __forceinline__

QS_HD double fma_rn(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

QS_HD dd quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = a + b;
    return {s, b - (s - a)};
}

QS_HD dd two_sum(double a, double b) {
    const double s = a + b;
    const double bb = s - a;
    return {s, (a - (s - bb)) + (b - bb)};
}

QS_HD dd two_prod(double a, double b) {
    const double p = a * b;
    return {p, fma_rn(a, b, -p)};
}

QS_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}

QS_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}

QS_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}

QS_HD dd dd_neg(dd a) { return {-a.hi, -a.lo}; }

// host only: a / b by three Newton-style quotient digits
dd dd_div(dd a, dd b) {
    const double q1 = a.hi / b.hi;
    dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    const double q2 = r.hi / b.hi;
    r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
    const double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return dd_add(q, dd{q3, 0.0});
}

// ---------------------------------------------------------------------------------------------
// tables (built on the host per call, O(l n_max + S_max^2 + g_max^3) entries)
// ---------------------------------------------------------------------------------------------
struct TdhoLayout {
    int l, n_max, e_max, g_max, s_max, m_min, m_max, slots;  // slots = largest number of orbitals sharing one m
    int64_t off_n, off_m, off_order, off_mstart, off_mcount, off_norm, off_pow2, off_w, off_fg, off_coef, bytes;
};

int64_t align16(int64_t v) { return (v + 15) / 16 * 16; }

int make_layout(const int64_t* n, const int64_t* m, int64_t l, TdhoLayout* L) {
    QS_REQUIRE(n && m && l > 0 && l < 46341, "qs_tdho_coulomb: bad quantum-number arrays");
    int n_max = 0, e_max = 0, m_min = 1 << 30, m_max = -(1 << 30);
    for (int64_t p = 0; p < l; ++p) {
        QS_REQUIRE(n[p] >= 0 && llabs(m[p]) < 4096, "qs_tdho_coulomb: orbital %lld has n = %lld, m = %lld",
                   (long long)p, (long long)n[p], (long long)m[p]);
        n_max = std::max<int>(n_max, (int)n[p]);
        e_max = std::max<int>(e_max, (int)(n[p] + llabs(m[p])));
        m_min = std::min<int>(m_min, (int)m[p]);
        m_max = std::max<int>(m_max, (int)m[p]);
    }
    // coefficients of (1+x)^gp (1-x)^gm are stored as exact doubles: needs gp + gm <= 52
    QS_REQUIRE(e_max <= 13,
               "qs_tdho_coulomb: n + |m| = %d exceeds 13 (integer coefficient tables are exact in FP64 up to the "
               "14th oscillator shell, l <= 105)",
               e_max);
    L->l = (int)l;
    L->n_max = n_max;
    L->e_max = e_max;
    L->g_max = 2 * e_max;
    L->s_max = 4 * e_max;
    L->m_min = m_min;
    L->m_max = m_max;
    std::vector<int> count(m_max - m_min + 1, 0);
    for (int64_t p = 0; p < l; ++p) count[m[p] - m_min]++;
    L->slots = *std::max_element(count.begin(), count.end());
    const int mr = m_max - m_min + 1;
    int64_t off = 0;
    L->off_n = off, off = align16(off + 4 * l);
    L->off_m = off, off = align16(off + 4 * l);
    L->off_order = off, off = align16(off + 4 * l);
    L->off_mstart = off, off = align16(off + 4 * mr);
    L->off_mcount = off, off = align16(off + 4 * mr);
    L->off_norm = off, off = align16(off + 8 * l);
    L->off_pow2 = off, off = align16(off + 8 * (int64_t)(L->s_max + 1));
    L->off_w = off, off = align16(off + 16 * l * (int64_t)(n_max + 1));
    L->off_fg = off, off = align16(off + 16 * (int64_t)(L->s_max + 1) * (L->s_max + 1));
    L->off_coef = off, off = align16(off + 8 * (int64_t)(L->g_max + 1) * (L->g_max + 1) * (2 * L->g_max + 1));
    L->bytes = off;
    return QS_OK;
}

void fill_tables(const int64_t* n, const int64_t* m, const TdhoLayout& L, char* base) {
    const int l = L.l;
    int* tn = reinterpret_cast<int*>(base + L.off_n);
    int* tm = reinterpret_cast<int*>(base + L.off_m);
    int* order = reinterpret_cast<int*>(base + L.off_order);
    int* mstart = reinterpret_cast<int*>(base + L.off_mstart);
    int* mcount = reinterpret_cast<int*>(base + L.off_mcount);
    double* norm = reinterpret_cast<double*>(base + L.off_norm);
    double* pow2neg = reinterpret_cast<double*>(base + L.off_pow2);
    dd* w = reinterpret_cast<dd*>(base + L.off_w);
    dd* fg = reinterpret_cast<dd*>(base + L.off_fg);
    double* coef = reinterpret_cast<double*>(base + L.off_coef);

    const int mr = L.m_max - L.m_min + 1;
    for (int t = 0; t < mr; ++t) mcount[t] = 0;
    for (int p = 0; p < l; ++p) {
        tn[p] = (int)n[p];
        tm[p] = (int)m[p];
        mcount[tm[p] - L.m_min]++;
    }
    for (int t = 0, run = 0; t < mr; ++t) mstart[t] = run, run += mcount[t];
    std::vector<int> cursor(mstart, mstart + mr);
    for (int p = 0; p < l; ++p) order[cursor[tm[p] - L.m_min]++] = p;  // stable: ascending p inside one m

    // k! and Gamma(k + 1/2) / sqrt(pi) = prod_{i<k} (i + 1/2), in double-double
    std::vector<dd> fact(L.s_max + 1), ghalf(L.s_max + 1);
    fact[0] = {1.0, 0.0};
    ghalf[0] = {1.0, 0.0};
    for (int k = 1; k <= L.s_max; ++k) {
        fact[k] = dd_mul_d(fact[k - 1], (double)k);
        ghalf[k] = dd_mul_d(ghalf[k - 1], (double)k - 0.5);
    }
    const int S1 = L.s_max + 1;
    for (int S = 0; S <= L.s_max; ++S) pow2neg[S] = ldexp(1.0, -S);
    for (int S = 0; S <= L.s_max; ++S)
        for (int lam = 0; lam <= L.s_max; ++lam)
            fg[S * S1 + lam] = lam <= S ? dd_mul(fact[lam], ghalf[S - lam]) : dd{0.0, 0.0};

    // per-orbital weights w(n, m, j) = (-1)^j binom(n + |m|, n - j) / j!  and  sqrt(n! / (n + |m|)!)
    for (int p = 0; p < l; ++p) {
        const int np_ = tn[p], am = abs(tm[p]);
        for (int j = 0; j <= L.n_max; ++j) {
            dd v = {0.0, 0.0};
            if (j <= np_) {
                // binom(n + |m|, n - j) = (n+|m|)! / ((n-j)! (j+|m|)!)
                v = dd_div(fact[np_ + am], dd_mul(fact[np_ - j], fact[j + am]));
                v = dd_div(v, fact[j]);
                if (j & 1) v = dd_neg(v);
            }
            w[p * (L.n_max + 1) + j] = v;
        }
        const dd ratio = dd_div(fact[np_], fact[np_ + am]);
        norm[p] = sqrt(ratio.hi);
    }

    // integer coefficients of (1 + x)^gp (1 - x)^gm, exact in FP64 for gp + gm <= 52
    const int G1 = L.g_max + 1, LW = 2 * L.g_max + 1;
    std::vector<double> poly(LW);
    for (int gp = 0; gp <= L.g_max; ++gp) {
        std::fill(poly.begin(), poly.end(), 0.0);
        poly[0] = 1.0;
        for (int t = 0; t < gp; ++t)
            for (int k = LW - 1; k >= 1; --k) poly[k] += poly[k - 1];
        for (int gm = 0; gm <= L.g_max; ++gm) {
            if (gm > 0)
                for (int k = LW - 1; k >= 1; --k) poly[k] -= poly[k - 1];
            double* row = coef + ((int64_t)gp * G1 + gm) * LW;
            for (int k = 0; k < LW; ++k) row[k] = poly[k];
        }
    }
}

struct TdhoParams {
    const int *n, *m, *order, *mstart, *mcount;
    const double *norm, *pow2neg;
    const dd *w, *fg;
    const double* coef;
    int l, n_max, g_max, s_max, m_min, m_max, slots;
    long long p_begin, planes;
    double scale;  // sqrt(pi / 2) * caller's factor
};

constexpr int kMaxConv = 32;  // >= 2 n_max + 1 (n_max <= 13)

QS_HD int up_part(int m) { return m > 0 ? m : 0; }
QS_HD int down_part(int m) { return m < 0 ? -m : 0; }
QS_HD int imax(int a, int b) { return a > b ? a : b; }
QS_HD int imin(int a, int b) { return a < b ? a : b; }

// One matrix element for orbitals (i, j, l, k) in the argument order of coulomb_elements.py:7.
QS_HD double tdho_element(const TdhoParams& P, int oi, int oj, int ol, int ok) {
    const int n_i = P.n[oi], n_j = P.n[oj], n_k = P.n[ok], n_l = P.n[ol];
    const int m_i = P.m[oi], m_j = P.m[oj], m_k = P.m[ok], m_l = P.m[ol];
    const int W1 = P.n_max + 1;
    const dd* w_i = P.w + oi * W1;
    const dd* w_j = P.w + oj * W1;
    const dd* w_k = P.w + ok * W1;
    const dd* w_l = P.w + ol * W1;

    // convolution of the weights of the (j, k) pair: A23[s23] = sum_{j2 + j3 = s23} w_j(j2) w_k(j3)
    dd a23[kMaxConv];
    for (int s23 = 0; s23 <= n_j + n_k; ++s23) {
        dd acc = {0.0, 0.0};
        for (int j2 = imax(0, s23 - n_k); j2 <= imin(n_j, s23); ++j2)
            acc = dd_add(acc, dd_mul(w_j[j2], w_k[s23 - j2]));
        a23[s23] = acc;
    }

    const int g1_0 = up_part(m_i) + down_part(m_l);  // coulomb_elements.py:49-52 with j = 0
    const int g2_0 = up_part(m_j) + down_part(m_k);
    const int g3_0 = up_part(m_k) + down_part(m_j);
    const int g4_0 = up_part(m_l) + down_part(m_i);
    const int G1 = P.g_max + 1, LW = 2 * P.g_max + 1, S1 = P.s_max + 1;

    dd element = {0.0, 0.0};
    for (int s14 = 0; s14 <= n_i + n_l; ++s14) {
        dd a14 = {0.0, 0.0};
        for (int j1 = imax(0, s14 - n_l); j1 <= imin(n_i, s14); ++j1)
            a14 = dd_add(a14, dd_mul(w_i[j1], w_l[s14 - j1]));
        const int g1 = g1_0 + s14, g4 = g4_0 + s14;
        for (int s23 = 0; s23 <= n_j + n_k; ++s23) {
            const int g2 = g2_0 + s23, g3 = g3_0 + s23;
            const int S = g1 + g2;  // == g3 + g4 by m conservation
            const double* left = P.coef + ((long long)g1 * G1 + g2) * LW;
            const double* right = P.coef + ((long long)g4 * G1 + g3) * LW;
            const dd* fg = P.fg + S * S1;
            dd inner = {0.0, 0.0};
            for (int lam = 0; lam <= S; ++lam) {
                const dd c = two_prod(left[lam], right[lam]);
                inner = dd_add(inner, dd_mul(fg[lam], c));
            }
            // (-1)^(g2+g3) 2^-S: the 2^-(G+1)/2 of coulomb_elements.py:115-117 with G = 2S (1/sqrt(2) is in scale)
            const double factor = ((g2 + g3) & 1 ? -1.0 : 1.0) * P.pow2neg[S];
            element = dd_add(element, dd_mul(dd_mul(a14, a23[s23]), dd_mul_d(inner, factor)));
        }
    }
    return (element.hi + element.lo) * (P.norm[oi] * P.norm[oj]) * (P.norm[ok] * P.norm[ol]) * P.scale;
}

__global__ void __launch_bounds__(128) tdho_coulomb_kernel(TdhoParams P, double* __restrict__ u) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = P.planes * P.l * (long long)P.l * P.slots;
    if (tid >= total) return;
    const int slot = (int)(tid % P.slots);
    long long rest = tid / P.slots;
    const int r = (int)(rest % P.l);
    rest /= P.l;
    const int q = (int)(rest % P.l);
    const long long p_local = rest / P.l;
    const int p = (int)(P.p_begin + p_local);

    // m conservation (coulomb_elements.py:19-20): m_p + m_q = m_r + m_s
    const int m_target = P.m[p] + P.m[q] - P.m[r];
    if (m_target < P.m_min || m_target > P.m_max) return;
    if (slot >= P.mcount[m_target - P.m_min]) return;
    const int s = P.order[P.mstart[m_target - P.m_min] + slot];

    // the reference passes (p, q, r, s) as (i, j, l, k): two_dim_helper.py:264-266 vs coulomb_elements.py:7
    u[((p_local * P.l + q) * P.l + r) * P.l + s] = tdho_element(P, p, q, r, s);
}

}  // namespace

extern "C" int qs_tdho_coulomb_workspace_bytes(const int64_t* host_n, const int64_t* host_m, int64_t l,
                                               int64_t* bytes) {
    QS_REQUIRE(bytes, "qs_tdho_coulomb_workspace_bytes: null pointer");
    TdhoLayout L;
    const int rc = make_layout(host_n, host_m, l, &L);
    if (rc != QS_OK) return rc;
    *bytes = L.bytes;
    return QS_OK;
}

extern "C" int qs_tdho_coulomb(const int64_t* host_n, const int64_t* host_m, int64_t l, double scale,
                               double* u_out, int64_t p_begin, int64_t p_end, void* workspace,
                               int64_t workspace_bytes, void* stream) {
    TdhoLayout L;
    const int rc = make_layout(host_n, host_m, l, &L);
    if (rc != QS_OK) return rc;
    QS_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= l, "qs_tdho_coulomb: bad plane range");
    if (p_end == p_begin) return QS_OK;  // an empty shard
    QS_REQUIRE(u_out && workspace, "qs_tdho_coulomb: null pointer");
    QS_REQUIRE(workspace_bytes >= L.bytes, "qs_tdho_coulomb: workspace of %lld bytes, %lld needed",
               (long long)workspace_bytes, (long long)L.bytes);
    QS_REQUIRE(2 * L.n_max + 1 <= kMaxConv, "qs_tdho_coulomb: n_max too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    std::vector<char> host(L.bytes, 0);
    fill_tables(host_n, host_m, L, host.data());
    // pageable source: the runtime stages the bytes before returning, so `host` may go out of scope
    QS_CUDA(cudaMemcpyAsync(workspace, host.data(), L.bytes, cudaMemcpyHostToDevice, st));

    const int64_t planes = p_end - p_begin;
    QS_CUDA(cudaMemsetAsync(u_out, 0, sizeof(double) * planes * l * l * l, st));

    char* base = static_cast<char*>(workspace);
    TdhoParams P;
    P.n = reinterpret_cast<const int*>(base + L.off_n);
    P.m = reinterpret_cast<const int*>(base + L.off_m);
    P.order = reinterpret_cast<const int*>(base + L.off_order);
    P.mstart = reinterpret_cast<const int*>(base + L.off_mstart);
    P.mcount = reinterpret_cast<const int*>(base + L.off_mcount);
    P.norm = reinterpret_cast<const double*>(base + L.off_norm);
    P.pow2neg = reinterpret_cast<const double*>(base + L.off_pow2);
    P.w = reinterpret_cast<const dd*>(base + L.off_w);
    P.fg = reinterpret_cast<const dd*>(base + L.off_fg);
    P.coef = reinterpret_cast<const double*>(base + L.off_coef);
    P.l = L.l, P.n_max = L.n_max, P.g_max = L.g_max, P.s_max = L.s_max;
    P.m_min = L.m_min, P.m_max = L.m_max, P.slots = L.slots;
    P.p_begin = p_begin, P.planes = planes;
    P.scale = scale * sqrt(M_PI / 2.0);

    const long long threads = planes * l * l * (long long)L.slots;
    const long long blocks = (threads + 127) / 128;
    QS_REQUIRE(blocks < (1LL << 31), "qs_tdho_coulomb: grid too large");
    int slot = -1;
    qs_timing_begin(QS_FAMILY_TDHO, (double)threads, stream, &slot);
    tdho_coulomb_kernel<<<(unsigned)blocks, 128, 0, st>>>(P, u_out);
    QS_LAUNCH_CHECK();
    qs_timing_end(slot, stream);
    return QS_OK;
}

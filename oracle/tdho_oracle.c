/*
 * tdho_oracle.c -- CPU oracle for the two-dimensional harmonic-oscillator Coulomb elements.
 *
 * TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the algorithm the reference runs in
 *   quantum_systems/quantum_dots/two_dim/coulomb_elements.py:6-152   (coulomb_ho and its log helpers)
 *   quantum_systems/quantum_dots/two_dim/two_dim_helper.py:111-166   (get_index_p, get_indices_nm)
 *   quantum_systems/quantum_dots/two_dim/two_dim_helper.py:250-268   (_get_coulomb_elements)
 * (Anisimovas & Matulis, J. Phys.: Condens. Matter 10, 601 (1998)): the eight nested loops, every term
 * evaluated as exp(sum of log-factorials + lgamma), in the reference's order.  The reference compiles
 * this with numba `fastmath=True`; this file is compiled WITHOUT fast-math, so the two agree to the
 * rounding of the alternating sums (measured: 2e-10 absolute at l = 36, where the reference itself is
 * 2e-10 away from the exact rational value), far inside the 1e-6 of the reference's own tests
 * (tests/test_two_dim_ho.py:70-90).
 *
 * Pinned in tests/test_oracle_golden.py against (i) the reference's golden table
 * tests/dat/two_dim_quantum_dots_coulomb_elements.dat and (ii) vectors produced by importing the
 * reference (tests/golden/make_golden_tdho.py).  Nothing in quantum_systems_b200/ links or loads this.
 *
 * Build: oracle/build_oracle.py  ->  oracle/libtdho_oracle.so   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* log(n!) as a running sum of logs -- coulomb_elements.py:95-102 */
static double log_factorial(int64_t n) {
    double fac = 0.0;
    for (int64_t a = 2; a <= n; ++a) fac += log((double)a);
    return fac;
}

/* -sum_i log(j_i!) -- coulomb_elements.py:105-112 */
static double log_ratio_1(const int64_t* j) {
    double ratio = 0.0;
    for (int i = 0; i < 4; ++i) ratio -= log_factorial(j[i]);
    return ratio;
}

/* -(G+1)/2 log 2 -- coulomb_elements.py:115-117 */
static double log_ratio_2(int64_t G) { return -0.5 * (double)(G + 1) * log(2.0); }

/* sqrt(prod_i n_i! / (n_i+|m_i|)!) -- coulomb_elements.py:120-128 */
static double log_product_1(const int64_t* n, const int64_t* m) {
    double prod = 0.0;
    for (int i = 0; i < 4; ++i) {
        prod += log_factorial(n[i]);
        prod -= log_factorial(n[i] + llabs(m[i]));
    }
    return exp(0.5 * prod);
}

/* sum_i log (n_i+|m_i|)! - log (n_i-j_i)! - log (j_i+|m_i|)! -- coulomb_elements.py:131-140 */
static double log_product_2(const int64_t* n, const int64_t* m, const int64_t* j) {
    double prod = 0.0;
    for (int i = 0; i < 4; ++i) {
        prod += log_factorial(n[i] + llabs(m[i]));
        prod -= log_factorial(n[i] - j[i]);
        prod -= log_factorial(j[i] + llabs(m[i]));
    }
    return prod;
}

/* sum_i log binom(g_i, l_i) -- coulomb_elements.py:143-152 */
static double log_product_3(const int64_t* l, const int64_t* g) {
    double prod = 0.0;
    for (int i = 0; i < 4; ++i) {
        prod += log_factorial(g[i]);
        prod -= log_factorial(l[i]);
        prod -= log_factorial(g[i] - l[i]);
    }
    return prod;
}

/* <ij|u|lk> in the convention of the paper; the caller passes (p, q, r, s) as (i, j, l, k), i.e. the
 * third argument pair is "l" and the fourth "k" -- coulomb_elements.py:6-92. */
double tdho_coulomb_ho(int64_t n_i, int64_t m_i, int64_t n_j, int64_t m_j, int64_t n_l, int64_t m_l,
                       int64_t n_k, int64_t m_k) {
    if (m_i + m_j != m_k + m_l) return 0.0; /* :19-20 */

    const int64_t M_i = (llabs(m_i) + m_i) / 2, dm_i = (llabs(m_i) - m_i) / 2; /* :22-32 */
    const int64_t M_j = (llabs(m_j) + m_j) / 2, dm_j = (llabs(m_j) - m_j) / 2;
    const int64_t M_k = (llabs(m_k) + m_k) / 2, dm_k = (llabs(m_k) - m_k) / 2;
    const int64_t M_l = (llabs(m_l) + m_l) / 2, dm_l = (llabs(m_l) - m_l) / 2;

    const int64_t n[4] = {n_i, n_j, n_k, n_l}; /* :34-35 */
    const int64_t m[4] = {m_i, m_j, m_k, m_l};
    int64_t j[4], l[4], g[4];
    double element = 0.0;

    for (j[0] = 0; j[0] <= n_i; ++j[0])
        for (j[1] = 0; j[1] <= n_j; ++j[1])
            for (j[2] = 0; j[2] <= n_k; ++j[2])
                for (j[3] = 0; j[3] <= n_l; ++j[3]) {
                    g[0] = j[0] + j[3] + M_i + dm_l; /* :49-52 */
                    g[1] = j[1] + j[2] + M_j + dm_k;
                    g[2] = j[2] + j[1] + M_k + dm_j;
                    g[3] = j[3] + j[0] + M_l + dm_i;
                    const int64_t G = g[0] + g[1] + g[2] + g[3];
                    const double ratio_1 = log_ratio_1(j);
                    const double prod_2 = log_product_2(n, m, j);
                    const double ratio_2 = log_ratio_2(G);

                    double temp = 0.0;
                    for (l[0] = 0; l[0] <= g[0]; ++l[0])
                        for (l[1] = 0; l[1] <= g[1]; ++l[1])
                            for (l[2] = 0; l[2] <= g[2]; ++l[2])
                                for (l[3] = 0; l[3] <= g[3]; ++l[3]) {
                                    if (l[0] + l[1] != l[2] + l[3]) continue; /* :69-70 */
                                    const int64_t L = l[0] + l[1] + l[2] + l[3];
                                    const double sign = (double)(-2 * ((g[1] + g[2] - l[1] - l[2]) & 1) + 1);
                                    temp += sign * exp(log_product_3(l, g) + lgamma(1.0 + 0.5 * (double)L) +
                                                       lgamma(0.5 * (double)(G - L + 1))); /* :74-82 */
                                }
                    const int64_t jsum = j[0] + j[1] + j[2] + j[3];
                    element += (double)(-2 * (jsum & 1) + 1) * exp(ratio_1 + prod_2 + ratio_2) * temp; /* :84-88 */
                }
    return element * log_product_1(n, m); /* :90 */
}

/* Orbital index p -> (n, m), shells filled in order of energy, m ascending inside a shell --
 * two_dim_helper.py:136-166. */
void tdho_indices_nm(int64_t p, int64_t* n_out, int64_t* m_out) {
    int64_t previous_shell = 0, current_shell = 1, shell_counter = 1;
    while (current_shell <= p) {
        shell_counter += 1;
        previous_shell = current_shell;
        current_shell = previous_shell + shell_counter;
    }
    const int64_t width = current_shell - previous_shell;
    if ((width & 1) == 1 && p == previous_shell + width / 2) { /* the m = 0 state of an odd shell */
        *n_out = shell_counter / 2;
        *m_out = 0;
        return;
    }
    if (2 * p < 2 * previous_shell + width) {
        *n_out = p - previous_shell;
        *m_out = -((shell_counter - 1) - 2 * (*n_out));
    } else {
        *n_out = (current_shell - 1) - p;
        *m_out = (shell_counter - 1) - 2 * (*n_out);
    }
}

/* (n, m) -> p, inverse of the above -- two_dim_helper.py:111-133. */
int64_t tdho_index_p(int64_t n, int64_t m) {
    const int64_t num_shells = 2 * n + llabs(m) + 1;
    int64_t previous_shell = 0;
    for (int64_t i = 1; i < num_shells; ++i) previous_shell += i;
    const int64_t current_shell = previous_shell + num_shells;
    if (m == 0) return n == 0 ? 0 : previous_shell + (current_shell - previous_shell) / 2;
    if (m < 0) return previous_shell + n;
    return current_shell - (n + 1);
}

/* u[p,q,r,s] = coulomb_ho(nm[p], nm[q], nm[r], nm[s]) for explicit quantum-number arrays (the B-field
 * variant passes its energy-sorted table, two_dim_helper.py:283-300; the plain oscillator passes
 * get_indices_nm(p), :250-268).  Parallel over p like the reference's numba.prange. */
void tdho_coulomb_elements_nm(const int64_t* n, const int64_t* m, int64_t num_orbitals, double* u) {
    const int64_t L = num_orbitals;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int64_t p = 0; p < L; ++p)
        for (int64_t q = 0; q < L; ++q)
            for (int64_t r = 0; r < L; ++r)
                for (int64_t s = 0; s < L; ++s)
                    u[((p * L + q) * L + r) * L + s] = tdho_coulomb_ho(n[p], m[p], n[q], m[q], n[r], m[r], n[s], m[s]);
}

void tdho_coulomb_elements(int64_t num_orbitals, double* u) {
    int64_t* n = (int64_t*)malloc(sizeof(int64_t) * (size_t)num_orbitals * 2);
    int64_t* m = n + num_orbitals;
    for (int64_t p = 0; p < num_orbitals; ++p) tdho_indices_nm(p, &n[p], &m[p]);
    tdho_coulomb_elements_nm(n, m, num_orbitals, u);
    free(n);
}

/* A list of individual elements (for sampling large l without filling l^4 values). */
void tdho_coulomb_sample(const int64_t* n, const int64_t* m, const int64_t* pqrs, int64_t count, double* out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t t = 0; t < count; ++t) {
        const int64_t p = pqrs[4 * t], q = pqrs[4 * t + 1], r = pqrs[4 * t + 2], s = pqrs[4 * t + 3];
        out[t] = tdho_coulomb_ho(n[p], m[p], n[q], m[q], n[r], m[r], n[s], m[s]);
    }
}

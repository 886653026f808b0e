/*
 * CPU oracle in plain C for the all-pairs beta-diversity path.  TEST INFRASTRUCTURE ONLY:
 * used by tests/ (parity checker at sizes the NumPy oracle is too slow for) and by
 * bench.py's CPU-baseline legs.  The product library never links or calls this file.
 *
 * Restates, with the same per-term fp64 arithmetic and NumPy's pairwise summation:
 *   _nodes_by_counts + _traverse_reduce   /root/reference/skbio/diversity/_phylogenetic.pyx:73-222
 *   _unweighted_unifrac                   /root/reference/skbio/diversity/beta/_unifrac.py:340-371
 *   _weighted_unifrac                     /root/reference/skbio/diversity/beta/_unifrac.py:374-418
 *   _weighted_unifrac_normalized          /root/reference/skbio/diversity/beta/_unifrac.py:421-471
 *   _weighted_unifrac_branch_correction   /root/reference/skbio/diversity/beta/_unifrac.py:594-619
 *   scipy pdist 'braycurtis' / 'jaccard'  (third-party; definitions in scipy/spatial/distance.py)
 * Pair order is SciPy's condensed order (_pdist_callable).  Pinned against the reference's
 * outputs through oracle_np.py (tests/test_oracle.py compares the two on the goldens).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE 128) */
static double pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

/* a: (m x n) int64, zero-initialised by the caller; counts: (n x t) int64 row-major.
 * node_of_col[c] < 0 means the column is not placed. */
void orc_nodes_by_counts(int64_t n, int64_t t, int64_t m, const int64_t* counts,
                         const int32_t* node_of_col, const int32_t* parent, int64_t* a) {
    for (int64_t c = 0; c < t; ++c) {
        int32_t k = node_of_col[c];
        if (k < 0) continue;
        for (int64_t s = 0; s < n; ++s) a[(int64_t)k * n + s] = counts[s * t + c];
    }
    for (int64_t k = 0; k < m; ++k) {
        int32_t p = parent[k];
        if (p < 0) continue;
        for (int64_t s = 0; s < n; ++s) a[(int64_t)p * n + s] += a[k * n + s];
    }
}

/* at: (n x m) int64 = transpose of a (one contiguous node vector per sample) */
void orc_unweighted(int64_t n, int64_t m, const int64_t* at, const double* bl, double* out) {
#pragma omp parallel
    {
        double* tx = (double*)malloc(sizeof(double) * (size_t)m);
        double* to = (double*)malloc(sizeof(double) * (size_t)m);
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < n - 1; ++i) {
            int64_t k = i * n - (i * (i + 1)) / 2;
            const int64_t* u = at + i * m;
            for (int64_t j = i + 1; j < n; ++j, ++k) {
                const int64_t* v = at + j * m;
                for (int64_t q = 0; q < m; ++q) {
                    int pu = u[q] != 0, pv = v[q] != 0;
                    tx[q] = bl[q] * (double)(pu ^ pv);
                    to[q] = bl[q] * (double)(pu | pv);
                }
                double ub = pairwise_sum(tx, m), ob = pairwise_sum(to, m);
                out[k] = ob == 0.0 ? 0.0 : ub / ob;
            }
        }
        free(tx);
        free(to);
    }
}

/* is_tip: (m) 0/1; d: tip-to-root distances with zeros on internal nodes (normalized only) */
void orc_weighted(int64_t n, int64_t m, const int64_t* at, const double* bl,
                  const uint8_t* is_tip, int normalized, const double* d, double* out) {
    int64_t* tot = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t s = 0; s < n; ++s) {
        int64_t t = 0;
        for (int64_t q = 0; q < m; ++q)
            if (is_tip[q]) t += at[s * m + q];
        tot[s] = t;
    }
#pragma omp parallel
    {
        double* pu = (double*)malloc(sizeof(double) * (size_t)m);
        double* pv = (double*)malloc(sizeof(double) * (size_t)m);
        double* tm = (double*)malloc(sizeof(double) * (size_t)m);
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < n - 1; ++i) {
            int64_t k = i * n - (i * (i + 1)) / 2;
            const int64_t* u = at + i * m;
            for (int64_t q = 0; q < m; ++q)
                pu[q] = tot[i] > 0 ? (double)u[q] / (double)tot[i] : (double)u[q];
            for (int64_t j = i + 1; j < n; ++j, ++k) {
                if (normalized && tot[i] == 0 && tot[j] == 0) {
                    out[k] = 0.0;
                    continue;
                }
                const int64_t* v = at + j * m;
                for (int64_t q = 0; q < m; ++q) {
                    pv[q] = tot[j] > 0 ? (double)v[q] / (double)tot[j] : (double)v[q];
                    tm[q] = bl[q] * fabs(pu[q] - pv[q]);
                }
                double wu = pairwise_sum(tm, m);
                if (normalized) {
                    for (int64_t q = 0; q < m; ++q) tm[q] = d[q] * (pu[q] + pv[q]);
                    double c = pairwise_sum(tm, m);
                    out[k] = wu / c;
                } else {
                    out[k] = wu;
                }
            }
        }
        free(pu);
        free(pv);
        free(tm);
    }
    free(tot);
}

/* Faith's PD per sample: sum bl * (count > 0)  (alpha/_pd.py:19-51) */
void orc_faith_pd(int64_t n, int64_t m, const int64_t* at, const double* bl, double* out) {
    double* tm = (double*)malloc(sizeof(double) * (size_t)m);
    for (int64_t s = 0; s < n; ++s) {
        for (int64_t q = 0; q < m; ++q) tm[q] = bl[q] * (double)(at[s * m + q] > 0);
        out[s] = pairwise_sum(tm, m);
    }
    free(tm);
}

/* x: (n x f) doubles.  Bray-Curtis: sum|u-v| / sum|u+v| (0/0 -> NaN). */
void orc_braycurtis(int64_t n, int64_t f, const double* x, double* out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n - 1; ++i) {
        int64_t k = i * n - (i * (i + 1)) / 2;
        for (int64_t j = i + 1; j < n; ++j, ++k) {
            double s1 = 0.0, s2 = 0.0;
            const double* u = x + i * f;
            const double* v = x + j * f;
            for (int64_t q = 0; q < f; ++q) {
                s1 += fabs(u[q] - v[q]);
                s2 += fabs(u[q] + v[q]);
            }
            out[k] = s1 / s2;
        }
    }
}

/* Jaccard on (x > 0): #unequal / #nonzero, 0 when both rows are empty. */
void orc_jaccard(int64_t n, int64_t f, const double* x, double* out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n - 1; ++i) {
        int64_t k = i * n - (i * (i + 1)) / 2;
        for (int64_t j = i + 1; j < n; ++j, ++k) {
            int64_t ne = 0, nz = 0;
            const double* u = x + i * f;
            const double* v = x + j * f;
            for (int64_t q = 0; q < f; ++q) {
                int a = u[q] > 0.0, b = v[q] > 0.0;
                ne += a ^ b;
                nz += a | b;
            }
            out[k] = nz ? (double)ne / (double)nz : 0.0;
        }
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

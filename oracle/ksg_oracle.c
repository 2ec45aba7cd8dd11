/*
 * ksg_oracle.c -- plain-C restatement of the KSG path's arithmetic.  TEST INFRASTRUCTURE
 * ONLY: linked by nothing in syntropy_b200; loaded through oracle/ksg_oracle_c.py by tests.
 *
 * Same definitions as oracle/ksg_oracle.py (which cites the reference line by line):
 *   radius : kp-th smallest over all j of max_d |q_d - p_jd|         (syntropy/knn/utils.py:29-33)
 *   count  : #{j : max_d |q_d - p_jd| < eps}  (or <=)  + bias         (syntropy/knn/utils.py:75-84)
 *   pairs  : #{(i,j) : max_w |x[i+w] - y[j+w]| <= r}, w < m           (syntropy/knn/temporal.py:71-72,142-143)
 * Brute force in float64, OpenMP over the queries; SoA layout base[d*ld + j].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads of the brute-force loops (0: leave the OpenMP default); returns the count in effect.
 * torchrun exports OMP_NUM_THREADS=1 to every rank: callers that want the host's cores say so. */
int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static double cheb(const double* pts, int64_t ld, int dim, int64_t j, const double* q) {
    double m = 0.0;
    for (int d = 0; d < dim; ++d) {
        double a = fabs(q[d] - pts[(int64_t)d * ld + j]);
        if (a > m) m = a;
    }
    return m;
}

/* eps_out[i] = kp-th smallest distance from query i (coordinates qry[d*ldq + i]) to the n points */
void oracle_knn_radius(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                       int64_t nq, int kp, double* eps_out) {
#pragma omp parallel
    {
        double* best = (double*)malloc(sizeof(double) * (size_t)kp);
        double q[64];
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < nq; ++i) {
            for (int d = 0; d < dim; ++d) q[d] = qry[(int64_t)d * ldq + i];
            for (int r = 0; r < kp; ++r) best[r] = INFINITY;
            for (int64_t j = 0; j < n; ++j) {
                double m = cheb(pts, ld, dim, j, q);
                if (m < best[kp - 1]) {
                    int r = kp - 1;
                    while (r > 0 && best[r - 1] > m) { best[r] = best[r - 1]; --r; }
                    best[r] = m;
                }
            }
            eps_out[i] = best[kp - 1];
        }
        free(best);
    }
}

/* counts_out[i] = #{j : dist(i,j) < eps[i]} + bias   (strict != 0), <= otherwise */
void oracle_count_within(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                         int64_t nq, const double* eps, int strict, int64_t bias, int64_t* counts_out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nq; ++i) {
        double q[64];
        for (int d = 0; d < dim; ++d) q[d] = qry[(int64_t)d * ldq + i];
        int64_t c = 0;
        for (int64_t j = 0; j < n; ++j) {
            double m = cheb(pts, ld, dim, j, q);
            c += strict ? (m < eps[i]) : (m <= eps[i]);
        }
        counts_out[i] = c + bias;
    }
}

/* out[0] = small-template pairs, out[1] = big-template pairs (closed ball, ordered, self included) */
void oracle_pair_count(const double* x, const double* y, int64_t nt, int m, double r, int64_t* out) {
    int64_t small = 0, big = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : small, big)
    for (int64_t i = 0; i < nt; ++i)
        for (int64_t j = 0; j < nt; ++j) {
            int in = 1;
            for (int w = 0; w < m && in; ++w) in = fabs(x[i + w] - y[j + w]) <= r;
            if (in) {
                ++small;
                if (fabs(x[i + m] - y[j + m]) <= r) ++big;
            }
        }
    out[0] = small;
    out[1] = big;
}

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

/* ---- Minkowski p-norms, 1 <= p < inf: cKDTree's p-th-power-space arithmetic,
 * see the comment block in oracle/ksg_oracle.py.  pow() here is the libm one scipy's C++ calls. */
static double pow_dist(const double* pts, int64_t ld, int dim, int64_t j, const double* q, double p) {
    if (p == 1.0) {
        double s = 0.0;
        for (int d = 0; d < dim; ++d) s += fabs(q[d] - pts[(int64_t)d * ld + j]);
        return s;
    }
    if (p == 2.0) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int d = 0;
        for (; d + 4 <= dim; d += 4)
            for (int u = 0; u < 4; ++u) {
                double t = q[d + u] - pts[(int64_t)(d + u) * ld + j];
                acc[u] += t * t;
            }
        double s = acc[0] + acc[1] + acc[2] + acc[3];
        for (; d < dim; ++d) {
            double t = q[d] - pts[(int64_t)d * ld + j];
            s += t * t;
        }
        return s;
    }
    double s = 0.0;
    for (int d = 0; d < dim; ++d) s += pow(fabs(q[d] - pts[(int64_t)d * ld + j]), p);
    return s;
}
static double root_p(double s, double p) { return p == 1.0 ? s : (p == 2.0 ? sqrt(s) : pow(s, 1.0 / p)); }
static double bound_p(double r, double p) { return p == 1.0 ? r : (p == 2.0 ? r * r : pow(r, p)); }

/* eps_out[i] = distance to the kp-th nearest point; idx_out (optional, [nq][kp]): the kp nearest, ascending
 * by (p-th-power distance, index) */
void oracle_knn_radius_p(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                         int64_t nq, int kp, double p, double* eps_out, int64_t* idx_out) {
#pragma omp parallel
    {
        double* best = (double*)malloc(sizeof(double) * (size_t)kp);
        int64_t* bidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)kp);
        double q[64];
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < nq; ++i) {
            for (int d = 0; d < dim; ++d) q[d] = qry[(int64_t)d * ldq + i];
            for (int r = 0; r < kp; ++r) { best[r] = INFINITY; bidx[r] = n; }
            for (int64_t j = 0; j < n; ++j) {
                double m = pow_dist(pts, ld, dim, j, q, p);
                if (m < best[kp - 1]) {
                    int r = kp - 1;
                    while (r > 0 && best[r - 1] > m) { best[r] = best[r - 1]; bidx[r] = bidx[r - 1]; --r; }
                    best[r] = m; bidx[r] = j;
                }
            }
            eps_out[i] = root_p(best[kp - 1], p);
            if (idx_out) for (int r = 0; r < kp; ++r) idx_out[i * kp + r] = bidx[r];
        }
        free(best); free(bidx);
    }
}

/* counts_out[i] = #{j : s_ij <= bound_p(r_i)} + bias, r_i = nextafter(eps_i, -inf) when strict */
void oracle_count_within_p(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                           int64_t nq, const double* eps, int strict, int64_t bias, double p, int64_t* counts_out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nq; ++i) {
        double q[64];
        for (int d = 0; d < dim; ++d) q[d] = qry[(int64_t)d * ldq + i];
        const double ub = bound_p(strict ? nextafter(eps[i], -INFINITY) : eps[i], p);
        int64_t c = 0;
        for (int64_t j = 0; j < n; ++j) c += pow_dist(pts, ld, dim, j, q, p) <= ub;
        counts_out[i] = c + bias;
    }
}

/* ordered pairs (i, j) of the columns of u (nu) and v (nv), both [dim][ld], with s_ij <= bound_p(r) */
int64_t oracle_pair_count_p(const double* u, int64_t ldu, int64_t nu, const double* v, int64_t ldv, int64_t nv,
                            int dim, double r, double p) {
    const double ub = bound_p(r, p);
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int64_t i = 0; i < nu; ++i) {
        double q[64];
        for (int d = 0; d < dim; ++d) q[d] = u[(int64_t)d * ldu + i];
        for (int64_t j = 0; j < nv; ++j) total += pow_dist(v, ldv, dim, j, q, p) <= ub;
    }
    return total;
}

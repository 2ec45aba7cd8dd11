// Minkowski p-norms other than the maximum norm (1 <= p < inf): exact float64 dense kernels.
//
// Reference: every `p` / `norm` argument of syntropy.knn ends in scipy's cKDTree
// (knn/utils.py:33,81; knn/shannon.py:401-402; knn/temporal.py:71-72,142-143), which works in
// "p-th power space" (published algorithm; its CPU restatement is pinned on the installed scipy by the tests):
//     p = 1    s = sum_d |x_d - y_d|          sequential over d                 distance = s
//     p = 2    s = sum_d (x_d - y_d)^2        four partial sums over d mod 4, added as
//                                             ((a0 + a1) + a2) + a3, then the d % 4 remainder
//                                             sequentially                      distance = sqrt(s)
//     other p  s = sum_d pow(|x_d - y_d|, p)  sequential                        distance = pow(s, 1/p)
// and a ball of radius r holds the points with s <= r, r*r, pow(r, p).  The kernels below do
// exactly this arithmetic with the round-to-nearest intrinsics (no contraction), so radii and
// counts are bit-exact for p = 1 and p = 2; for other p the device pow() may differ from the host
// libm in the last ulps (documented tolerance).
//
// Layout and structure follow exact_f64.cuh: candidate tiles staged SoA in shared memory, one
// query per thread, blockIdx.y splits the candidate range, partial lists merged by
// knn_merge_kernel.  Dimensions are run-time values (<= KSG_MAX_DIM): this is the "next" row
// of SURVEY.md 8(f), not the hot path.
#pragma once

#include "common.cuh"
#include "exact_f64.cuh"

namespace ksg {

enum { NORM_1 = 1, NORM_2 = 2, NORM_P = 3 };

// p-th-power-space distance between q[0..nd) and the candidate whose coordinate i is c(i)
template <int NORM, class Cand>
__device__ __forceinline__ double pow_space_distance(const double* q, Cand c, int nd, double p) {
    if (NORM == NORM_1) {
        double s = 0.0;
        for (int i = 0; i < nd; ++i) s = __dadd_rn(s, fabs(__dsub_rn(q[i], c(i))));
        return s;
    }
    if (NORM == NORM_2) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = 0;
        for (; i + 4 <= nd; i += 4) {
            const double t0 = __dsub_rn(q[i], c(i)), t1 = __dsub_rn(q[i + 1], c(i + 1));
            const double t2 = __dsub_rn(q[i + 2], c(i + 2)), t3 = __dsub_rn(q[i + 3], c(i + 3));
            a0 = __dadd_rn(a0, __dmul_rn(t0, t0));
            a1 = __dadd_rn(a1, __dmul_rn(t1, t1));
            a2 = __dadd_rn(a2, __dmul_rn(t2, t2));
            a3 = __dadd_rn(a3, __dmul_rn(t3, t3));
        }
        double s = __dadd_rn(__dadd_rn(__dadd_rn(a0, a1), a2), a3);
        for (; i < nd; ++i) {
            const double t = __dsub_rn(q[i], c(i));
            s = __dadd_rn(s, __dmul_rn(t, t));
        }
        return s;
    }
    double s = 0.0;
    for (int i = 0; i < nd; ++i) s = __dadd_rn(s, pow(fabs(__dsub_rn(q[i], c(i))), p));
    return s;
}

template <int NORM>
__device__ __forceinline__ double pow_space_root(double s, double p) {
    if (NORM == NORM_1) return s;
    if (NORM == NORM_2) return __dsqrt_rn(s);
    return pow(s, __ddiv_rn(1.0, p));
}
template <int NORM>
__device__ __forceinline__ double pow_space_bound(double r, double p) {
    if (NORM == NORM_1) return r;
    if (NORM == NORM_2) return __dmul_rn(r, r);
    return pow(r, p);
}

// ---- k-NN: part_dist[(split * kp + r) * nq + q] = r-th smallest POWER-SPACE distance of this split
// (ties by candidate index, as the max-norm kernel); knn_merge_kernel merges, root_kernel finishes.
template <int NORM, bool WITH_IDX>
static __global__ void __launch_bounds__(F64_THREADS)
knn_pow_kernel(const double* __restrict__ pts, int64_t ld, int64_t n, const double* __restrict__ qry, int64_t ldq,
               int64_t q_begin, int64_t nq, int dim, int kp, double p, int64_t cand_per_split,
               double* __restrict__ part_dist, int64_t* __restrict__ part_idx) {
    extern __shared__ double smem_f64[];
    double* tile = smem_f64;                       // dim * F64_TILE
    double* lst = tile + dim * F64_TILE;           // kp * F64_THREADS
    int64_t* lsti = (int64_t*)(lst + kp * F64_THREADS);
    const int tid = threadIdx.x;
    const int64_t qi = (int64_t)blockIdx.x * F64_THREADS + tid;
    const bool active = qi < nq;
    double q[KSG_MAX_DIM];
    for (int d = 0; d < dim; ++d) q[d] = active ? qry[(int64_t)d * ldq + q_begin + qi] : 0.0;
    for (int r = 0; r < kp; ++r) {
        lst[r * F64_THREADS + tid] = CUDART_INF;
        if (WITH_IDX) lsti[r * F64_THREADS + tid] = n;
    }
    double thr = CUDART_INF;
    const int64_t c0 = (int64_t)blockIdx.y * cand_per_split;
    const int64_t c1 = min(n, c0 + cand_per_split);
    for (int64_t base = c0; base < c1; base += F64_TILE) {
        const int64_t cnt = min((int64_t)F64_TILE, c1 - base);
        __syncthreads();
        stage_tile_f64(tile, pts, ld, base, cnt, dim, F64_TILE, CUDART_INF);
        __syncthreads();
        for (int j = 0; j < (int)cnt; ++j) {
            const double m = pow_space_distance<NORM>(q, [&](int i) { return tile[i * F64_TILE + j]; }, dim, p);
            if (m < thr) {  // strict: a candidate tying the current kp-th value cannot change it
                int r = kp - 1;
                while (r > 0 && lst[(r - 1) * F64_THREADS + tid] > m) {
                    lst[r * F64_THREADS + tid] = lst[(r - 1) * F64_THREADS + tid];
                    if (WITH_IDX) lsti[r * F64_THREADS + tid] = lsti[(r - 1) * F64_THREADS + tid];
                    --r;
                }
                lst[r * F64_THREADS + tid] = m;
                if (WITH_IDX) lsti[r * F64_THREADS + tid] = base + j;
                thr = lst[(kp - 1) * F64_THREADS + tid];
            }
        }
    }
    if (active) {
        for (int r = 0; r < kp; ++r) {
            const int64_t o = ((int64_t)blockIdx.y * kp + r) * nq + qi;
            part_dist[o] = lst[r * F64_THREADS + tid];
            if (WITH_IDX) part_idx[o] = lsti[r * F64_THREADS + tid];
        }
    }
}

template <int NORM>
static __global__ void pow_root_kernel(double* __restrict__ eps, int64_t nq, double p) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi < nq) eps[qi] = pow_space_root<NORM>(eps[qi], p);
}

// ---- counts: counts[s * nq + q] += #{candidates of this split with s_sub(q, c) <= bound(r_q)},
// r_q = nextafter(eps_q, -inf) when strict (knn/utils.py:75-76), eps per query or per (subspace, query).
struct SubDims {
    int8_t n[KSG_MAX_SUBSPACES];
    int8_t d[KSG_MAX_SUBSPACES][KSG_MAX_DIM];   // coordinates of each subspace, in the caller's row order
};

template <int NORM>
static __global__ void __launch_bounds__(F64_THREADS)
count_pow_kernel(const double* __restrict__ pts, int64_t ld, int64_t n, const double* __restrict__ qry, int64_t ldq,
                 int64_t q_begin, int64_t nq, int dim, const __grid_constant__ SubDims subs, int n_sub,
                 const double* __restrict__ eps, int64_t eps_stride, int strict, double p, int64_t cand_per_split,
                 int32_t* __restrict__ counts) {
    extern __shared__ double smem_f64[];
    double* tile = smem_f64;
    const int tid = threadIdx.x;
    const int64_t qi = (int64_t)blockIdx.x * F64_THREADS + tid;
    const bool active = qi < nq;
    const int s = blockIdx.z;  // one subspace per grid slice
    const int nd = subs.n[s];
    double q[KSG_MAX_DIM];
    for (int i = 0; i < nd; ++i) q[i] = active ? qry[(int64_t)subs.d[s][i] * ldq + q_begin + qi] : 0.0;
    double ub = -1.0;  // nothing is inside (power-space distances are >= 0)
    if (active) {
        double r = eps[(int64_t)s * eps_stride + qi];
        if (strict) r = nextafter(r, -CUDART_INF);
        // (eps == 0 and strict: r is the negative denormal next to 0 -- empty ball for p = 1, but
        //  r * r = +0 for p = 2, which keeps exact duplicates, exactly as cKDTree does)
        ub = pow_space_bound<NORM>(r, p);
    }
    int cnt = 0;
    const int64_t c0 = (int64_t)blockIdx.y * cand_per_split;
    const int64_t c1 = min(n, c0 + cand_per_split);
    for (int64_t base = c0; base < c1; base += F64_TILE) {
        const int jn = (int)min((int64_t)F64_TILE, c1 - base);
        __syncthreads();
        stage_tile_f64(tile, pts, ld, base, jn, dim, F64_TILE, 0.0);
        __syncthreads();
        for (int j = 0; j < jn; ++j) {
            const double m = pow_space_distance<NORM>(q, [&](int i) { return tile[subs.d[s][i] * F64_TILE + j]; }, nd, p);
            cnt += m <= ub;
        }
    }
    if (active && cnt) atomicAdd(&counts[(int64_t)s * nq + qi], cnt);
}

// ---- KSG-2: eps_s(q) = max over the k neighbours of ||x_q - x_nbr||_p in subspace s, the way
// numpy.linalg.norm(.., ord=p, axis=1) forms it on the reference's (N, d) difference array
// (knn/shannon.py:253-259): sum |.| (p = 1), sqrt(sum .^2) (p = 2), pow(sum pow(|.|, p), 1/p) --
// every sum sequential over the subspace's coordinates (the reduction runs over the strided axis
// of an F-ordered array: no pairwise blocking).
template <int NORM>
static __global__ void subspace_radius_pow_kernel(const double* __restrict__ pts, int64_t ld, int64_t q_begin,
                                                  int64_t nq, int64_t n, const int64_t* __restrict__ idx, int kp,
                                                  int skip_first, const __grid_constant__ SubDims subs, int n_sub,
                                                  double p, double* __restrict__ eps_out) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    for (int s = 0; s < n_sub; ++s) {
        double e = -CUDART_INF;
        for (int r = skip_first ? 1 : 0; r < kp; ++r) {
            const int64_t nb = idx[qi * kp + r];
            if (nb < 0 || nb >= n) { e = CUDART_INF; continue; }
            double acc = 0.0;
            for (int i = 0; i < subs.n[s]; ++i) {
                const int d = subs.d[s][i];
                const double t = __dsub_rn(pts[(int64_t)d * ld + q_begin + qi], pts[(int64_t)d * ld + nb]);
                const double term = NORM == NORM_1 ? fabs(t) : (NORM == NORM_2 ? __dmul_rn(t, t) : pow(fabs(t), p));
                acc = i == 0 ? term : __dadd_rn(acc, term);
            }
            const double nrm = NORM == NORM_1 ? acc : (NORM == NORM_2 ? __dsqrt_rn(acc) : pow(acc, __ddiv_rn(1.0, p)));
            e = fmax(e, nrm);
        }
        eps_out[(int64_t)s * nq + qi] = e;
    }
}

// ---- pair counts of the temporal estimators under a p-norm: templates of m and m + 1 samples,
// ordered pairs with s <= bound(r) (cKDTree.count_neighbors(.., r, p=norm), knn/temporal.py:71-72,142-143).
// out[0] = small (m) templates, out[1] = big (m + 1) templates.
template <int NORM>
static __global__ void __launch_bounds__(PC_THREADS)
paircount_pow_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t nt, int m, double r, double p,
                     int64_t i_begin, int64_t ni, int64_t cand_per_split, unsigned long long* __restrict__ out) {
    __shared__ double ys[PC_TILE + KSG_MAX_DIM];
    __shared__ unsigned long long red[2];
    const int tid = threadIdx.x;
    const int64_t ii = (int64_t)blockIdx.x * PC_THREADS + tid;
    const bool active = ii < ni;
    double xi[KSG_MAX_DIM];
    for (int w = 0; w <= m; ++w) xi[w] = active ? x[i_begin + ii + w] : 0.0;
    if (tid < 2) red[tid] = 0ull;
    const double ub = pow_space_bound<NORM>(r, p);
    unsigned long long small_total = 0ull, big_total = 0ull;
    const int64_t c0 = (int64_t)blockIdx.y * cand_per_split;
    const int64_t c1 = min(nt, c0 + cand_per_split);
    for (int64_t base = c0; base < c1; base += PC_TILE) {
        const int jn = (int)min((int64_t)PC_TILE, c1 - base);
        __syncthreads();
        for (int i = tid; i < jn + m; i += PC_THREADS) ys[i] = y[base + i];
        __syncthreads();
        unsigned int cs = 0, cb = 0;
        for (int j = 0; j < jn; ++j) {
            // the two template sizes are separate trees in the reference: each sum in its own order
            cs += pow_space_distance<NORM>(xi, [&](int w) { return ys[j + w]; }, m, p) <= ub;
            cb += pow_space_distance<NORM>(xi, [&](int w) { return ys[j + w]; }, m + 1, p) <= ub;
        }
        small_total += cs;
        big_total += cb;
    }
    if (!active) { small_total = 0ull; big_total = 0ull; }
    for (int off = 16; off > 0; off >>= 1) {
        small_total += __shfl_down_sync(0xffffffffu, small_total, off);
        big_total += __shfl_down_sync(0xffffffffu, big_total, off);
    }
    __syncthreads();
    if ((tid & 31) == 0) {
        atomicAdd(&red[0], small_total);
        atomicAdd(&red[1], big_total);
    }
    __syncthreads();
    if (tid == 0) {
        if (red[0]) atomicAdd(&out[0], red[0]);
        if (red[1]) atomicAdd(&out[1], red[1]);
    }
}

}  // namespace ksg

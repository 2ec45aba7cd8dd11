// Exact 1-D neighbour counts WITHOUT sorted coordinates (prepared sets with "lazy axes", axes_mode > 0).
//
// The headline step -- mutual information of two scalar series -- needs every coordinate's sorted array
// for one thing only: counts[q] = #{j : |fl64(x_q - x_j)| < eps_q}.  Two float64 radix sorts (eight
// latency-bound passes each, the one part of ksg_prepare that neither shards over ranks nor overlaps
// with the persistent k-NN kernel: it is starved until that kernel drains) are a high price for it.
// ksg_prepare already has a monotone map of every coordinate -- the curve table -- so the values are
// instead GROUPED into 2^bbits value buckets (~16 values each) by a counting sort whose histogram rides
// on curve_key_kernel (prepare.cuh: bucket_scan_kernel, bucket_scatter_kernel), and a query counts
//     whole buckets strictly inside its ball      by a difference of prefix sums,
//     the buckets its ball's two ends fall into   value by value with the exact predicate.
// Exactness.  M(v) = bucket of v is monotone non-decreasing in v (curve_pos), so M(v) < M(t) implies
// v < t and M(v) > M(t) implies v > t.  With u = 2^-53, |fl(d)| <= |d| (1 + u) and >= |d| (1 - u), so
// |x - v| <= e (1 - 2u) is inside for sure and |x - v| >= e (1 + 2u) outside for sure: with
//     A <= x - e (1 + 2u),  A' >= x - e (1 - 2u),  B' <= x + e (1 - 2u),  B >= x + e (1 + 2u)
// (directed rounding below) buckets b < M(A) or b > M(B) hold no counted value, buckets M(A') < b < M(B')
// only counted ones, and the rest -- normally the two end buckets -- are read.  Tied data can fill a
// bucket with thousands of equal values: bucket_scan_kernel raises `heavy` then, this kernel reports
// every query as unsettled (n_flagged) and the caller builds the exact axes after all (ksg_prepare_axes).
#pragma once

#include "fast_count.cuh"
#include "prepare.cuh"

namespace ksg {

constexpr int BUCKET_SCAN_CAP = 8 * BUCKET_HEAVY;  // values a query reads at most (else: flagged)

// grid = (query blocks, n_axes).  sorted: [dim][ld] prepared order; bvals: [dim][ldv] values grouped by
// bucket; bstart: [dim][lds] prefix sums (2^bbits + 1 entries per coordinate); counts: [n_axes][nq].
static __global__ void __launch_bounds__(256)
axis_count_bucket_kernel(const double* __restrict__ sorted, int64_t ld, const double* __restrict__ bvals, int64_t ldv,
                         const int32_t* __restrict__ bstart, int64_t lds, const float* __restrict__ table,
                         const double* __restrict__ center, int bbits, const __grid_constant__ AxisList axes, int64_t q_begin,
                         int64_t nq, const double* __restrict__ eps, int strict, int32_t bias,
                         int32_t* __restrict__ counts, const int32_t* __restrict__ heavy,
                         int32_t* __restrict__ n_flagged) {
    __shared__ float s_tb[CURVE_TABLE];
    const int a = axes.axis[blockIdx.y];
    for (int i = threadIdx.x; i < CURVE_TABLE; i += 256) s_tb[i] = table[(int64_t)a * CURVE_TABLE + i];
    __syncthreads();
    if (*heavy) {  // tied data: not this kernel's job (uniform over the grid)
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(n_flagged, 1);
        return;
    }
    const int64_t qi = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (qi >= nq) return;
    const double x = sorted[(int64_t)a * ld + q_begin + qi], e = eps[qi], c = center[a];
    const double* __restrict__ v = bvals + (int64_t)a * ldv;
    const int32_t* __restrict__ st = bstart + (int64_t)a * lds;
    // (1 +- 8u: four times the margin the argument needs)
    const double e_hi = __dmul_ru(e, 1.0 + 8.8817841970012523e-16), e_lo = __dmul_rd(e, 1.0 - 8.8817841970012523e-16);
    // the two values of an end (A, A' or B', B) almost always fall into the same gap of the table: one search
    auto buckets_of = [&](double t1, double t2, int& b1, int& b2) {
        const float f1 = __double2float_rn(t1 - c), f2 = __double2float_rn(t2 - c);
        const int g1 = curve_gap(f1, s_tb);
        b1 = (int)curve_scale(curve_pos_in_gap(f1, g1, s_tb), bbits);
        const bool same = (g1 == 0 || s_tb[g1 - 1] <= f2) && (g1 >= CURVE_TABLE || !(s_tb[g1] <= f2));
        b2 = (int)curve_scale(curve_pos_in_gap(f2, same ? g1 : curve_gap(f2, s_tb), s_tb), bbits);
    };
    int lo1, lo2, hi1, hi2;
    buckets_of(__dsub_rd(x, e_hi), __dsub_ru(x, e_lo), lo1, lo2);
    buckets_of(__dadd_rd(x, e_lo), __dadd_ru(x, e_hi), hi1, hi2);
    int cnt = 0;
    bool too_long = false;
    auto inside = [&](double vj) -> int {
        const double d = fabs(x - vj);
        return (strict ? d < e : d <= e) ? 1 : 0;
    };
    auto scan = [&](int b0, int b1) {  // every value of the buckets [b0, b1], exact predicate
        const int s = st[b0], t = st[b1 + 1];
        if (t - s > BUCKET_SCAN_CAP) { too_long = true; return; }
        int j = s;
        for (; j + 4 <= t; j += 4) {  // (independent loads first: the loop is latency-bound otherwise)
            const double v0 = v[j], v1 = v[j + 1], v2 = v[j + 2], v3 = v[j + 3];
            cnt += inside(v0) + inside(v1) + inside(v2) + inside(v3);
        }
        for (; j < t; ++j) cnt += inside(v[j]);
    };
    if (lo2 < hi1) {
        scan(lo1, lo2);
        cnt += st[hi1] - st[lo2 + 1];
        scan(hi1, hi2);
    } else {
        scan(lo1, hi2);
    }
    if (too_long) atomicAdd(n_flagged, 1);
    counts[(int64_t)blockIdx.y * nq + qi] = cnt + bias;
}

}  // namespace ksg

// Host-side harness for axis_count_kernel (tests/test_axis_count_host.py): the 1-D marginal counts
// of the path (both marginals of the headline MI) are two binary searches with the exact float64
// predicate, the upper levels on a shared-memory sample of the sorted array.  The kernel is cut out
// of csrc/fast_count.cuh and run block by block on the CPU (its only barrier separates the table
// fill from the searches and the fill is idempotent: every block is simply executed twice), against
// a brute-force count of |fl64(x - v)| < eps (strict) / <= eps (closed).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <vector>
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
#define __shared__ static
#define __syncthreads() ((void)0)
struct Dim3 { unsigned x; };
static Dim3 blockIdx{0}, blockDim{1024}, threadIdx{0};
#include "axis_count_kernel.inc"

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 300;
    std::mt19937_64 rng(5);
    long checked = 0, wrong = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int64_t n = (trial < 8) ? 1 + trial : 1 + (int64_t)(rng() % 6000);
        const int kind = (int)(rng() % 4);  // 0 normal, 1 quantised, 2 heavy tails, 3 constant
        const int strict = (int)(rng() % 2);
        const int32_t bias = (int32_t)(rng() % 3) - 1;
        std::normal_distribution<double> nd(0.0, 1.0);
        std::cauchy_distribution<double> cd(0.0, 1.0);
        std::vector<double> row(n);  // the coordinate in the prepared order
        for (auto& v : row) {
            v = kind == 2 ? cd(rng) : nd(rng);
            if (kind == 1) v = std::round(v * 4) / 4;
            if (kind == 3) v = 1.25;
        }
        std::vector<double> vals = row;
        std::sort(vals.begin(), vals.end());
        const int64_t q_begin = (int64_t)(rng() % n);
        const int64_t nq = 1 + (int64_t)(rng() % (n - q_begin));
        std::vector<double> eps(nq);
        for (auto& e : eps) {
            const int pick = (int)(rng() % 10);
            e = pick == 0 ? 0.0 : pick == 1 ? std::numeric_limits<double>::infinity() : pick == 2 ? 0.25
                : pick == 3 ? 1e-300 : std::fabs(nd(rng)) * (pick < 7 ? 0.01 : 1.0);
        }
        // position of every sorted value in the prepared order (ties: any consistent assignment)
        std::vector<int32_t> pos(n);
        {
            std::vector<int32_t> order(n);
            for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return row[a] < row[b]; });
            for (int64_t i = 0; i < n; ++i) pos[i] = order[i];
        }
        std::vector<int32_t> counts(nq, -12345), counts_by_value(nq, -12345);
        const unsigned blocks = (unsigned)((nq + 1023) / 1024);
        for (unsigned b = 0; b < blocks; ++b)
            for (int pass = 0; pass < 2; ++pass)
                for (unsigned t = 0; t < 1024; ++t) {
                    blockIdx.x = b;
                    threadIdx.x = t;
                    axis_count_kernel(row.data(), vals.data(), n, q_begin, nq, eps.data(), strict, bias, counts.data());
                }
#ifdef HAVE_SORTED_KERNEL
        const unsigned blocks_n = (unsigned)((n + 1023) / 1024);
        for (unsigned b = 0; b < blocks_n; ++b)
            for (int pass = 0; pass < 2; ++pass)
                for (unsigned t = 0; t < 1024; ++t) {
                    blockIdx.x = b;
                    threadIdx.x = t;
                    axis_count_sorted_kernel(vals.data(), pos.data(), n, q_begin, nq, eps.data(), strict, bias,
                                             counts_by_value.data());
                }
        for (int64_t qi = 0; qi < nq; ++qi)
            if (counts_by_value[qi] != counts[qi] && ++wrong < 10)
                printf("WRONG (by value) trial %d qi %ld: %d vs %d\n", trial, (long)qi, counts_by_value[qi], counts[qi]);
#endif
        for (int64_t qi = 0; qi < nq; ++qi) {
            const double x = row[q_begin + qi], e = eps[qi];
            int32_t want = bias;
            for (int64_t j = 0; j < n; ++j) {
                const double d = std::fabs(x - vals[j]);
                want += strict ? (d < e) : (d <= e);
            }
            ++checked;
            if (counts[qi] != want && ++wrong < 10)
                printf("WRONG trial %d qi %ld: %d want %d (n %ld eps %g strict %d kind %d)\n", trial, (long)qi, counts[qi], want,
                       (long)n, e, strict, kind);
        }
    }
    printf("checked %ld wrong %ld\n", checked, wrong);
    return wrong ? 1 : 0;
}

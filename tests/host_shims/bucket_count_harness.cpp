// Host-side harness for axis_count_bucket_kernel (tests/test_bucket_count_host.py): the bucketed exact
// 1-D counts of sets with lazy axes.  curve_pos / curve_scale / bucket_bits (csrc/prepare.cuh) and the
// kernel (csrc/axis_buckets.cuh) are cut out of the sources and compiled for the CPU; the curve table,
// the histogram, its scan and the scatter are redone here the way the device kernels do them
// (curve_table_kernel: min, a sorted sample of 1024 values, max, centred on the sample median).
// Every count the kernel settles must equal the brute-force count of |fl64(x - v)| < eps (<= when
// closed) -- ties, heavy tails, constant coordinates, offsets, radii 0 / tiny / huge / inf / NaN.
#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <vector>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
#define __grid_constant__
#define __shared__ static
#define __syncthreads() ((void)0)
#define CUDART_INF (std::numeric_limits<double>::infinity())
struct Dim3 { unsigned x, y; };
static Dim3 blockIdx{0, 0}, threadIdx{0, 0};
constexpr int KSG_MAX_DIM = 32;
static inline int atomicAdd(int32_t* p, int v) { int o = *p; *p += v; return o; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __double2float_rn(double x) { return (float)x; }
// directed rounding through the floating-point environment (compiled with -frounding-math)
template <class F>
static inline double rounded(int mode, F f) {
    const int old = fegetround();
    fesetround(mode);
    volatile double r = f();
    fesetround(old);
    return r;
}
static inline double __dmul_ru(double a, double b) { volatile double x = a, y = b; return rounded(FE_UPWARD, [&] { return x * y; }); }
static inline double __dmul_rd(double a, double b) { volatile double x = a, y = b; return rounded(FE_DOWNWARD, [&] { return x * y; }); }
static inline double __dadd_ru(double a, double b) { volatile double x = a, y = b; return rounded(FE_UPWARD, [&] { return x + y; }); }
static inline double __dadd_rd(double a, double b) { volatile double x = a, y = b; return rounded(FE_DOWNWARD, [&] { return x + y; }); }
static inline double __dsub_ru(double a, double b) { volatile double x = a, y = b; return rounded(FE_UPWARD, [&] { return x - y; }); }
static inline double __dsub_rd(double a, double b) { volatile double x = a, y = b; return rounded(FE_DOWNWARD, [&] { return x - y; }); }
using std::fabs;
struct AxisList { int32_t axis[KSG_MAX_DIM]; };
constexpr int CURVE_SAMPLES = 1024;
constexpr int CURVE_TABLE = CURVE_SAMPLES + 2;
#include "bucket_kernel.inc"

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 200;
    std::mt19937_64 rng(11);
    long checked = 0, wrong = 0, flagged_runs = 0, heavy_runs = 0, scanned = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int64_t n = trial < 6 ? 1 + trial : 40 + (int64_t)(rng() % 30000);
        const int kind = (int)(rng() % 6);  // 0 normal, 1 quantised, 2 heavy tails, 3 constant, 4 offset + tiny scale, 5 few values
        const int strict = (int)(rng() % 2);
        const int32_t bias = (int32_t)(rng() % 3) - 1;
        std::normal_distribution<double> nd(0.0, 1.0);
        std::cauchy_distribution<double> cd(0.0, 1.0);
        std::vector<double> row(n);
        for (auto& v : row) {
            v = kind == 2 ? cd(rng) : nd(rng);
            if (kind == 1) v = std::round(v * 64) / 64;
            if (kind == 3) v = 1.25;
            if (kind == 4) v = 1e6 + v * 1e-7;
            if (kind == 5) v = (double)(rng() % 7);
        }
        // ---- curve table as curve_table_kernel builds it
        std::vector<double> sample(CURVE_SAMPLES);
        for (int s = 0; s < CURVE_SAMPLES; ++s) {
            const uint64_t lo = (uint64_t)s * (uint64_t)n / CURVE_SAMPLES, hi = (uint64_t)(s + 1) * (uint64_t)n / CURVE_SAMPLES;
            uint64_t e = lo + (hi > lo ? rng() % (hi - lo) : 0);
            if (e >= (uint64_t)n) e = (uint64_t)n - 1;
            sample[s] = row[e];
        }
        std::sort(sample.begin(), sample.end());
        const double c = sample[CURVE_SAMPLES / 2];
        const double mn = *std::min_element(row.begin(), row.end()), mx = *std::max_element(row.begin(), row.end());
        std::vector<float> table(CURVE_TABLE);
        table[0] = (float)(mn - c);
        table[CURVE_TABLE - 1] = (float)(mx - c);
        for (int s = 0; s < CURVE_SAMPLES; ++s) table[1 + s] = (float)(sample[s] - c);
        // ---- histogram, scan, scatter (curve_key_kernel / bucket_scan_kernel / bucket_scatter_kernel)
        const int bbits = bucket_bits(n);
        const int B = 1 << bbits;
        std::vector<uint32_t> bid(n);
        std::vector<int32_t> start(B + 1, 0), cursor(B, 0);
        int occupancy = 0;
        for (int64_t i = 0; i < n; ++i) {
            bid[i] = curve_scale(curve_pos((float)(row[i] - c), table.data()), bbits);
            ++cursor[bid[i]];
        }
        for (int b = 0, run = 0; b < B; ++b) { const int cnt = cursor[b]; occupancy = std::max(occupancy, cnt); start[b] = run; cursor[b] = run; run += cnt; start[b + 1] = run; }
        std::vector<double> grouped(n);
        std::vector<int64_t> order(n);
        for (int64_t i = 0; i < n; ++i) order[i] = i;
        std::shuffle(order.begin(), order.end(), rng);  // (the device's order inside a bucket is arbitrary)
        for (int64_t i : order) grouped[cursor[bid[i]]++] = row[i];
        int32_t heavy = occupancy > BUCKET_HEAVY ? 1 : 0;
        if (trial % 7 == 3) heavy = 0;  // also exercise long scans (and the per-query cap) on tied data
        heavy_runs += heavy;
        // ---- queries: every point, radii of all kinds
        const int64_t q_begin = (int64_t)(rng() % n);
        const int64_t nq = 1 + (int64_t)(rng() % (n - q_begin));
        std::vector<double> eps(nq);
        for (auto& e : eps) {
            const int pick = (int)(rng() % 12);
            const double scale = kind == 4 ? 1e-7 : 1.0;
            e = pick == 0 ? 0.0 : pick == 1 ? std::numeric_limits<double>::infinity() : pick == 2 ? 0.25 * scale
              : pick == 3 ? std::numeric_limits<double>::quiet_NaN() : pick == 4 ? 1e-300 : pick == 5 ? 1e300
              : std::fabs(nd(rng)) * scale * (pick < 9 ? 0.01 : 0.5);
            if (pick == 6 && n > 1) e = std::fabs(row[rng() % n] - row[rng() % n]);  // radii that ARE pair distances
        }
        std::vector<int32_t> counts(nq, -777);
        int32_t n_flagged = 0;
        AxisList axes{};
        axes.axis[0] = 0;
        // (the kernel's only barrier separates the table fill from its use and the fill is idempotent:
        //  every block is simply executed twice, thread by thread)
        for (int64_t blk = 0; blk * 256 < nq; ++blk)
            for (int pass = 0; pass < 2; ++pass) {
                int32_t flagged_pass = 0;
                for (unsigned t = 0; t < 256; ++t) {
                    blockIdx.x = (unsigned)blk; blockIdx.y = 0; threadIdx.x = t;
                    axis_count_bucket_kernel(row.data(), n, grouped.data(), n, start.data(), B + 1, table.data(), &c, bbits,
                                             axes, q_begin, nq, eps.data(), strict, bias, counts.data(), &heavy, &flagged_pass);
                }
                if (pass == 1) n_flagged += flagged_pass;
            }
        if (n_flagged) { ++flagged_runs; continue; }  // (the caller falls back to the sorted arrays)
        for (int64_t qi = 0; qi < nq; ++qi) {
            const double x = row[q_begin + qi], e = eps[qi];
            int32_t want = bias;
            for (int64_t j = 0; j < n; ++j) {
                const double d = std::fabs(x - row[j]);
                want += (strict ? d < e : d <= e) ? 1 : 0;
            }
            ++checked;
            if (counts[qi] != want && ++wrong < 10)
                printf("WRONG trial %d kind %d n %ld qi %ld: got %d want %d (x %.17g eps %.17g strict %d)\n", trial, kind,
                       (long)n, (long)qi, counts[qi], want, x, e, strict);
        }
        (void)scanned;
    }
    printf("checked %ld wrong %ld flagged_runs %ld heavy_runs %ld\n", checked, wrong, flagged_runs, heavy_runs);
    return wrong ? 1 : 0;
}

// Host-side harness for count_fix_kernel (tests/test_count_fix_host.py): the exactness argument of the
// fp32 count filter, checked on the CPU.  The filter counts #{j : m32(q, j) < T_q} with m32 computed
// from the centred fp32 rows and T_q = fp32_up(eps_q + delta_q) + 1 ulp; count_fix_kernel must turn
// that into the exact #{j : m64(q, j) < eps_q} (strict) / <= eps_q (closed) for every subspace, or
// flag the query.  The kernel source is cut out of csrc/fast_count.cuh and compiled for the host.
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
#define __host__
#define __restrict__
#define __launch_bounds__(x)
#define __forceinline__ inline
#define CUDART_INF (std::numeric_limits<double>::infinity())
#define CUDART_INF_F (std::numeric_limits<float>::infinity())
struct Dim3 { unsigned x; };
static Dim3 blockIdx{0}, blockDim{128}, threadIdx{0};
static inline int atomicAdd(int32_t* p, int v) { int o = *p; *p += v; return o; }
static inline float __double2float_ru(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafter(f, std::numeric_limits<float>::infinity());
    return f;
}
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
using std::fabs;
using std::fmax;
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
constexpr int KSG_MAX_DIM = 16;
constexpr int KSG_MAX_SUBSPACES = 64;
#include "count_fix_kernel.inc"  // band_query_f64, query_maxabs, count_threshold, COUNT_MIN_T, the searches, the kernel

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 200;
    std::mt19937_64 rng(77);
    long checked = 0, wrong = 0, flagged = 0, corrected = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int dim = 1 + (int)(rng() % 6);
        const int dp = (dim + 1) & ~1;
        const int64_t n = 300 + (int64_t)(rng() % 700);
        const int kind = (int)(rng() % 5);  // 0 normal, 1 off-centre, 2 heavy tails, 3 quantised, 4 tiny scale
        const int strict = (int)(rng() % 2);
        const int k = 1 + (int)(rng() % 6);
        std::vector<double> pts((size_t)dim * n);
        std::normal_distribution<double> nd(0.0, 1.0);
        std::cauchy_distribution<double> cd(0.0, 1.0);
        for (int d = 0; d < dim; ++d)
            for (int64_t i = 0; i < n; ++i) {
                double v = kind == 2 ? cd(rng) : nd(rng);
                if (kind == 1) v += 1e4 * (d + 1);
                if (kind == 3) v = std::round(v * 8) / 8;
                if (kind == 4) v *= 1e-12;
                pts[(size_t)d * n + i] = v;
            }
        // centre = median, shadow = fp32(x - centre), zero padding coordinate
        std::vector<double> center(dim), maxabs(dim, 0.0);
        std::vector<double> axis_vals((size_t)dim * n);
        std::vector<int32_t> axis_pos((size_t)dim * n);
        for (int d = 0; d < dim; ++d) {
            std::vector<std::pair<double, int32_t>> v(n);
            for (int64_t i = 0; i < n; ++i) v[i] = {pts[(size_t)d * n + i], (int32_t)i};
            std::sort(v.begin(), v.end());
            for (int64_t i = 0; i < n; ++i) { axis_vals[(size_t)d * n + i] = v[i].first; axis_pos[(size_t)d * n + i] = v[i].second; }
            center[d] = v[n / 2].first;
        }
        std::vector<float> shadow((size_t)n * dp, 0.0f);
        for (int64_t i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d) shadow[(size_t)i * dp + d] = (float)(pts[(size_t)d * n + i] - center[d]);
        // subspaces
        const int n_sub = 1 + (int)(rng() % 4);
        MaskTable64 masks{};
        for (int s = 0; s < n_sub; ++s) {
            uint32_t m = 0;
            while (!m) m = (uint32_t)(rng() % (1u << dim));
            masks.m[s] = m;
        }
        // radii: exact (k+1)-th smallest joint distance (self included), as the estimators use them
        std::vector<double> eps(n);
        for (int64_t q = 0; q < n; ++q) {
            std::vector<double> dist(n);
            for (int64_t j = 0; j < n; ++j) {
                double m = 0.0;
                for (int d = 0; d < dim; ++d) m = fmax(m, fabs(pts[(size_t)d * n + q] - pts[(size_t)d * n + j]));
                dist[j] = m;
            }
            std::nth_element(dist.begin(), dist.begin() + k, dist.end());
            eps[q] = dist[k];
            if (rng() % 97 == 0) eps[q] = CUDART_INF;
        }
        std::vector<float> thresholds(n);
        for (int64_t q = 0; q < n; ++q) thresholds[q] = count_threshold(eps[q], query_maxabs(shadow.data(), q, dp));
        // the filter's counts (fp32) and the truth (fp64)
        std::vector<int32_t> counts((size_t)n_sub * n, 0), truth((size_t)n_sub * n, 0);
        for (int s = 0; s < n_sub; ++s)
            for (int64_t q = 0; q < n; ++q) {
                int c32 = 0, c64 = 0;
                for (int64_t j = 0; j < n; ++j) {
                    float m32 = 0.0f;
                    double m64 = 0.0;
                    for (int d = 0; d < dim; ++d)
                        if ((masks.m[s] >> d) & 1u) {
                            m32 = std::max(m32, std::fabs(__fsub_rn(shadow[(size_t)q * dp + d], shadow[(size_t)j * dp + d])));
                            m64 = fmax(m64, fabs(pts[(size_t)d * n + q] - pts[(size_t)d * n + j]));
                        }
                    c32 += m32 < thresholds[q];
                    c64 += strict ? (m64 < eps[q]) : (m64 <= eps[q]);
                }
                counts[(size_t)s * n + q] = c32;
                truth[(size_t)s * n + q] = c64;
            }
        std::vector<int32_t> before = counts, flags(n, 0);
        int32_t n_flagged = 0;
        for (int64_t q = 0; q < n; ++q) {
            blockIdx.x = (unsigned)(q / 128);
            threadIdx.x = (unsigned)(q % 128);
            count_fix_kernel(pts.data(), n, shadow.data(), dim, dp, n, axis_vals.data(), axis_pos.data(), maxabs.data(), 0, n,
                             eps.data(), thresholds.data(), masks, n_sub, strict, counts.data(), flags.data(), &n_flagged, nullptr);
        }
        for (int64_t q = 0; q < n; ++q) {
            if (flags[q]) { ++flagged; continue; }
            for (int s = 0; s < n_sub; ++s) {
                ++checked;
                if (before[(size_t)s * n + q] != counts[(size_t)s * n + q]) ++corrected;
                if (counts[(size_t)s * n + q] != truth[(size_t)s * n + q] && ++wrong < 10)
                    printf("WRONG trial %d q %ld sub %d: %d want %d (filter %d) dim %d kind %d strict %d eps %.17g\n", trial,
                           (long)q, s, counts[(size_t)s * n + q], truth[(size_t)s * n + q], before[(size_t)s * n + q], dim, kind,
                           strict, eps[q]);
            }
        }
    }
    printf("checked %ld wrong %ld flagged %ld corrected %ld\n", checked, wrong, flagged, corrected);
    return wrong ? 1 : 0;
}

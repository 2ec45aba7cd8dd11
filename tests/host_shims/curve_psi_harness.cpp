// Host-side harness (tests/test_curve_psi_host.py) for two plain-C++ device functions:
//  * hilbert_index (csrc/prepare.cuh): on a full grid of 2^bits cells per axis the keys must be a
//    bijection onto [0, 2^(bits*dim)) and consecutive keys must be NEIGHBOURING cells (exactly one
//    coordinate changes, by one) -- the property that makes runs of consecutive rows compact blobs;
//  * psi_of_int (csrc/epilogue.cuh): prints psi at the integers given on stdin (checked against
//    scipy's digamma by the Python side).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>
#define __global__
#define __device__
#define __host__
#define __restrict__
#define __forceinline__ inline
#define CUDART_INF (std::numeric_limits<double>::infinity())
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
constexpr double EULER_GAMMA = 0.57721566490153286061;
#include "device_functions.inc"  // hilbert_index, psi_of_int (cut out of the headers)

template <int DIM>
static long check_curve(int bits) {
    const uint64_t cells = 1ull << (bits * DIM);
    std::vector<std::pair<unsigned long long, uint64_t>> keyed(cells);
    for (uint64_t c = 0; c < cells; ++c) {
        unsigned int X[DIM];
        for (int d = 0; d < DIM; ++d) X[d] = (unsigned int)((c >> (d * bits)) & ((1u << bits) - 1u));
        keyed[c] = {hilbert_index<DIM>(X, DIM, bits), c};
    }
    std::sort(keyed.begin(), keyed.end());
    long bad = 0;
    for (uint64_t i = 0; i < cells; ++i) {
        if (keyed[i].first != i) ++bad;  // bijection onto [0, cells)
        if (i == 0) continue;
        int changed = 0, step = 0;
        for (int d = 0; d < DIM; ++d) {
            const int a = (int)((keyed[i - 1].second >> (d * bits)) & ((1u << bits) - 1u));
            const int b = (int)((keyed[i].second >> (d * bits)) & ((1u << bits) - 1u));
            if (a != b) { ++changed; step = std::abs(a - b); }
        }
        if (changed != 1 || step != 1) ++bad;
    }
    return bad;
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'c') {
        long bad = 0;
        bad += check_curve<2>(1) + check_curve<2>(5) + check_curve<2>(8);
        bad += check_curve<3>(2) + check_curve<3>(5);
        bad += check_curve<4>(1) + check_curve<4>(4);
        bad += check_curve<5>(3) + check_curve<6>(3) + check_curve<8>(2);
        printf("curve_violations %ld\n", bad);
        return bad ? 1 : 0;
    }
    long long n;
    while (scanf("%lld", &n) == 1) printf("%.17g\n", psi_of_int((int64_t)n));
    return 0;
}

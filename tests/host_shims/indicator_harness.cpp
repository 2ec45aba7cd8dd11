// Host-side harness (tests/test_indicator_host.py) for the fused count kernel's indicator
// (csrc/fast_count.cuh): "m < T" evaluated as one saturating FMA  sat((-m) * B + T * B)  with a
// per-query power of two B.  Claim: for every threshold T >= 2^-100 (smaller ones are flagged) the
// result is exactly 1.0f when m < T and 0.0f otherwise -- also one ulp either side of T, for
// m = +inf, T = +inf and the NaN of inf - inf.  The PTX instruction is modelled by fmaf + clamp
// (fma.rn.sat.f32: single rounding, NaN -> +0).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#define __device__
#define __forceinline__ inline
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float fma_sat(float a, float b, float c) {
    const float r = fmaf(a, b, c);
    if (!(r == r)) return 0.0f;
    return r < 0.0f ? 0.0f : (r > 1.0f ? 1.0f : r);
}
#include "indicator.inc"  // COUNT_MIN_T, IndScale, make_ind_scale, ind_lt (asm replaced by fma_sat)

int main() {
    std::mt19937_64 rng(9);
    long checked = 0, wrong = 0;
    const float inf = std::numeric_limits<float>::infinity();
    auto check = [&](float m, float T) {
        if (!(T >= COUNT_MIN_T)) return;  // flagged for the float64 kernel
        const float got = ind_lt(m, make_ind_scale(T));
        const float want = (m < T) ? 1.0f : 0.0f;
        ++checked;
        if (got != want && ++wrong < 10) printf("WRONG m %.9g T %.9g: %g want %g\n", m, T, got, want);
    };
    for (long i = 0; i < 3000000; ++i) {
        // thresholds over the whole exponent range above 2^-100
        const int e = 27 + (int)(rng() % 227);  // biased exponent 27..253  (2^-100 .. 2^126)
        const unsigned mant = (unsigned)(rng() & 0x7fffffu);
        const float T = __uint_as_float(((unsigned)e << 23) | mant);
        float m;
        switch (rng() % 8) {
            case 0: m = T; break;
            case 1: m = std::nextafter(T, 0.0f); break;
            case 2: m = std::nextafter(T, inf); break;
            case 3: m = 0.0f; break;
            case 4: m = inf; break;
            case 5: m = T * (0.5f + (float)(rng() % 1000) / 1000.0f); break;
            case 6: m = __uint_as_float((unsigned)(rng() & 0x7f7fffffu)); break;  // any finite non-negative float
            default: m = std::numeric_limits<float>::quiet_NaN(); break;        // inf - inf of a padding row
        }
        check(m, T);
    }
    for (float m : {0.0f, 1.0f, 3.4e38f, inf, std::numeric_limits<float>::quiet_NaN()}) check(m, inf);  // infinite radius
    // dead slots (T <= 0) never count
    for (float T : {0.0f, -1.0f})
        for (float m : {0.0f, 1.0f, inf}) {
            ++checked;
            if (ind_lt(m, make_ind_scale(T)) != 0.0f) { ++wrong; printf("WRONG dead slot T %g m %g\n", T, m); }
        }
    printf("checked %ld wrong %ld\n", checked, wrong);
    return wrong ? 1 : 0;
}

// K2 fast path: strict-radius neighbour counts in several subspaces from ONE pass over
// the candidates, fp32 filter + exact fp64 resolution of the boundary band.
//
// count_f32_kernel<Spec, Q>: same tiling as the k-NN filter (prepared fp32 AoS shadow,
// bulk-copied tiles, Q queries per thread in registers).  Per (query, candidate):
// DP/2 FADD2 form all coordinate differences once; the Spec turns their magnitudes into
// the S subspace maxima sharing partial maxima (3-input FMNMX3); each subspace costs one
// FSET (1.0f / 0.0f) against the query's threshold T and half a packed FADD2 into a
// float counter (exact: flushed to int32 every tile).  T = fp32_up(eps + delta) so the
// pass over-counts only inside the error band; no branch, no second compare.
//
// count_fix_kernel: for every query, the only candidates the filter can have
// misjudged have, in some coordinate, | |x_q - x_j| - eps | <= 4*delta.  They are found
// in the per-coordinate sorted arrays of the prepared set by binary search (a handful
// per query), re-evaluated with the filter's exact fp32 arithmetic AND in float64, and
// the counts corrected.  Queries whose band is too crowded (tie-saturated data) are
// flagged for the dense fp64 kernel.
//
// axis_count_kernel: a 1-D marginal needs no pass at all -- its count is the length of
// the run around x_q in that coordinate's sorted array that satisfies the exact
// predicate |fl64(x_q - v)| < eps, found by two binary searches.
#pragma once

#include "common.cuh"
#include "fast_knn.cuh"
#include "prepare.cuh"

namespace ksg {

__device__ __forceinline__ float max3f(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

// acc + a * b on both halves (exact here: every operand is a small integer)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// sat(x + c): 0 below -c, 1 above 1 - c
__device__ __forceinline__ float add_sat(float x, float c) {
    float r;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(c));
    return r;
}

// Mq of the query at `row` of the fp32 shadow (upper bound of max_d |x_d - centre_d|)
__device__ __forceinline__ double query_maxabs(const float* __restrict__ shadow, int64_t row, int dim_padded) {
    float m = 0.0f;
    for (int d = 0; d < dim_padded; ++d) m = fmaxf(m, fabsf(shadow[row * dim_padded + d]));
    return (double)m * (1.0 + 1e-6) + 1e-300;
}
// The filter threshold for a query: the smallest fp32 >= eps + delta(eps), one ulp up.
__device__ __forceinline__ float count_threshold(double eps, double mq) {
    if (!(eps < CUDART_INF)) return CUDART_INF_F;
    const float t = __double2float_ru(eps + band_query_f64(mq, eps));
    return nextafterf(t, CUDART_INF_F);
}

// ---- "m < T" as 1.0f / 0.0f on the FMA pipe.  FSET runs on the half-rate ALU pipe, which the
// subspace maxima (FMNMX / FMNMX3) already saturate (ncu, leave-one-out of 8: ALU 80 % busy,
// FMA 15 %).  One saturating FMA does the same: with B = 2^e chosen per query so that
// T*B ~ 2^100,   sat((-m)*B + T*B)   is 1 when m < T (the gap is >= one ulp of T, i.e.
// >= 2^-24 T, so the scaled difference is >= 2^76), 0 when m >= T (non-positive), 0 for the
// NaN of inf - inf (padding rows against an infinite radius) -- single rounding, power-of-two
// scaling, so the sign is exact.  Thresholds below 2^-100 cannot be scaled that far: those
// queries are flagged for the float64 kernel (count_fix_kernel).
constexpr float COUNT_MIN_T = 7.8886090522101181e-31f;  // 2^-100
struct IndScale {
    float neg_b, tb;  // -B and T*B
};
__device__ __forceinline__ IndScale make_ind_scale(float T) {
    IndScale s;
    if (!(T > 0.0f)) { s.neg_b = -1.0f; s.tb = -1.0f; return s; }  // dead slot: never counts
    const int e = (int)((__float_as_uint(T) >> 23) & 0xffu);        // T in [2^(e-127), 2^(e-126))
    int be = 354 - e;                                               // biased exponent of 2^(100 - (e - 127))
    be = be > 253 ? 253 : (be < 1 ? 1 : be);
    const float b = __uint_as_float((unsigned)be << 23);
    s.neg_b = -b;
    s.tb = T * b;  // exact (power of two), inf for an infinite radius
    return s;
}
__device__ __forceinline__ float ind_lt(float m, IndScale s) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(s.neg_b), "f"(s.tb));
    return r;
}

// ---------------------------------------------------------------- subspace specs
// a[d] = |q_d - c_d| for the DP (even-padded) joint dimensions; m[s] = subspace maxima.
template <int LO, int HI>
__device__ __forceinline__ float max_range(const float* a) {
    static_assert(HI > LO, "empty range");
    if constexpr (HI - LO == 1) return a[LO];
    else if constexpr (HI - LO == 2) return fmaxf(a[LO], a[LO + 1]);
    else if constexpr (HI - LO == 3) return max3f(a[LO], a[LO + 1], a[LO + 2]);
    else return max3f(a[LO], a[LO + 1], max_range<LO + 2, HI>(a));
}

// two disjoint groups: x = [0,DX), y = [DX,DX+DY)            (mutual information)
template <int DX, int DY>
struct SpecPair {
    static constexpr int D = DX + DY, DP = (D + 1) & ~1, S = 2;
    static constexpr bool INDICATOR_SUM = false;
    __device__ __forceinline__ static void eval(const float* a, float* m) {
        m[0] = max_range<0, DX>(a);
        m[1] = max_range<DX, DX + DY>(a);
    }
};

// three groups x, y, z; subspaces z, xz, yz                   (conditional MI)
template <int DX, int DY, int DZ>
struct SpecCmi {
    static constexpr int D = DX + DY + DZ, DP = (D + 1) & ~1, S = 3;
    static constexpr bool INDICATOR_SUM = false;
    __device__ __forceinline__ static void eval(const float* a, float* m) {
        const float mz = max_range<DX + DY, D>(a);
        m[0] = mz;
        if constexpr (DX == 1) m[1] = fmaxf(a[0], mz);
        else if constexpr (DX == 2) m[1] = max3f(a[0], a[1], mz);
        else m[1] = fmaxf(max_range<0, DX>(a), mz);
        if constexpr (DY == 1) m[2] = fmaxf(a[DX], mz);
        else if constexpr (DY == 2) m[2] = max3f(a[DX], a[DX + 1], mz);
        else m[2] = fmaxf(max_range<DX, DX + DY>(a), mz);
    }
};

// all D leave-one-out subspaces                               (DTC, O-, S-information)
//
// From 5 coordinates up the kernel does not form the D maxima at all (INDICATOR_SUM below).  With
// i_d = [a_d < T] per COORDINATE (the same saturating FMA, so the same predicate: max_{d != s} a_d < T
// <=> every i_d, d != s, is 1) and sigma = sum_d i_d,
//     count_s += [sigma == DP] + [sigma == DP - 1] * (1 - i_s)
// -- a candidate counts in every subspace when all coordinates are inside, and otherwise in exactly the one
// subspace that drops its only outside coordinate.  DP FFMA.SAT, a packed add tree, two saturating adds,
// DP/2 packed FMAs: 36 FP32 lane-operations per pair for D = 8, all on the FMA pipe.  Measured on the
// inner loop in isolation (profiles/pipe_microbench2.cu, r2s_pipe_microbench2.txt; a 250 000-sample pass):
// maxima form (24 FMA-pipe lane-ops + 14 FMNMX / FMNMX3) 81.0 ms, this form 70.8 ms, pair maxima on the ALU
// pipe + 12 indicators 80.8 ms, integer sign-bit forms 85-89 ms: on sm_100a the 16-lane ALU pipe does not
// overlap with the FMA pipe in such a loop (the times follow the SUM of both pipes' cycles, a packed
// FADD2 / FFMA2 occupying the FMA pipe for two), so the cheapest form is the one with the fewest lane-ops.
template <int DD>
struct SpecLoo {
    static constexpr int D = DD, DP = (D + 1) & ~1, S = DD;
#ifdef KSG_LOO_MAXIMA
    static constexpr bool INDICATOR_SUM = false;
#else
    static constexpr bool INDICATOR_SUM = DD >= 5;
#endif
    __device__ __forceinline__ static void eval(const float* a, float* m) {
        if constexpr (D == 3) {
            m[0] = fmaxf(a[1], a[2]); m[1] = fmaxf(a[0], a[2]); m[2] = fmaxf(a[0], a[1]);
        } else if constexpr (D == 4) {
            const float p01 = fmaxf(a[0], a[1]), p23 = fmaxf(a[2], a[3]);
            m[0] = fmaxf(a[1], p23); m[1] = fmaxf(a[0], p23); m[2] = fmaxf(a[3], p01); m[3] = fmaxf(a[2], p01);
        } else if constexpr (D == 5) {
            const float p01 = fmaxf(a[0], a[1]), p23 = fmaxf(a[2], a[3]);
            m[0] = max3f(a[1], p23, a[4]); m[1] = max3f(a[0], p23, a[4]);
            m[2] = max3f(a[3], p01, a[4]); m[3] = max3f(a[2], p01, a[4]); m[4] = fmaxf(p01, p23);
        } else if constexpr (D == 6) {
            const float p01 = fmaxf(a[0], a[1]), p23 = fmaxf(a[2], a[3]), p45 = fmaxf(a[4], a[5]);
            m[0] = max3f(a[1], p23, p45); m[1] = max3f(a[0], p23, p45);
            m[2] = max3f(a[3], p01, p45); m[3] = max3f(a[2], p01, p45);
            m[4] = max3f(a[5], p01, p23); m[5] = max3f(a[4], p01, p23);
        } else if constexpr (D == 7) {
            const float p01 = fmaxf(a[0], a[1]), p23 = fmaxf(a[2], a[3]), p45 = fmaxf(a[4], a[5]);
            const float r = fmaxf(p45, a[6]), s = fmaxf(p01, p23);
            m[0] = max3f(a[1], p23, r); m[1] = max3f(a[0], p23, r);
            m[2] = max3f(a[3], p01, r); m[3] = max3f(a[2], p01, r);
            m[4] = max3f(a[5], a[6], s); m[5] = max3f(a[4], a[6], s); m[6] = fmaxf(p45, s);
        } else {
            static_assert(D == 8, "SpecLoo supports 3..8 dimensions");
            const float p01 = fmaxf(a[0], a[1]), p23 = fmaxf(a[2], a[3]);
            const float p45 = fmaxf(a[4], a[5]), p67 = fmaxf(a[6], a[7]);
            const float lo = fmaxf(p01, p23), hi = fmaxf(p45, p67);
            m[0] = max3f(a[1], p23, hi); m[1] = max3f(a[0], p23, hi);
            m[2] = max3f(a[3], p01, hi); m[3] = max3f(a[2], p01, hi);
            m[4] = max3f(a[5], p67, lo); m[5] = max3f(a[4], p67, lo);
            m[6] = max3f(a[7], p45, lo); m[7] = max3f(a[6], p45, lo);
        }
    }
};

// up to 4 arbitrary subspaces given as run-time bitmasks (uniform) -- generic fallback
template <int DPP>
struct SpecMasked {
    static constexpr int D = DPP, DP = DPP, S = 4;
    static constexpr bool INDICATOR_SUM = false;
};

struct MaskQuad {
    uint32_t m[4];
};

// ---------------------------------------------------------------- filter kernel
// counts32[s][nq] (int32) are ADDED to (atomicAdd when the candidate range is split).
template <class Spec, int Q, int STAGES, bool MASKED>
static __global__ void __launch_bounds__(FAST_THREADS)
count_f32_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box, int64_t n_tiles,
                 int64_t q_begin, int64_t nq, const float* __restrict__ thresholds, int windowed, MaskQuad masks,
                 int n_sub, int32_t* __restrict__ counts, unsigned long long* __restrict__ tile_counter) {
    constexpr int DP = Spec::DP, S = Spec::S, SP = (S + 1) / 2;
    constexpr int QB = FAST_THREADS * Q;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ TileQueueShared sh;

    const int tid = threadIdx.x;
    const int64_t qb0 = (int64_t)blockIdx.x * QB;
    const int splits = gridDim.y;

    float2 q[Q][DP / 2];
    float T[Q];
    IndScale sc[Q];
    int cnt[Q][S];
    float blo[DP], bhi[DP];
    float tmax = -CUDART_INF_F;
#pragma unroll
    for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        const bool live = qi < nq;
        const int64_t row = q_begin + (live ? qi : 0);
        load_row<DP>(shadow + row * DP, q[j]);
        T[j] = live ? thresholds[qi] : -1.0f;
        sc[j] = make_ind_scale(T[j]);
#pragma unroll
        for (int s = 0; s < S; ++s) cnt[j][s] = 0;
        if (live) {
            tmax = fmaxf(tmax, T[j]);
#pragma unroll
            for (int h = 0; h < DP / 2; ++h) {
                blo[2 * h] = fminf(blo[2 * h], q[j][h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[j][h].x);
                blo[2 * h + 1] = fminf(blo[2 * h + 1], q[j][h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[j][h].y);
            }
        }
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    block_query_box<DP>(blo, bhi, sh);
    const float Tmax = block_max(tmax, sh);  // a tile is needed iff some subspace's box gap is < Tmax

    auto subspace_maxima = [&](const float* a, float* m) {
        if constexpr (MASKED) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float v = 0.0f;
#pragma unroll
                for (int d = 0; d < DP; ++d)
                    if ((masks.m[s] >> d) & 1u) v = fmaxf(v, a[d]);
                m[s] = masks.m[s] ? v : CUDART_INF_F;  // unused slot: never inside
            }
        } else {
            Spec::eval(a, m);
        }
    };
    auto test = [&](int64_t t) -> bool {
        float gap[DP], m[S];
        box_gaps<DP>(tile_box, t, sh, gap);
        subspace_maxima(gap, m);
        bool need = false;
#pragma unroll
        for (int s = 0; s < S; ++s) need = need || (s < n_sub && m[s] < Tmax);
        return need;
    };
    auto refresh = [&]() {};
    auto compute = [&](const float* tile, int64_t) {
        if constexpr (Spec::INDICATOR_SUM) {
            // acc[j][h] = sum over the tile of e * (i_2h, i_2h+1);  ae[j] = sum of ([sigma == DP], e),
            // e = [sigma == DP - 1]  (see SpecLoo).  Small integers: every operation is exact.
            constexpr int H = DP / 2;
            float2 acc[Q][H], ae[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                ae[j] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int h = 0; h < H; ++h) acc[j][h] = make_float2(0.0f, 0.0f);
            }
#pragma unroll 2
            for (int c = 0; c < FAST_TILE; ++c) {
                float2 cc[H];
                load_row<DP>(tile + c * DP, cc);
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    float2 ind[H];
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        const float2 d = sub2(q[j][h], cc[h]);
                        ind[h] = make_float2(ind_lt(fabsf(d.x), sc[j]), ind_lt(fabsf(d.y), sc[j]));
                    }
                    float2 s2 = ind[0];
#pragma unroll
                    for (int h = 1; h < H; ++h) s2 = add2(s2, ind[h]);
                    const float sigma = s2.x + s2.y;
                    const float all = add_sat(sigma, -(float)(DP - 1));      // [sigma == DP]
                    const float e = add_sat(sigma, -(float)(DP - 2)) - all;  // [sigma == DP - 1]
                    ae[j] = add2(ae[j], make_float2(all, e));
                    const float2 ee = make_float2(e, e);
#pragma unroll
                    for (int h = 0; h < H; ++h) acc[j][h] = fma2(ind[h], ee, acc[j][h]);
                }
            }
#pragma unroll
            for (int j = 0; j < Q; ++j)
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const float in_s = (s & 1) ? acc[j][s / 2].y : acc[j][s / 2].x;
                    cnt[j][s] += (int)(ae[j].x + (ae[j].y - in_s));
                }
            return;
        } else {
        float2 acc[Q][SP];
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
            for (int h = 0; h < SP; ++h) acc[j][h] = make_float2(0.0f, 0.0f);
#pragma unroll 2
        for (int c = 0; c < FAST_TILE; ++c) {
            float2 cc[DP / 2];
            load_row<DP>(tile + c * DP, cc);
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                float a[DP];
#pragma unroll
                for (int h = 0; h < DP / 2; ++h) {
                    const float2 d = sub2(q[j][h], cc[h]);
                    a[2 * h] = fabsf(d.x);
                    a[2 * h + 1] = fabsf(d.y);
                }
                float m[S];
                subspace_maxima(a, m);
                float ind[2 * SP];
#pragma unroll
                for (int s = 0; s < S; ++s) ind[s] = ind_lt(m[s], sc[j]);
                if constexpr (S & 1) ind[S] = 0.0f;
#pragma unroll
                for (int h = 0; h < SP; ++h) acc[j][h] = add2(acc[j][h], make_float2(ind[2 * h], ind[2 * h + 1]));
            }
        }
        // flush the float counters (<= FAST_TILE, exact) into integers
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
            for (int s = 0; s < S; ++s) cnt[j][s] += (int)((s & 1) ? acc[j][s / 2].y : acc[j][s / 2].x);
        }
    };

    const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
    const int tiles_done = run_tile_queue<DP, STAGES>(shadow, tiles, full_bar, sh, mid / FAST_TILE, n_tiles,
                                                      (int)blockIdx.y, splits, windowed != 0, test, compute, refresh);
    if (tid == 0 && tile_counter) atomicAdd(tile_counter, (unsigned long long)tiles_done * (FAST_TILE * QB));
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if (qi < nq) {
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (s < n_sub) {
                    if (splits == 1) counts[(int64_t)s * nq + qi] += cnt[j][s];
                    else if (cnt[j][s]) atomicAdd(&counts[(int64_t)s * nq + qi], cnt[j][s]);
                }
        }
    }
}

// thresholds[qi] = count_threshold(eps[qi]);  counts initialised to `bias`
// exact32 (prepared flags[8]): the fp32 distances are the float64 distances themselves, so the filter's
// "m < T" is made the exact predicate -- strict: d < eps  <=>  d < ceil32(eps); closed: d <= eps  <=>
// d < nextup(floor32(eps)) -- and count_fix_kernel has nothing to correct.
static __global__ void count_thresholds_kernel(const double* __restrict__ eps, int64_t q_begin, int64_t nq,
                                        const float* __restrict__ shadow, int dim_padded, int strict,
                                        const int32_t* __restrict__ exact32, float* __restrict__ thresholds) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const double e = eps[qi];
    if (exact32 && *exact32 && e < CUDART_INF && e >= 0.0) {
        thresholds[qi] = strict ? __double2float_ru(e) : nextafterf(__double2float_rd(e), CUDART_INF_F);
        return;
    }
    thresholds[qi] = count_threshold(e, query_maxabs(shadow, q_begin + qi, dim_padded));
}

// ---------------------------------------------------------------- exact resolution of the band
constexpr int FIX_CAP = 512;  // band candidates examined per query, coordinate and side

__device__ __forceinline__ int64_t lower_bound_f64(const double* __restrict__ v, int64_t n, double x) {
    int64_t lo = 0, hi = n;  // first index with v[i] >= x
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int64_t upper_bound_f64(const double* __restrict__ v, int64_t n, double x) {
    int64_t lo = 0, hi = n;  // first index with v[i] > x
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (v[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct MaskTable64 {
    uint32_t m[KSG_MAX_SUBSPACES];
};

static __global__ void __launch_bounds__(128)
count_fix_kernel(const double* __restrict__ sorted, int64_t ld, const float* __restrict__ shadow, int dim,
                 int dim_padded, int64_t n, const double* __restrict__ axis_vals,
                 const int32_t* __restrict__ axis_pos, const double* __restrict__ maxabs, int64_t q_begin,
                 int64_t nq, const double* __restrict__ eps, const float* __restrict__ thresholds,
                 MaskTable64 masks, int n_sub, int strict, int32_t* __restrict__ counts,
                 int32_t* __restrict__ flags, int32_t* __restrict__ n_flagged, const int32_t* __restrict__ exact32) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const double e = eps[qi];
    if (!(e < CUDART_INF)) return;  // infinite radius: the filter is exact (everything finite is inside)
    const float T = thresholds[qi];
    if (exact32 && *exact32 && e >= 0.0 && fp32_range_ok(maxabs, dim)) {
        // exact fp32 copy: the filter already counted with the exact predicate (count_thresholds_kernel);
        // only thresholds the scaled compare cannot represent go to the float64 kernel
        if (T > 0.0f && T < COUNT_MIN_T) {
            flags[qi] = 1;
            atomicAdd(n_flagged, 1);
        }
        return;
    }
    // (also: shadow rows beyond the fp32 range look like padding to the filter, see fp32_range_ok)
    if (T < COUNT_MIN_T || !fp32_range_ok(maxabs, dim)) {  // the filter's scaled compare is not exact this far down (see ind_lt)
        flags[qi] = 1;
        atomicAdd(n_flagged, 1);
        return;
    }
    uint32_t used = 0;
    for (int s = 0; s < n_sub; ++s) used |= masks.m[s];
    const int64_t row_q = q_begin + qi;
    const double w = 4.0 * band_query_f64(query_maxabs(shadow, row_q, dim_padded), e);

    double xq[KSG_MAX_DIM];
    for (int d = 0; d < dim; ++d) xq[d] = sorted[(int64_t)d * ld + row_q];
    bool crowded = false;

    for (int d = 0; d < dim && !crowded; ++d) {
        if (!((used >> d) & 1u)) continue;
        const double* vals = axis_vals + (int64_t)d * ld;
        const int32_t* pos = axis_pos + (int64_t)d * ld;
        // superset of {v : | |xq - v| - e | <= w}: two intervals (or one when they touch)
        const double ws = w * 1.01 + 1e-300;
        double lo_a = xq[d] - e - ws, hi_a = xq[d] - e + ws;
        double lo_b = xq[d] + e - ws, hi_b = xq[d] + e + ws;
        int n_int = 2;
        if (hi_a >= lo_b) { hi_a = hi_b; n_int = 1; }
        for (int side = 0; side < n_int && !crowded; ++side) {
            const double lo_v = side == 0 ? lo_a : lo_b, hi_v = side == 0 ? hi_a : hi_b;
            const int64_t i0 = lower_bound_f64(vals, n, lo_v), i1 = upper_bound_f64(vals, n, hi_v);
            if (i1 - i0 > FIX_CAP) { crowded = true; break; }
            for (int64_t i = i0; i < i1; ++i) {
                const int64_t j = pos[i];
                // in-band test in float64, the same for every coordinate (also dedupes)
                double a64[KSG_MAX_DIM];
                bool first_band = true, in_band_d = false;
                for (int dd = 0; dd < dim; ++dd) {
                    a64[dd] = fabs(xq[dd] - sorted[(int64_t)dd * ld + j]);
                    const bool inb = ((used >> dd) & 1u) && fabs(a64[dd] - e) <= w;
                    if (dd < d && inb) first_band = false;   // handled when coordinate dd was visited
                    if (dd == d) in_band_d = inb;
                }
                if (!in_band_d || !first_band) continue;
                for (int s = 0; s < n_sub; ++s) {
                    float m32 = 0.0f;
                    double m64 = 0.0;
                    for (int dd = 0; dd < dim; ++dd)
                        if ((masks.m[s] >> dd) & 1u) {
                            const float df = __fsub_rn(shadow[row_q * dim_padded + dd], shadow[j * dim_padded + dd]);
                            m32 = fmaxf(m32, fabsf(df));
                            m64 = fmax(m64, a64[dd]);
                        }
                    const int counted = m32 < T;
                    const int truth = strict ? (m64 < e) : (m64 <= e);
                    if (counted != truth) atomicAdd(&counts[(int64_t)s * nq + qi], truth - counted);
                }
            }
        }
    }
    if (crowded) {
        flags[qi] = 1;
        atomicAdd(n_flagged, 1);
    }
}

// ---------------------------------------------------------------- 1-D marginals
// counts[qi] = #{j : |fl64(x_q - v_j)| < eps} + bias   over the coordinate's sorted array: the
// length of the run around x_q that satisfies the exact predicate (monotone along the sorted
// array because IEEE rounding is monotone), found by two binary searches.  The searches are
// latency-bound chains of dependent loads, so their upper levels run in shared memory: every CTA
// first samples the array at AXIS_TAB evenly spaced positions (the last element of each chunk),
// a search picks its chunk there and only bisects inside the chunk in global memory.
constexpr int AXIS_THREADS = 1024;
constexpr int AXIS_TAB = 1024;

static __global__ void __launch_bounds__(AXIS_THREADS)
axis_count_kernel(const double* __restrict__ sorted_row, const double* __restrict__ vals, int64_t n,
                  int64_t q_begin, int64_t nq, const double* __restrict__ eps, int strict, int32_t bias,
                  int32_t* __restrict__ counts) {
    __shared__ double tab[AXIS_TAB];
    const int64_t chunk = (n + AXIS_TAB - 1) / AXIS_TAB;          // elements per chunk (>= 1)
    const int n_chunks = (int)((n + chunk - 1) / chunk);          // <= AXIS_TAB
    for (int i = threadIdx.x; i < n_chunks; i += AXIS_THREADS) {
        const int64_t last = ((int64_t)i + 1) * chunk - 1;
        tab[i] = vals[last < n ? last : n - 1];
    }
    __syncthreads();
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const double x = sorted_row[q_begin + qi];
    const double e = eps[qi];
    // first index whose value satisfies pred (pred is false ... false true ... true along the array)
    auto first_true = [&](auto pred) -> int64_t {
        int lo = 0, hi = n_chunks;  // first chunk whose LAST element satisfies pred
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (pred(tab[mid])) hi = mid; else lo = mid + 1;
        }
        if (lo == n_chunks) return n;
        int64_t a = (int64_t)lo * chunk, b = a + chunk < n ? a + chunk : n;
        b -= 1;  // the chunk's last element satisfies pred: the answer is in [a, b]
        while (a < b) {
            const int64_t mid = (a + b) >> 1;
            if (pred(vals[mid])) b = mid; else a = mid + 1;
        }
        return a;
    };
    // first index inside on the left: (v > x) || inside(x - v)
    const int64_t left = first_true([&](double v) { return (v > x) || (strict ? (x - v < e) : (x - v <= e)); });
    // first index beyond the run on the right: (v > x) && !inside(v - x)
    const int64_t right = first_true([&](double v) { return (v > x) && !(strict ? (v - x < e) : (v - x <= e)); });
    counts[qi] = (int32_t)(right - left) + bias;
}

// The same counts with the threads in the order of the COORDINATE: thread i owns the i-th smallest
// value x = vals[i] (pos[i] = its row in the prepared order, where its radius lives and its count
// goes).  Both ends of its run lie near i, so they are found by galloping outwards from i (1, 2, 4, ...
// steps, then a bisection of the last step): ~2 log2(count) loads instead of 2 log2(n), and the runs of
// neighbouring threads nearly coincide, so a warp touches one or two cache lines per step instead of 32.
// Rows outside the query slice [q_begin, q_begin + nq) are skipped (several ranks).
__device__ __forceinline__ void axis_count_sorted_one(const double* __restrict__ vals, const int32_t* __restrict__ pos,
                                                      int64_t n, int64_t q_begin, int64_t nq,
                                                      const double* __restrict__ eps, int strict, int32_t bias,
                                                      int32_t* __restrict__ counts, int64_t i) {
    if (i >= n) return;
    const int64_t qi = (int64_t)pos[i] - q_begin;
    if (qi < 0 || qi >= nq) return;
    const double x = vals[i];
    const double e = eps[qi];
    // first index at which a monotone predicate (false ... false true ... true) holds; pred(n) = true
    auto first_true_from = [&](auto pred) -> int64_t {
        if (pred(x)) {
            // true at i: gallop to the left for the last false
            int64_t hi = i, lo = i - 1, step = 1;
            while (lo >= 0 && pred(vals[lo])) { hi = lo; lo -= step; step <<= 1; }
            if (lo < -1) lo = -1;
            while (hi - lo > 1) {
                const int64_t mid = lo + ((hi - lo) >> 1);
                if (pred(vals[mid])) hi = mid; else lo = mid;
            }
            return hi;
        }
        // false at i: gallop to the right for the first true
        int64_t lo = i, hi = i + 1, step = 1;
        while (hi < n && !pred(vals[hi])) { lo = hi; hi += step; step <<= 1; }
        if (hi > n) hi = n;
        while (hi - lo > 1) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (pred(vals[mid])) hi = mid; else lo = mid;
        }
        return hi;
    };
    const int64_t left = first_true_from([&](double v) { return (v > x) || (strict ? (x - v < e) : (x - v <= e)); });
    const int64_t right = first_true_from([&](double v) { return (v > x) && !(strict ? (v - x < e) : (v - x <= e)); });
    counts[qi] = (int32_t)(right - left) + bias;
}

static __global__ void __launch_bounds__(256)
axis_count_sorted_kernel(const double* __restrict__ vals, const int32_t* __restrict__ pos, int64_t n,
                         int64_t q_begin, int64_t nq, const double* __restrict__ eps, int strict, int32_t bias,
                         int32_t* __restrict__ counts) {
    axis_count_sorted_one(vals, pos, n, q_begin, nq, eps, strict, bias, counts,
                          (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

static __global__ void scatter_f64_kernel(const double* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                                   double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm[i]] = src[i];
}

// out[0] += sum of counts (int64), one atomic per block
static __global__ void __launch_bounds__(256)
sum_i32_kernel(const int32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += v[i];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __shared__ long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(out, (unsigned long long)t);
    }
}

static __global__ void scatter_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                                   int32_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm[i]] = src[i];
}
static __global__ void gather_f64_kernel(const double* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                                  double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

// Several 1-D marginals in ONE launch (blockIdx.y = which): the searches are latency-bound, two
// coordinates side by side cost little more than one.  counts: [n_axes][nq].
struct AxisList {
    int32_t axis[KSG_MAX_DIM];
};
static __global__ void __launch_bounds__(256)
axis_count_sorted_multi_kernel(const double* __restrict__ axis_vals, const int32_t* __restrict__ axis_pos, int64_t ld,
                               const __grid_constant__ AxisList axes, int64_t n, int64_t q_begin, int64_t nq,
                               const double* __restrict__ eps, int strict, int32_t bias, int32_t* __restrict__ counts) {
    const int64_t a = axes.axis[blockIdx.y];
    axis_count_sorted_one(axis_vals + a * ld, axis_pos + a * ld, n, q_begin, nq, eps, strict, bias,
                          counts + (int64_t)blockIdx.y * nq, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

}  // namespace ksg

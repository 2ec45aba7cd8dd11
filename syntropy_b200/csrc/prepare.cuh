// Prepared point set: everything the fast (fp32-filtered, exact) kernels need, built
// once per estimator call on the device.
//
//   * the points in a space-filling order -- the Hilbert curve over the per-coordinate ranks
//     (primary = -1), the leaf order of a balanced kd-tree (primary = -2) or one coordinate's
//     sorted order (primary >= 0): neighbouring queries then see neighbouring candidates, which
//     keeps the k-NN thresholds tight from the first tile on and gives 512-row tiles and 32-row
//     sub-tiles tight bounding boxes (the windowed engine's pruning);
//   * an fp32 AoS shadow copy, centred per dimension so that its rounding error is
//     2^-24 * |x - centre| instead of 2^-24 * |x|, padded with +inf rows to a whole
//     number of tiles (padded candidates can never be a neighbour or be counted);
//   * every coordinate sorted on its own (values + position in the primary order):
//     1-D marginal counts become two binary searches with the exact fp64 predicate,
//     and the fp32 count kernel's boundary band is resolved by looking at the handful
//     of candidates whose coordinate difference is within the band of the radius.
#pragma once

#include <stdlib.h>

#include "common.cuh"
#include "sample_sort.cuh"

namespace ksg {

constexpr int FAST_TILE = 512;       // candidates per shared-memory tile (fast kernels)
constexpr int FAST_SUB = 32;         // rows of a sub-tile (own bounding box: warp-level culling)
constexpr int FAST_NSUB = FAST_TILE / FAST_SUB;
constexpr int STATS_BLOCKS = 256;    // partial min/max blocks per dimension
constexpr int SORT_LANES = 4;        // per-coordinate sorts of ksg_prepare that run concurrently

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// (layout computation lives in fast_abi.cu: it needs CUB's temp-storage size)
struct PreparedLayout {
    int64_t ld;        // row stride of the fp64 arrays
    int64_t n_padded;  // shadow rows
    int dim_padded;
    size_t off_perm, off_sorted, off_shadow, off_axis_vals, off_axis_pos, off_center, off_maxabs, off_flags;
    size_t off_tile_box, off_sub_box;
    size_t off_iota, off_ranks, off_mkeys, off_mkeys_out, off_invperm, off_cub, off_curve_table, off_stats, off_exact;
    size_t cub_bytes;  // per sort lane
    int sort_lanes;
    size_t total;
};

// ---- non-finite flag (and per-dimension partial min / max, used by the cross-set queries' range)
static __global__ void __launch_bounds__(256)
stats_stage1_kernel(const double* __restrict__ pts, int64_t ld, int64_t n, double* __restrict__ partial,
                    int32_t* __restrict__ flags) {
    const int d = blockIdx.y;
    const double* row = pts + (int64_t)d * ld;
    double lo = CUDART_INF, hi = -CUDART_INF;
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = row[i];
        bad |= !isfinite(v);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    __shared__ double slo[256], shi[256];
    slo[threadIdx.x] = lo;
    shi[threadIdx.x] = hi;
    const int anybad = __syncthreads_or(bad);
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + off]);
            shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[((int64_t)d * STATS_BLOCKS + blockIdx.x) * 2 + 0] = slo[0];
        partial[((int64_t)d * STATS_BLOCKS + blockIdx.x) * 2 + 1] = shi[0];
        if (anybad) atomicExch(&flags[0], 1);
    }
}

// centre[d] = median of coordinate d (middle entry of its sorted array); maxabs[d] = max |x - centre|;
// flags[0] = 1 if any coordinate holds a NaN or an infinity (the radix order puts -inf and negative
// NaNs first, +inf and positive NaNs last: the two ends of every sorted array tell)
static __global__ void center_median_kernel(const double* __restrict__ axis_vals, int64_t ld, int64_t n, int dim,
                                            double* __restrict__ center, double* __restrict__ maxabs,
                                            int32_t* __restrict__ flags) {
    const int d = threadIdx.x;
    if (d >= dim) return;
    const double* v = axis_vals + (int64_t)d * ld;
    double c = v[n / 2], lo = v[0], hi = v[n - 1];
    if (!isfinite(lo) || !isfinite(hi)) atomicExch(&flags[0], 1);
    if (!isfinite(c)) c = 0.0;  // non-finite input is flagged elsewhere; keep the math finite
    if (!isfinite(lo)) lo = c;
    if (!isfinite(hi)) hi = c;
    center[d] = c;
    maxabs[d] = fmax(fabs(hi - c), fabs(c - lo)) * (1.0 + 1e-12) + 1e-300;
}

static __global__ void iota_kernel(int32_t* __restrict__ p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

// ---- "the fp32 copy is EXACT" (quantised / integer data).  If every centred coordinate x - centre is
// a multiple of one power of two q_d and smaller than 2^23 q_d, then the subtraction x - centre, its
// conversion to fp32 and every difference of two shadow values are exact: the fp32 filter computes
// the float64 distances themselves, the k'-th smallest fp32 distance IS the exact radius (ties need
// no identities), the fp32 counts are the exact counts, and nothing needs a recheck or a fallback.
// exact_info (per block of 256 rows, no atomics): [b][0] = an inexact step was seen, [b][1 + 2 d] =
// min exponent of the lowest set bit of coordinate d's shadow values, [b][2 + 2 d] = max exponent of
// their highest bit; exact_flag_kernel folds the blocks.
constexpr int EXACT_INFO_INTS = 2 * 8 + 2;  // the fp32 filters serve up to 8 coordinates
constexpr int EXACT_NONE_LOW = 0x7fffffff, EXACT_NONE_HIGH = -0x7fffffff;

__device__ __forceinline__ void shadow_bit_range(float s, int& low, int& high) {
    const unsigned int b = __float_as_uint(s) & 0x7fffffffu;
    const int e = (int)(b >> 23);
    const unsigned int mant = (b & 0x7fffffu) | (e ? 0x800000u : 0u);
    const int base = (e ? e : 1) - 150;  // value = mant * 2^base
    low = base + (__ffs((int)mant) - 1);
    high = base + (31 - __clz((int)mant));
}

// sorted_f64[d][i] = pts[d][perm[i]];  shadow[i][d] = fp32(sorted - centre);  +inf padding rows
static __global__ void __launch_bounds__(256)
gather_sorted_kernel(const double* __restrict__ pts, int64_t ld_in, int64_t n, int dim, int dim_padded,
                     const int32_t* __restrict__ perm, const double* __restrict__ center,
                     double* __restrict__ sorted, int64_t ld, float* __restrict__ shadow, int64_t n_padded,
                     int32_t* __restrict__ exact_info) {
    __shared__ int s_low[8][8], s_high[8][8], s_bad[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_range = i < n_padded;
    const bool real = i < n;
    const int64_t src = real ? perm[i] : 0;
    const bool analyse = exact_info != nullptr && dim <= 8;
    int inexact = 0;
    for (int d = 0; d < dim_padded; ++d) {
        float s;
        int low = EXACT_NONE_LOW, high = EXACT_NONE_HIGH;
        if (d < dim) {
            if (real) {
                const double v = pts[(int64_t)d * ld_in + src], c = center[d];
                sorted[(int64_t)d * ld + i] = v;
                const double t = __dsub_rn(v, c);
                s = __double2float_rn(t);
                if (analyse) {
                    // error-free transformation of v - c (TwoSum) and of the conversion
                    const double bb = __dsub_rn(t, v);
                    const double err = __dadd_rn(__dsub_rn(v, __dsub_rn(t, bb)), __dsub_rn(-c, bb));
                    inexact |= !(err == 0.0) || !((double)s == t);
                    if (s != 0.0f && isfinite(s)) shadow_bit_range(s, low, high);
                }
            } else {
                s = CUDART_INF_F;
            }
        } else {
            s = 0.0f;
        }
        if (in_range) shadow[i * dim_padded + d] = s;
        if (analyse && d < dim) {
            low = __reduce_min_sync(0xffffffffu, low);
            high = __reduce_max_sync(0xffffffffu, high);
            if (lane == 0) { s_low[warp][d] = low; s_high[warp][d] = high; }
        }
    }
    if (!analyse) return;
    inexact = __any_sync(0xffffffffu, inexact);
    if (lane == 0) s_bad[warp] = inexact;
    __syncthreads();
    int32_t* out = exact_info + (int64_t)blockIdx.x * EXACT_INFO_INTS;
    if (threadIdx.x < dim) {
        const int d = threadIdx.x;
        int low = EXACT_NONE_LOW, high = EXACT_NONE_HIGH;
        for (int w = 0; w < 8; ++w) { low = min(low, s_low[w][d]); high = max(high, s_high[w][d]); }
        out[1 + 2 * d] = low;
        out[2 + 2 * d] = high;
    }
    if (threadIdx.x == 0) {
        int bad = 0;
        for (int w = 0; w < 8; ++w) bad |= s_bad[w];
        out[0] = bad;
    }
}

// flags[8] = 1 when the fp32 copy is exact (see above); one CTA of 256 threads folds the block records
static __global__ void __launch_bounds__(256)
exact_flag_kernel(const int32_t* __restrict__ exact_info, int64_t n_blocks, int dim, int force_off,
                  int32_t* __restrict__ flags) {
    __shared__ int s_low[256], s_high[256], s_ok;
    if (threadIdx.x == 0) s_ok = (dim <= 8 && !force_off) ? 1 : 0;
    __syncthreads();
    if (dim > 8 || force_off) { if (threadIdx.x == 0) flags[8] = 0; return; }
    int bad = 0;
    for (int64_t b = threadIdx.x; b < n_blocks; b += 256) bad |= exact_info[b * EXACT_INFO_INTS];
    if (bad) s_ok = 0;
    for (int d = 0; d < dim; ++d) {
        int low = EXACT_NONE_LOW, high = EXACT_NONE_HIGH;
        for (int64_t b = threadIdx.x; b < n_blocks; b += 256) {
            low = min(low, exact_info[b * EXACT_INFO_INTS + 1 + 2 * d]);
            high = max(high, exact_info[b * EXACT_INFO_INTS + 2 + 2 * d]);
        }
        s_low[threadIdx.x] = low;
        s_high[threadIdx.x] = high;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) {
                s_low[threadIdx.x] = min(s_low[threadIdx.x], s_low[threadIdx.x + off]);
                s_high[threadIdx.x] = max(s_high[threadIdx.x], s_high[threadIdx.x + off]);
            }
            __syncthreads();
        }
        // (a coordinate that is all zeros has no bits at all: fine)
        if (threadIdx.x == 0 && s_low[0] != EXACT_NONE_LOW && s_high[0] - s_low[0] > 22) s_ok = 0;
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) flags[8] = s_ok;
}

// gather_sorted_kernel + invert_perm_kernel + tile_box_kernel + exact_flag_kernel as ONE launch, one CTA
// per 512-row tile (two passes of 256 rows; warp w of pass p owns sub-tile 8 p + w, so the sub-tile
// boxes are warp reductions of values still in registers).  The CTA that finishes last (ticket =
// flags[9], zeroed with the flag block at the start of ksg_prepare) folds the per-tile exactness
// records into flags[8].  Same outputs, bit for bit, as the four kernels.
static __global__ void __launch_bounds__(256)
gather_tile_kernel(const double* __restrict__ pts, int64_t ld_in, int64_t n, int dim, int dim_padded,
                   const int32_t* __restrict__ perm, const double* __restrict__ center,
                   double* __restrict__ sorted, int64_t ld, float* __restrict__ shadow,
                   int32_t* __restrict__ invperm, float* __restrict__ tile_box, float* __restrict__ sub_box,
                   int32_t* __restrict__ exact_info, int force_off, int32_t* __restrict__ flags) {
    static_assert(FAST_TILE == 512 && FAST_SUB == 32, "two passes of 256 rows, one sub-tile per warp and pass");
    __shared__ float slo[8][KSG_MAX_DIM], shi[8][KSG_MAX_DIM];
    __shared__ int s_low[8][8], s_high[8][8], s_bad[8], s_last;
    const int64_t t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool analyse = dim <= 8;
    int inexact = 0;
    int alow[8], ahigh[8];  // (dim <= 8 only) per coordinate, this thread's two rows
#pragma unroll
    for (int d = 0; d < 8; ++d) { alow[d] = EXACT_NONE_LOW; ahigh[d] = EXACT_NONE_HIGH; }
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
        const int64_t i = t * FAST_TILE + p * 256 + threadIdx.x;
        const bool real = i < n;
        const int64_t src = real ? perm[i] : 0;
        if (real && invperm) invperm[src] = (int32_t)i;
        const int sub = p * 8 + warp;
        for (int d = 0; d < dim_padded; ++d) {
            float s;
            if (d < dim) {
                if (real) {
                    const double v = pts[(int64_t)d * ld_in + src], c = center[d];
                    sorted[(int64_t)d * ld + i] = v;
                    const double tt = __dsub_rn(v, c);
                    s = __double2float_rn(tt);
                    if (analyse) {
                        // error-free transformation of v - c (TwoSum) and of the conversion
                        const double bb = __dsub_rn(tt, v);
                        const double err = __dadd_rn(__dsub_rn(v, __dsub_rn(tt, bb)), __dsub_rn(-c, bb));
                        inexact |= !(err == 0.0) || !((double)s == tt);
                        if (s != 0.0f && isfinite(s)) {
                            int low, high;
                            shadow_bit_range(s, low, high);
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (e == d) { alow[e] = min(alow[e], low); ahigh[e] = max(ahigh[e], high); }
                        }
                    }
                } else {
                    s = CUDART_INF_F;
                }
            } else {
                s = 0.0f;
            }
            shadow[i * dim_padded + d] = s;  // (n_padded is a multiple of the tile: every row exists)
            float lo = s, hi = s;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
            }
            if (lane == 0) {
                sub_box[((t * FAST_NSUB + sub) * dim_padded + d) * 2 + 0] = lo;
                sub_box[((t * FAST_NSUB + sub) * dim_padded + d) * 2 + 1] = hi;
                slo[warp][d] = p == 0 ? lo : fminf(slo[warp][d], lo);
                shi[warp][d] = p == 0 ? hi : fmaxf(shi[warp][d], hi);
            }
        }
    }
    if (analyse) {
#pragma unroll
        for (int d = 0; d < 8; ++d)
            if (d < dim) {
                const int low = __reduce_min_sync(0xffffffffu, alow[d]);
                const int high = __reduce_max_sync(0xffffffffu, ahigh[d]);
                if (lane == 0) { s_low[warp][d] = low; s_high[warp][d] = high; }
            }
        inexact = __any_sync(0xffffffffu, inexact);
        if (lane == 0) s_bad[warp] = inexact;
    }
    __syncthreads();
    if (threadIdx.x < dim_padded) {
        const int d = threadIdx.x;
        float lo = slo[0][d], hi = shi[0][d];
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w][d]); hi = fmaxf(hi, shi[w][d]); }
        tile_box[(t * dim_padded + d) * 2 + 0] = lo;
        tile_box[(t * dim_padded + d) * 2 + 1] = hi;
    }
    if (!analyse || force_off) {
        if (blockIdx.x == 0 && threadIdx.x == 0) flags[8] = 0;
        return;
    }
    int32_t* rec = exact_info + t * EXACT_INFO_INTS;
    if (threadIdx.x < dim) {
        const int d = threadIdx.x;
        int low = EXACT_NONE_LOW, high = EXACT_NONE_HIGH;
        for (int w = 0; w < 8; ++w) { low = min(low, s_low[w][d]); high = max(high, s_high[w][d]); }
        rec[1 + 2 * d] = low;
        rec[2 + 2 * d] = high;
    }
    if (threadIdx.x == 0) {
        int bad = 0;
        for (int w = 0; w < 8; ++w) bad |= s_bad[w];
        rec[0] = bad;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&flags[9], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // ---- the last CTA folds every tile's record (exact_flag_kernel)
    __threadfence();
    __shared__ int f_low[256], f_high[256], f_ok;
    if (threadIdx.x == 0) f_ok = 1;
    __syncthreads();
    const int64_t n_rec = gridDim.x;
    int bad = 0;
    for (int64_t b = threadIdx.x; b < n_rec; b += 256) bad |= __ldcg(&exact_info[b * EXACT_INFO_INTS]);
    if (bad) f_ok = 0;
    for (int d = 0; d < dim; ++d) {
        int low = EXACT_NONE_LOW, high = EXACT_NONE_HIGH;
        for (int64_t b = threadIdx.x; b < n_rec; b += 256) {
            low = min(low, __ldcg(&exact_info[b * EXACT_INFO_INTS + 1 + 2 * d]));
            high = max(high, __ldcg(&exact_info[b * EXACT_INFO_INTS + 2 + 2 * d]));
        }
        f_low[threadIdx.x] = low;
        f_high[threadIdx.x] = high;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) {
                f_low[threadIdx.x] = min(f_low[threadIdx.x], f_low[threadIdx.x + off]);
                f_high[threadIdx.x] = max(f_high[threadIdx.x], f_high[threadIdx.x + off]);
            }
            __syncthreads();
        }
        // (a coordinate that is all zeros has no bits at all: fine)
        if (threadIdx.x == 0 && f_low[0] != EXACT_NONE_LOW && f_high[0] - f_low[0] > 22) f_ok = 0;
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) flags[8] = f_ok;
}

// ---- space-filling-curve order on per-coordinate RANKS (rank-uniform: cells hold ~equal point counts)
// rank[d][orig] = position of sample `orig` in coordinate d's sorted order
// (grid.y = coordinate; both arrays have row stride ld)
static __global__ void rank_scatter_kernel(const int32_t* __restrict__ sorted_orig, int64_t ld, int64_t n,
                                           int32_t* __restrict__ ranks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)blockIdx.y * ld;
    if (i < n) ranks[row + sorted_orig[row + i]] = (int32_t)i;
}

// key = Hilbert-curve index of the point in rank space: the top `bits` bits of every
// coordinate's rank are taken as integer axes, converted to the "transposed" Hilbert form
// with Skilling's axes-to-transpose bit manipulation (J. Skilling, "Programming the Hilbert
// curve", AIP Conf. Proc. 707, 2004) and bit-interleaved (level-major, MSB first).  Unlike
// the plain Z-order interleave the Hilbert curve has no jumps: any run of consecutive
// points is a connected blob, so 512-point tiles and query blocks get tight boxes.
// X[0..dim) (the top `bits` bits of every coordinate's rank) -> Hilbert index, see above
template <int CAP>
__device__ __forceinline__ unsigned long long hilbert_index(unsigned int (&X)[CAP], int dim, int bits) {
    if (dim > 1) {
        const unsigned int M = 1u << (bits - 1);
        for (unsigned int Q = M; Q > 1u; Q >>= 1) {
            const unsigned int P = Q - 1u;
#pragma unroll
            for (int d = 0; d < CAP; ++d)
                if (d < dim) {
                    if (X[d] & Q) {
                        X[0] ^= P;  // invert the low bits of axis 0
                    } else {
                        const unsigned int t = (X[0] ^ X[d]) & P;  // exchange low bits of axes 0 and d
                        X[0] ^= t;
                        X[d] ^= t;
                    }
                }
        }
#pragma unroll
        for (int d = 1; d < CAP; ++d)
            if (d < dim) X[d] ^= X[d - 1];  // Gray encode
        unsigned int t = 0u;
        for (unsigned int Q = M; Q > 1u; Q >>= 1) {
            unsigned int last = 0u;
#pragma unroll
            for (int d = 0; d < CAP; ++d)
                if (d == dim - 1) last = X[d];
            if (last & Q) t ^= Q - 1u;
        }
#pragma unroll
        for (int d = 0; d < CAP; ++d)
            if (d < dim) X[d] ^= t;
    }
    unsigned long long key = 0ull;
    for (int level = 0; level < bits; ++level)
#pragma unroll
        for (int d = 0; d < CAP; ++d)
            if (d < dim) key = (key << 1) | ((X[d] >> (bits - 1 - level)) & 1u);
    return key;
}

// ---- approximate ranks for the curve (primary = -1, "early order").  The Hilbert order only needs a
// monotone, roughly rank-uniform map of every coordinate -- not its exact ranks -- so the curve keys
// need not wait for the exact coordinate sorts (eight radix passes each): a stratified sample of
// CURVE_SAMPLES values per coordinate is sorted by one CTA, and a value's curve coordinate is its
// position among [min, sample..., max] refined by linear interpolation inside its gap.  The exact sorts
// (needed by the 1-D counts and the band look-ups, i.e. AFTER the k-NN stage) then run on the
// library's side streams underneath the k-NN kernels.  Any monotone map gives a valid, spatially
// coherent order; results never depend on it.
constexpr int CURVE_SAMPLES = 1024;
constexpr int CURVE_TABLE = CURVE_SAMPLES + 2;  // [min, sorted sample, max]

// grid = dim CTAs of 512 threads.  partial: the (min, max) pairs of stats_stage1_kernel.
// table[d] = fp32 [min - c, sorted sample - c, max - c] with c = center[d] = the sample median (centred
// like the shadow copy, so data at a large offset keep their resolution; fp32 rounding is monotone).
static __global__ void __launch_bounds__(512)
curve_table_kernel(const double* __restrict__ pts, int64_t ld, int64_t n, const double* __restrict__ partial,
                   float* __restrict__ table, double* __restrict__ center, double* __restrict__ maxabs) {
    __shared__ uint64_t sk[CURVE_SAMPLES];
    __shared__ uint32_t si[CURVE_SAMPLES];
    __shared__ double s_lo[512], s_hi[512];
    __shared__ double s_center;
    const int d = blockIdx.x, tid = threadIdx.x;
    const double* row = pts + (int64_t)d * ld;
    for (int s = tid; s < CURVE_SAMPLES; s += 512) {
        const uint64_t lo = (uint64_t)s * (uint64_t)n / CURVE_SAMPLES, hi = (uint64_t)(s + 1) * (uint64_t)n / CURVE_SAMPLES;
        uint64_t e = lo + (hi > lo ? ss_hash((uint32_t)s * 2654435761u + 977u * (uint32_t)d) % (hi - lo) : 0);
        if (e >= (uint64_t)n) e = (uint64_t)n - 1;
        sk[s] = ss_key_of(row[e]);
        si[s] = (uint32_t)s;
    }
    double lo = CUDART_INF, hi = -CUDART_INF;
    for (int b = tid; b < STATS_BLOCKS; b += 512) {
        lo = fmin(lo, partial[((int64_t)d * STATS_BLOCKS + b) * 2 + 0]);
        hi = fmax(hi, partial[((int64_t)d * STATS_BLOCKS + b) * 2 + 1]);
    }
    s_lo[tid] = lo;
    s_hi[tid] = hi;
    __syncthreads();
    ss_bitonic(sk, si, CURVE_SAMPLES, tid, 512);
    for (int off = 256; off > 0; off >>= 1) {
        if (tid < off) {
            s_lo[tid] = fmin(s_lo[tid], s_lo[tid + off]);
            s_hi[tid] = fmax(s_hi[tid], s_hi[tid + off]);
        }
        __syncthreads();
    }
    float* out = table + (int64_t)d * CURVE_TABLE;
    if (tid == 0) {
        double c;
        ss_store_key(&c, sk[CURVE_SAMPLES / 2]);  // sample median: the centre of the fp32 copy
        double mn = s_lo[0], mx = s_hi[0];
        if (!isfinite(c)) c = 0.0;  // non-finite input is flagged (stats_stage1_kernel); keep the math finite
        if (!isfinite(mn)) mn = c;
        if (!isfinite(mx)) mx = c;
        out[0] = __double2float_rn(mn - c);
        out[CURVE_TABLE - 1] = __double2float_rn(mx - c);
        center[d] = c;
        maxabs[d] = fmax(fabs(mx - c), fabs(c - mn)) * (1.0 + 1e-12) + 1e-300;
        s_center = c;
    }
    __syncthreads();
    const double c = s_center;
    for (int s = tid; s < CURVE_SAMPLES; s += 512) {
        double v;
        ss_store_key(&v, sk[s]);
        out[1 + s] = __double2float_rn(v - c);
    }
}

// position of the centred value x among the CURVE_TABLE ascending table values, in [0, CURVE_TABLE - 1]:
// the number of entries <= x, refined by linear interpolation inside the gap.  Monotone (non-decreasing) in x:
// the entry count is, and inside a gap x - a, the product with the approximate reciprocal of b - a, the
// clamps and the final sum are all monotone in x under round-to-nearest.
__device__ __forceinline__ int curve_gap(float x, const float* __restrict__ tb) {  // number of table entries <= x
    int lo = 0, hi = CURVE_TABLE;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tb[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ float curve_pos_in_gap(float x, int lo, const float* __restrict__ tb) {  // lo = curve_gap(x, tb)
    float pos = 0.0f;
    if (lo >= CURVE_TABLE) {
        pos = (float)(CURVE_TABLE - 1);
    } else if (lo > 0) {
        const float a = tb[lo - 1], b = tb[lo];
        float f = b > a ? __fdividef(x - a, b - a) : 0.0f;
        if (!(f >= 0.0f)) f = 0.0f;
        if (f > 1.0f) f = 1.0f;
        pos = (float)(lo - 1) + f;
    }
    return pos;
}
__device__ __forceinline__ float curve_pos(float x, const float* __restrict__ tb) {
    return curve_pos_in_gap(x, curve_gap(x, tb), tb);
}
// ... scaled to an integer in [0, 2^bits) (monotone in pos)
__device__ __forceinline__ unsigned int curve_scale(float pos, int bits) {
    const float scaled = pos * ((float)(1u << bits) / (float)(CURVE_TABLE - 1));
    unsigned int c = (unsigned int)scaled;
    const unsigned int top = (1u << bits) - 1u;
    return c > top ? top : c;
}
// curve coordinate of the centred value x in [0, 2^bits): monotone in x (tb: CURVE_TABLE ascending values)
__device__ __forceinline__ unsigned int curve_coord(float x, const float* __restrict__ tb, int bits) {
    return curve_scale(curve_pos(x, tb), bits);
}

// ---- value buckets for the exact 1-D counts (axes_mode > 0, "lazy axes"; see axis_buckets.cuh).
// bucket of a value = its curve position at `bbits` bits (~16 points per bucket for rank-uniform data).
constexpr int BUCKET_MAX_BITS = 22;
constexpr int BUCKET_HEAVY = 512;  // a fuller bucket makes the set "heavy": the bucketed counts give up (flag)
constexpr int BUCKET_CHUNK = 1024; // buckets per CTA of the two scan kernels
// 2^bits buckets, ~4 values each for rank-uniform data (2^bits < n / 2: start, cursor and chunk sums of a
// coordinate fit into its axis_pos row)
static inline int bucket_bits(int64_t n) {
    int rb = 1;
    while ((1ll << rb) < n) ++rb;
    int b = rb > 3 ? rb - 2 : 1;
    return b > BUCKET_MAX_BITS ? BUCKET_MAX_BITS : b;
}

// keys[i] = Hilbert index of point i's curve coordinates.  Dynamic shared memory: dim * CURVE_TABLE floats.
template <int DIM>
static __global__ void __launch_bounds__(256)
curve_key_kernel(const double* __restrict__ pts, int64_t ld, int64_t n, int dim_rt, const float* __restrict__ table,
                 const double* __restrict__ center, int bits, unsigned long long* __restrict__ keys,
                 int32_t* __restrict__ iota, int bbits, int32_t* __restrict__ bucket_hist, int64_t ldh,
                 uint32_t* __restrict__ bucket_id, int64_t ldb) {
    // bucket_hist / bucket_id (optional, lazy axes): bucket_id[d * ldb + i] = the value bucket of coordinate d of
    // point i (its curve position at bbits bits), counted into bucket_hist[d * ldh + bucket]
    extern __shared__ float s_table[];
    const int dim = DIM > 0 ? DIM : dim_rt;
    for (int i = threadIdx.x; i < dim * CURVE_TABLE; i += 256) s_table[i] = table[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    unsigned int X[DIM > 0 ? DIM : KSG_MAX_DIM];
#pragma unroll
    for (int d = 0; d < (DIM > 0 ? DIM : KSG_MAX_DIM); ++d)
        if (d < dim) {
            const float pos = curve_pos(__double2float_rn(pts[(int64_t)d * ld + i] - center[d]), s_table + d * CURVE_TABLE);
            X[d] = curve_scale(pos, bits);
            if (bucket_hist) {
                const unsigned int b = curve_scale(pos, bbits);
                bucket_id[(int64_t)d * ldb + i] = b;
                atomicAdd(&bucket_hist[(int64_t)d * ldh + b], 1);
            }
        }
    keys[i] = hilbert_index(X, dim, bits);
    if (iota) iota[i] = (int32_t)i;
}

// Exclusive prefix sums of the bucket histogram, two launches (grid = (chunks of BUCKET_CHUNK buckets, dim),
// 1024 threads, one bucket per thread):  bucket_sum_kernel: chunk_sum[d][chunk] = values in the chunk, and
// heavy[0] |= 1 when a bucket holds more than BUCKET_HEAVY values (tied data);  bucket_scan_kernel:
// start[d][b] = values in the buckets before b (start[B] = n), cursor[d][b] = start[d][b] in place of the
// histogram (the scatter's bump cursors).
static __device__ __forceinline__ int block_sum_1024(int v, int* s_warp) {  // all threads get the total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    int t = s_warp[lane];
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    return t;
}
static __global__ void __launch_bounds__(BUCKET_CHUNK)
bucket_sum_kernel(const int32_t* __restrict__ hist, int64_t ldh, int B, int32_t* __restrict__ chunk_sum, int64_t ldc,
                  int32_t* __restrict__ heavy) {
    __shared__ int s_warp[32];
    const int b = blockIdx.x * BUCKET_CHUNK + threadIdx.x;
    const int c = b < B ? hist[(int64_t)blockIdx.y * ldh + b] : 0;
    const int total = block_sum_1024(c, s_warp);
    if (threadIdx.x == 0) chunk_sum[(int64_t)blockIdx.y * ldc + blockIdx.x] = total;
    if (__syncthreads_or(c > BUCKET_HEAVY) && threadIdx.x == 0) atomicOr(heavy, 1);
}
static __global__ void __launch_bounds__(BUCKET_CHUNK)
bucket_scan_kernel(int32_t* __restrict__ hist_cursor, int64_t ldh, int32_t* __restrict__ start, int64_t lds, int B,
                   const int32_t* __restrict__ chunk_sum, int64_t ldc) {
    __shared__ int s_warp[32], s_scan[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // values in the chunks before this one
    int before = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += BUCKET_CHUNK) before += chunk_sum[(int64_t)blockIdx.y * ldc + i];
    before = block_sum_1024(before, s_warp);
    const int b = blockIdx.x * BUCKET_CHUNK + threadIdx.x;
    int32_t* h = hist_cursor + (int64_t)blockIdx.y * ldh;
    const int c = b < B ? h[b] : 0;
    int incl = c;  // inclusive scan over the CTA: warp scans, then the warp totals
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    __syncthreads();
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int wt = s_scan[lane];
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wt, off);
        if (lane >= off) wt += v;
    }
    const int warp_before = __shfl_sync(0xffffffffu, wt, warp) - s_scan[warp];
    const int excl = before + warp_before + incl - c;
    if (b < B) {
        start[(int64_t)blockIdx.y * lds + b] = excl;
        h[b] = excl;
        if (b == B - 1) start[(int64_t)blockIdx.y * lds + B] = excl + c;
    }
}

// vals[d][cursor[d][bucket]++] = pts[d][i]: the values of every coordinate grouped by bucket (any order inside
// a bucket: the counts do not depend on it).  grid = (blocks, dim), four points per thread (the atomics are
// latency-bound: four independent ones in flight).
constexpr int BUCKET_SCATTER_PER_THREAD = 4;
static __global__ void __launch_bounds__(256)
bucket_scatter_kernel(const double* __restrict__ pts, int64_t ld_in, int64_t n, const uint32_t* __restrict__ bucket_id,
                      int64_t ldb, int32_t* __restrict__ cursor, int64_t ldh, double* __restrict__ vals, int64_t ldv) {
    const int d = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * (256 * BUCKET_SCATTER_PER_THREAD) + threadIdx.x;
    unsigned int b[BUCKET_SCATTER_PER_THREAD];
    double v[BUCKET_SCATTER_PER_THREAD];
    int pos[BUCKET_SCATTER_PER_THREAD];
#pragma unroll
    for (int u = 0; u < BUCKET_SCATTER_PER_THREAD; ++u) {
        const int64_t i = i0 + u * 256;
        b[u] = i < n ? bucket_id[(int64_t)d * ldb + i] : 0u;
        v[u] = i < n ? pts[(int64_t)d * ld_in + i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < BUCKET_SCATTER_PER_THREAD; ++u)
        pos[u] = i0 + u * 256 < n ? atomicAdd(&cursor[(int64_t)d * ldh + b[u]], 1) : 0;
#pragma unroll
    for (int u = 0; u < BUCKET_SCATTER_PER_THREAD; ++u)
        if (i0 + u * 256 < n) vals[(int64_t)d * ldv + pos[u]] = v[u];
}

// bits per coordinate of the curve for n points in `dim` coordinates (shared by ksg_prepare and
// ksg_cross_prepare, which must place foreign points on the SAME curve)
// The curve only has to resolve the points down to a sub-tile: with b = rank_bits - 4 bits per
// coordinate a rank slice holds <= 16 points, so even when the coordinates are functions of each
// other (all points of a slice in ONE cell) a cell is smaller than a 32-row sub-tile; everything
// finer would only lengthen the key sort (one radix pass per 8 bits of dim * b).
// KSG_B200_KEY_BITS (read once): cap on the TOTAL width of a curve key.  The key sort is latency-bound radix
// passes of 8 bits each; a coarser curve (ties keep their input order: the sort is stable) trades passes
// against slightly wider sub-tile boxes.  Any order gives the same radii and counts.
static inline int curve_key_bits_cap() {
    static const int cap = [] {
        const char* e = getenv("KSG_B200_KEY_BITS");
        return e ? atoi(e) : 0;
    }();
    return cap;
}
static inline void curve_bits(int64_t n, int dim, int* rank_bits, int* bits) {
    int rb = 1;
    while ((1ll << rb) < n) ++rb;
    int b = 63 / dim;
    const int need = rb > 5 ? rb - 4 : 1;
    if (b > need) b = need;
    // Scalar pairs (the headline step): 16x as many curve cells as points is enough -- at n = 1e6 a 24-bit key
    // (three radix passes) against 32 bits (four) costs 0.9 % more evaluated pairs and saves 22 us of the
    // 225 us prepare (profiles/r2t_key_bits.txt); 20 bits: +15 % pairs.  More coordinates keep the full width
    // (prepare is negligible against their pair work, and coarse cells hurt where the points are dense).
    int cap = curve_key_bits_cap();
    if (cap <= 0 && dim <= 2) cap = rb + 4 > 24 ? rb + 4 : 24;
    if (cap > 0 && b * dim > cap) b = cap / dim > 1 ? cap / dim : 1;
    *rank_bits = rb;
    *bits = b;
}

template <int DIM>  // DIM > 0: compile-time dimension (coordinates stay in registers); 0: run-time `dim`
static __global__ void morton_key_kernel(const int32_t* __restrict__ ranks, int64_t ld, int64_t n, int dim_rt,
                                         int rank_bits, int bits, unsigned long long* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dim = DIM > 0 ? DIM : dim_rt;
    unsigned int X[DIM > 0 ? DIM : KSG_MAX_DIM];
#pragma unroll
    for (int d = 0; d < (DIM > 0 ? DIM : KSG_MAX_DIM); ++d)
        if (d < dim) X[d] = ((unsigned int)ranks[(int64_t)d * ld + i]) >> (rank_bits - bits);
    keys[i] = hilbert_index(X, dim, bits);
}

// ---- cross-set queries (ksg_cross_prepare): foreign points placed on the candidate set's curve.
// rank of a query coordinate = number of candidate values below it (binary search in the
// coordinate's sorted array), clamped into the candidates' rank range.
template <int DIM>
static __global__ void cross_key_kernel(const double* __restrict__ qry, int64_t ldq, int64_t nq,
                                        const double* __restrict__ axis_vals, int64_t ld, int64_t n, int dim_rt,
                                        int rank_bits, int bits, unsigned long long* __restrict__ keys,
                                        int32_t* __restrict__ iota) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int dim = DIM > 0 ? DIM : dim_rt;
    unsigned int X[DIM > 0 ? DIM : KSG_MAX_DIM];
#pragma unroll
    for (int d = 0; d < (DIM > 0 ? DIM : KSG_MAX_DIM); ++d)
        if (d < dim) {
            const double x = qry[(int64_t)d * ldq + i];
            const double* v = axis_vals + (int64_t)d * ld;
            int64_t lo = 0, hi = n;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (v[mid] < x) lo = mid + 1; else hi = mid;
            }
            if (lo > n - 1) lo = n - 1;
            X[d] = ((unsigned int)lo) >> (rank_bits - bits);
        }
    keys[i] = hilbert_index(X, dim, bits);
    iota[i] = (int32_t)i;
}

// queries in cross order: float64 SoA copy, fp32 rows centred with the CANDIDATE set's centre
// (+inf padding rows), and the candidate row next to each query on the curve
static __global__ void __launch_bounds__(256)
cross_gather_kernel(const double* __restrict__ qry, int64_t ldq, int64_t nq, int dim, int dim_padded,
                    const int32_t* __restrict__ perm, const unsigned long long* __restrict__ keys_sorted,
                    const double* __restrict__ center, const unsigned long long* __restrict__ cand_keys, int64_t n,
                    double* __restrict__ sorted, int64_t ld, float* __restrict__ shadow, int32_t* __restrict__ home) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= nq) return;
    const int64_t src = perm[i];
    for (int d = 0; d < dim_padded; ++d) {
        float s = 0.0f;
        if (d < dim) {
            const double v = qry[(int64_t)d * ldq + src];
            sorted[(int64_t)d * ld + i] = v;
            s = __double2float_rn(__dsub_rn(v, center[d]));
        }
        shadow[i * dim_padded + d] = s;
    }
    const unsigned long long key = keys_sorted[i];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cand_keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    home[i] = (int32_t)(lo > n - 1 ? n - 1 : lo);
}

// maxabs2[d] = max(candidates' maxabs[d], max over the queries of |x - centre|) (rounded up)
static __global__ void cross_maxabs_kernel(const double* __restrict__ partial, int dim, const double* __restrict__ center,
                                           const double* __restrict__ maxabs, double* __restrict__ out) {
    const int d = threadIdx.x;
    if (d >= dim) return;
    double lo = CUDART_INF, hi = -CUDART_INF;
    for (int b = 0; b < STATS_BLOCKS; ++b) {
        lo = fmin(lo, partial[((int64_t)d * STATS_BLOCKS + b) * 2 + 0]);
        hi = fmax(hi, partial[((int64_t)d * STATS_BLOCKS + b) * 2 + 1]);
    }
    if (!(isfinite(lo) && isfinite(hi))) { lo = center[d]; hi = center[d]; }
    const double m = fmax(hi - center[d], center[d] - lo) * (1.0 + 1e-12) + 1e-300;
    out[d] = fmax(maxabs[d], m);
}

// ---- kd order (primary = -2): the order of the leaves of a balanced kd-tree.  Level after level,
// every segment (a contiguous run of positions, boundaries at multiples of 32 rows) is sorted along
// one coordinate -- by global rank, cycling through the coordinates -- and cut in half at a 32-row
// boundary, until a segment is one 32-row sub-tile.  Unlike the curve over GLOBAL ranks the splits
// are CONDITIONAL medians, so the boxes follow the data when coordinates are correlated or the
// points sit near a low-dimensional manifold: sub-tiles, tiles and query blocks are nodes of the
// tree with tight, non-overlapping boxes.  key = (segment of the position at this level, rank).
static __global__ void kd_key_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ rank_row, int64_t n,
                                     int level, int64_t n_groups, unsigned long long* __restrict__ keys) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t g = p / FAST_SUB;
    int64_t lo = 0, hi = n_groups;
    unsigned long long id = 0ull;
    for (int l = 0; l < level; ++l) {
        const int64_t mid = lo + ((hi - lo + 1) >> 1);
        if (g < mid) { hi = mid; id <<= 1; } else { lo = mid; id = (id << 1) | 1ull; }
    }
    keys[p] = (id << 32) | (unsigned long long)(unsigned int)rank_row[perm[p]];
}

static __global__ void invert_perm_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int32_t)i;
}

// axis_pos[d][i] (an original sample index) -> its position in the final order (grid.y = coordinate)
static __global__ void remap_axis_kernel(int32_t* __restrict__ pos, int64_t ld, int64_t n,
                                         const int32_t* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int32_t* p = pos + (int64_t)blockIdx.y * ld + i;
        *p = inv[*p];
    }
}

// tile_box[(t * dim_padded + d) * 2 + {0,1}] = min / max of the fp32 shadow over tile t;
// sub_box[((t * FAST_NSUB + s) * dim_padded + d) * 2 + {0,1}] = the same over sub-tile s of tile t
static __global__ void __launch_bounds__(128)
tile_box_kernel(const float* __restrict__ shadow, int dim_padded, float* __restrict__ tile_box,
                float* __restrict__ sub_box) {
    const int64_t t = blockIdx.x;
    const float* rows = shadow + t * FAST_TILE * dim_padded;
    __shared__ float slo[4][KSG_MAX_DIM], shi[4][KSG_MAX_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 0; d < dim_padded; ++d) {
        float tlo = CUDART_INF_F, thi = -CUDART_INF_F;
        for (int s = warp; s < FAST_NSUB; s += 4) {
            const float v = rows[(s * FAST_SUB + lane) * dim_padded + d];
            float lo = v, hi = v;
            for (int off = 16; off > 0; off >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
            }
            if (lane == 0) {
                sub_box[((t * FAST_NSUB + s) * dim_padded + d) * 2 + 0] = lo;
                sub_box[((t * FAST_NSUB + s) * dim_padded + d) * 2 + 1] = hi;
            }
            tlo = fminf(tlo, lo);
            thi = fmaxf(thi, hi);
        }
        if (lane == 0) { slo[warp][d] = tlo; shi[warp][d] = thi; }
    }
    __syncthreads();
    if (threadIdx.x < dim_padded) {
        const int d = threadIdx.x;
        tile_box[(t * dim_padded + d) * 2 + 0] = fminf(fminf(slo[0][d], slo[1][d]), fminf(slo[2][d], slo[3][d]));
        tile_box[(t * dim_padded + d) * 2 + 1] = fmaxf(fmaxf(shi[0][d], shi[1][d]), fmaxf(shi[2][d], shi[3][d]));
    }
}

}  // namespace ksg

// Exact float64 dense kernels: the correctness anchor of the path and the permanent
// device-side fallback for queries the fp32 filter cannot decide (tie-saturated or
// quantised data).  Everything here evaluates max_d |fl64(q_d - p_d)| exactly as the
// reference's cKDTree does, so radii and counts are bit-exact by construction.
//
// Layout: candidates are staged tile by tile into shared memory as SoA rows
// (row d of the tile = coordinate d of TILE consecutive points); one thread owns one
// query whose coordinates stay in registers; all lanes read the same candidate, so
// shared-memory reads are broadcasts.  blockIdx.y splits the candidate range so that
// small query shards still fill 148 SMs.
#pragma once

#include "common.cuh"

namespace ksg {

constexpr int F64_THREADS = 128;
constexpr int F64_TILE = 512;

// DT > 0: compile-time dimensionality (registers, fully unrolled).
// DT == 0: run-time dimensionality up to KSG_MAX_DIM (local array, generic path).
template <int DT>
struct DimOf {
    static constexpr int CAP = DT > 0 ? DT : KSG_MAX_DIM;
    __device__ __forceinline__ static int get(int rt) { return DT > 0 ? DT : rt; }
};

__device__ __forceinline__ void stage_tile_f64(double* __restrict__ tile, const double* __restrict__ pts,
                                               int64_t ld, int64_t base, int64_t cnt, int dim, int tile_len,
                                               double pad) {
    for (int i = threadIdx.x; i < dim * tile_len; i += blockDim.x) {
        int d = i / tile_len;
        int j = i - d * tile_len;
        tile[i] = (j < cnt) ? pts[(int64_t)d * ld + base + j] : pad;
    }
}

// ---------------------------------------------------------------------------------- K1
// part_dist[(split * kp + r) * nq + q] = r-th smallest distance seen by this split.
template <int DT, bool WITH_IDX>
static __global__ void __launch_bounds__(F64_THREADS)
knn_f64_kernel(const double* __restrict__ pts, int64_t ld, int64_t n,
               const double* __restrict__ qry, int64_t ldq, int64_t q_begin, int64_t nq,
               int dim_rt, int kp, int64_t cand_per_split,
               double* __restrict__ part_dist, int64_t* __restrict__ part_idx) {
    extern __shared__ double smem_f64[];
    const int dim = DimOf<DT>::get(dim_rt);
    double* tile = smem_f64;                       // dim * F64_TILE
    double* lst = tile + dim * F64_TILE;           // kp * F64_THREADS  (rank-major)
    int64_t* lsti = (int64_t*)(lst + kp * F64_THREADS);

    const int tid = threadIdx.x;
    const int64_t qi = (int64_t)blockIdx.x * F64_THREADS + tid;
    const bool active = qi < nq;

    double q[DimOf<DT>::CAP];
#pragma unroll
    for (int d = 0; d < DimOf<DT>::CAP; ++d)
        if (d < dim) q[d] = active ? qry[(int64_t)d * ldq + q_begin + qi] : 0.0;

    for (int r = 0; r < kp; ++r) {
        lst[r * F64_THREADS + tid] = CUDART_INF;
        if (WITH_IDX) lsti[r * F64_THREADS + tid] = n;  // cKDTree's "missing neighbour" index
    }
    double thr = CUDART_INF;  // current kp-th smallest

    const int64_t c0 = (int64_t)blockIdx.y * cand_per_split;
    const int64_t c1 = min(n, c0 + cand_per_split);

    for (int64_t base = c0; base < c1; base += F64_TILE) {
        const int64_t cnt = min((int64_t)F64_TILE, c1 - base);
        __syncthreads();
        stage_tile_f64(tile, pts, ld, base, cnt, dim, F64_TILE, CUDART_INF);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < F64_TILE; ++j) {
            // fp64 max is a 3-instruction sequence on sm_100 (DSETP + 2 FSEL); the hot
            // test is a chain of predicated compares instead, the max is formed on a hit.
            // strict '<': a candidate tying the current kp-th value cannot change it.
            double a[DimOf<DT>::CAP];
            bool hit = true;
#pragma unroll
            for (int d = 0; d < DimOf<DT>::CAP; ++d)
                if (d < dim) {
                    a[d] = fabs(q[d] - tile[d * F64_TILE + j]);
                    hit = hit && (a[d] < thr);
                }
            if (hit) {
                double m = a[0];
#pragma unroll
                for (int d = 1; d < DimOf<DT>::CAP; ++d)
                    if (d < dim) m = fmax(m, a[d]);
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
            int64_t o = ((int64_t)blockIdx.y * kp + r) * nq + qi;
            part_dist[o] = lst[r * F64_THREADS + tid];
            if (WITH_IDX) part_idx[o] = lsti[r * F64_THREADS + tid];
        }
    }
}

// Merge the per-split lists: kp-th smallest of the union (and the kp (dist, idx)
// pairs in ascending lexicographic order when indices are requested).
template <bool WITH_IDX>
static __global__ void knn_merge_kernel(const double* __restrict__ part_dist, const int64_t* __restrict__ part_idx,
                                 int64_t nq, int kp, int splits,
                                 double* __restrict__ eps_out, int64_t* __restrict__ idx_out) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    double best[KSG_MAX_KPRIME];
    int64_t besti[KSG_MAX_KPRIME];
    int have = 0;
    for (int s = 0; s < splits; ++s) {
        for (int r = 0; r < kp; ++r) {
            int64_t o = ((int64_t)s * kp + r) * nq + qi;
            double v = part_dist[o];
            int64_t vi = WITH_IDX ? part_idx[o] : 0;
            int pos = have < kp ? have : kp - 1;
            if (have == kp) {
                bool better = v < best[kp - 1] || (WITH_IDX && v == best[kp - 1] && vi < besti[kp - 1]);
                if (!better) break;  // lists are ascending: nothing further in this split helps
            }
            while (pos > 0 && (best[pos - 1] > v || (WITH_IDX && best[pos - 1] == v && besti[pos - 1] > vi))) {
                best[pos] = best[pos - 1];
                besti[pos] = besti[pos - 1];
                --pos;
            }
            best[pos] = v;
            besti[pos] = vi;
            if (have < kp) ++have;
        }
    }
    eps_out[qi] = best[kp - 1];
    if (WITH_IDX)
        for (int r = 0; r < kp; ++r) idx_out[qi * kp + r] = besti[r];
}

// ---------------------------------------------------------------------------------- K2
constexpr int F64_SUBCHUNK = 4;  // subspaces handled per launch slice (blockIdx.z)

struct MaskTable {
    uint32_t m[KSG_MAX_SUBSPACES];
};

static __global__ void fill_i32_kernel(int32_t* __restrict__ p, int64_t count, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

// counts[(s0 + s) * nq + q] += #{candidates in this split inside the ball of subspace s}
template <int DT>
static __global__ void __launch_bounds__(F64_THREADS)
count_f64_kernel(const double* __restrict__ pts, int64_t ld, int64_t n,
                 const double* __restrict__ qry, int64_t ldq, int64_t q_begin, int64_t nq,
                 int dim_rt, const MaskTable masks, int n_sub,
                 const double* __restrict__ eps, int64_t eps_stride, int strict,
                 int64_t cand_per_split, int32_t* __restrict__ counts) {
    extern __shared__ double smem_f64[];
    const int dim = DimOf<DT>::get(dim_rt);
    double* tile = smem_f64;

    const int tid = threadIdx.x;
    const int64_t qi = (int64_t)blockIdx.x * F64_THREADS + tid;
    const bool active = qi < nq;
    const int s0 = blockIdx.z * F64_SUBCHUNK;

    uint32_t mk[F64_SUBCHUNK];
    double e[F64_SUBCHUNK];
    int cnt[F64_SUBCHUNK];
#pragma unroll
    for (int s = 0; s < F64_SUBCHUNK; ++s) {
        const bool on = s0 + s < n_sub;
        mk[s] = on ? masks.m[s0 + s] : 0u;
        // an unused slot, or an inactive lane, gets a radius nothing is inside of
        e[s] = (on && active) ? eps[(int64_t)(s0 + s) * eps_stride + qi] : -1.0;
        cnt[s] = 0;
    }
    double q[DimOf<DT>::CAP];
#pragma unroll
    for (int d = 0; d < DimOf<DT>::CAP; ++d)
        if (d < dim) q[d] = active ? qry[(int64_t)d * ldq + q_begin + qi] : 0.0;

    const int64_t c0 = (int64_t)blockIdx.y * cand_per_split;
    const int64_t c1 = min(n, c0 + cand_per_split);

    for (int64_t base = c0; base < c1; base += F64_TILE) {
        const int jn = (int)min((int64_t)F64_TILE, c1 - base);
        __syncthreads();
        stage_tile_f64(tile, pts, ld, base, jn, dim, F64_TILE, 0.0);
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < jn; ++j) {
            double a[DimOf<DT>::CAP];
#pragma unroll
            for (int d = 0; d < DimOf<DT>::CAP; ++d)
                if (d < dim) a[d] = fabs(q[d] - tile[d * F64_TILE + j]);
#pragma unroll
            for (int s = 0; s < F64_SUBCHUNK; ++s) {
                // max_{d in s} a_d < e  <=>  every a_d of the subspace is < e
                bool in = true;
#pragma unroll
                for (int d = 0; d < DimOf<DT>::CAP; ++d)
                    if (d < dim && ((mk[s] >> d) & 1u)) in = in && (strict ? (a[d] < e[s]) : (a[d] <= e[s]));
                cnt[s] += in;
            }
        }
    }
    if (active) {
#pragma unroll
        for (int s = 0; s < F64_SUBCHUNK; ++s)
            if (s0 + s < n_sub && cnt[s]) atomicAdd(&counts[(int64_t)(s0 + s) * nq + qi], cnt[s]);
    }
}

// ---------------------------------------------------------------------------------- KSG-2 helper
static __global__ void subspace_radius_kernel(const double* __restrict__ pts, int64_t ld, int dim,
                                       int64_t q_begin, int64_t nq, int64_t n,
                                       const int64_t* __restrict__ idx, int kp, int skip_first,
                                       const MaskTable masks, int n_sub,
                                       double* __restrict__ eps_out) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    for (int s = 0; s < n_sub; ++s) {
        const uint32_t mk = masks.m[s];
        double e = -CUDART_INF;
        for (int r = skip_first ? 1 : 0; r < kp; ++r) {
            const int64_t nb = idx[qi * kp + r];
            if (nb < 0 || nb >= n) { e = CUDART_INF; continue; }  // missing neighbour (kprime > n)
            double m = 0.0;
            for (int d = 0; d < dim; ++d)
                if ((mk >> d) & 1u)
                    m = fmax(m, fabs(pts[(int64_t)d * ld + q_begin + qi] - pts[(int64_t)d * ld + nb]));
            e = fmax(e, m);
        }
        eps_out[(int64_t)s * nq + qi] = e;
    }
}

// ---------------------------------------------------------------------------------- K4
constexpr int PC_THREADS = 128;
constexpr int PC_TILE = 1024;

// MT > 0: compile-time small-template length m; MT == 0: run-time m (<= KSG_MAX_DIM - 1).
template <int MT>
static __global__ void __launch_bounds__(PC_THREADS)
paircount_f64_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t nt, int m_rt,
                     double r, int64_t i_begin, int64_t ni, int64_t cand_per_split,
                     unsigned long long* __restrict__ out) {
    constexpr int CAP = MT > 0 ? MT + 1 : KSG_MAX_DIM;
    const int m = MT > 0 ? MT : m_rt;
    __shared__ double ys[PC_TILE + KSG_MAX_DIM];
    __shared__ unsigned long long red[2];

    const int tid = threadIdx.x;
    const int64_t ii = (int64_t)blockIdx.x * PC_THREADS + tid;
    const bool active = ii < ni;
    double xi[CAP];
#pragma unroll
    for (int w = 0; w < CAP; ++w)
        if (w <= m) xi[w] = active ? x[i_begin + ii + w] : 0.0;

    if (tid < 2) red[tid] = 0ull;
    unsigned long long small_total = 0ull, big_total = 0ull;

    const int64_t c0 = (int64_t)blockIdx.y * cand_per_split;
    const int64_t c1 = min(nt, c0 + cand_per_split);
    for (int64_t base = c0; base < c1; base += PC_TILE) {
        const int jn = (int)min((int64_t)PC_TILE, c1 - base);
        __syncthreads();
        for (int i = tid; i < jn + m; i += PC_THREADS) ys[i] = y[base + i];
        __syncthreads();
        unsigned int cs = 0, cb = 0;
#pragma unroll 2
        for (int j = 0; j < jn; ++j) {
            // max_w |.| <= r  <=>  every |.| <= r  (no fp64 max in the hot loop)
            bool sm = fabs(xi[0] - ys[j]) <= r;
#pragma unroll
            for (int w = 1; w < CAP - 1; ++w)
                if (w < m) sm = sm && (fabs(xi[w] - ys[j + w]) <= r);
            bool bg = sm;
#pragma unroll
            for (int w = 1; w < CAP; ++w)
                if (w == m) bg = sm && (fabs(xi[w] - ys[j + w]) <= r);
            cs += sm;
            cb += bg;
        }
        small_total += cs;
        big_total += cb;
    }
    if (!active) { small_total = 0ull; big_total = 0ull; }
    // warp reduce, then one shared atomic per warp, one global atomic per block
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

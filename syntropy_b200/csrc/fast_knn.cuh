// K1 fast path: k-NN radius with an fp32 tile filter and an exact fp64 refine.
//
// Both filters work on the prepared set (prepare.cuh): fp32 AoS shadow rows in space-filling
// curve order, +inf padded, with a bounding box per 512-row tile and per 32-row sub-tile.  One
// thread owns Q queries whose coordinates live in registers as packed pairs; candidate tiles are
// bulk-copied (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) into a ring of
// shared-memory buffers; all lanes read the same candidate (LDS.128 broadcasts).
//
// Dense filter (knn_f32_kernel): every tile.  Per (query, candidate) DP/2 FADD2 (packed
// subtract, two dimensions per issue slot) + ~DP/2 FMNMX/FMNMX3 (|.| folded into the operands)
// + 1/2 FMNMX3 (running minimum over two candidates) -- no compare, no branch.  Every BATCH
// candidates the running minima are tested against the per-query thresholds; only on a hit is
// the batch re-walked to insert (d32, index) into the query's sorted list in shared memory.
// This is the brute-force roofline case, and the path of explicit query lists.
//
// Windowed filter (knn_plan_kernel -> knn_unit_table_kernel -> knn_units_kernel): same
// result from ~1e-3 of the pairs.  Planning seeds every query's threshold from its neighbours
// on the curve and lists, per block of queries, the tiles whose box is within reach; the lists
// are cut into work units pulled by a persistent kernel (4 filter warps + 1 copy warp per CTA);
// each filter warp culls 32-row sub-tiles against the box of its own 32*Q consecutive queries.
// The same kernels serve cross-set searches (queries that are not candidates: qshadow / qhome).
//
// Refine (knn_refine_kernel).  The lists keep the KCAP = k' + margin smallest fp32 distances.
// With r32 the k'-th smallest of them and delta the fp32 error bound, every point whose exact
// distance is <= the exact radius has d32 <= r32 + 2*delta; if the lists provably contain all
// of those (their tails lie above that bound) the exact radius is the k'-th smallest float64
// distance among them.  Otherwise the query is flagged and recomputed by the dense fp64 kernel.
#pragma once

#include "common.cuh"
#include "fast_launch.h"
#include "prepare.cuh"
#include "tile_queue.cuh"

namespace ksg {

constexpr int KNN_BATCH = 16;   // candidates between threshold checks
constexpr int KNN_MARGIN = 3;   // list entries kept beyond k'
constexpr int KNN_MAX_KCAP = 64; // longest neighbour list of the fp32 filter (k' + margin)
constexpr int KNN_SEED_W = 24;  // rows on each side of a query used to seed its threshold
constexpr float KNN_DEFER_CUT = 4.0f;   // defer a query whose seeded radius exceeds CUT x the CTA mean
constexpr float KNN_DEFERRED = -2.0f;   // sentinel distance marking a deferred query's list
constexpr int KNN_PLAN_CAP = 1024;      // tile-list pool: this many entries per query block ON AVERAGE
constexpr int KNN_UNIT_MIN_TILES = 16;  // a work unit = one query block x at most `chunk` planned tiles
constexpr int COUNT_UNIT_MIN_TILES = 64;  // count units are cheap per tile (culled / inside sub-tiles): a unit
                                          // must be long enough to amortise the latency of its first bulk copy
constexpr int KNN_UNIT_FACTOR = 2;      // chunk grows so that #units <= (1 + FACTOR) * #blocks

// fp32 error bound of a (query, candidate) pair, PER QUERY.  With centred values q~, c~
// (|q~_d| <= Mq) and a candidate within d of the query, |c~_d| <= Mq + d, so
//     |d32 - d64| <= 2^-24 (|q~| + |c~| + |d32|) <= 2^-24 (2 Mq + 2 d) (1 + tiny).
// Using the query's own magnitude instead of the data set's keeps the band narrow where the
// points are dense (near the centre = the per-coordinate median) even when a few outliers stretch
// the range by orders of magnitude (heavy-tailed data).
__host__ __device__ inline double band_query_f64(double mq, double d) {
    return 5.9604644775390625e-08 * (2.0 * mq + 2.0 * d) * 1.01 + 1e-37;
}

// The fp32 shadow holds x - centre: finite float64 input farther than FLT_MAX from the median would
// turn into +-inf rows that look like padding.  maxabs[d] = max |x_d - centre_d| of the prepared set
// (candidates and, for cross-set searches, queries): when it leaves the fp32 range every query is
// handed to the float64 kernels instead (cKDTree handles such input; so must we).
__host__ __device__ inline bool fp32_range_ok(const double* maxabs, int dim) {
    bool ok = true;
    for (int d = 0; d < dim; ++d) ok = ok && maxabs[d] < 3.0e38;
    return ok;
}

// Tight seed of the windowed filter: s = the k'-th smallest fp32 distance of the seed window (an upper
// bound of the final r32), mq = max_d |q~_d| of the fp32 row.  The refine examines every candidate
// with d32 <= T = r32 + 2 * band_query_f64(mq64, 2 * r32) = r32 + 2^-23 (2 mq64 + 4 r32) * 1.01 (+2e-37),
// mq64 <= mq (1 + 2e-6).  The value returned here exceeds that for every r32 <= s (factor 1.024
// against 1.01, every operation rounded up), so "inserted iff d32 < seed" captures them all; the
// refine still verifies T < seed and flags the query otherwise.
__device__ __forceinline__ float tight_seed(float s, float mq) {
    const float w = __fmul_ru(1.2207031e-07f, __fadd_ru(__fadd_ru(mq, mq), __fmul_ru(4.0f, s)));  // 1.024 * 2^-23 (2 mq + 4 s)
    return __fadd_ru(s, __fadd_ru(w, 1e-36f));  // inf stays inf
}

// ---- packed helpers
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "sub.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

template <int DP>
__device__ __forceinline__ float cheb32(const float2 (&q)[DP / 2], const float2 (&c)[DP / 2]) {
    const float2 d0 = sub2(q[0], c[0]);
    float m = fmaxf(fabsf(d0.x), fabsf(d0.y));
#pragma unroll
    for (int h = 1; h < DP / 2; ++h) {
        const float2 d = sub2(q[h], c[h]);
        m = fmaxf(m, fmaxf(fabsf(d.x), fabsf(d.y)));
    }
    return m;
}

template <int DP>
__device__ __forceinline__ void load_row(const float* __restrict__ row, float2 (&c)[DP / 2]) {
    if constexpr (DP % 4 == 0) {
#pragma unroll
        for (int h = 0; h < DP / 4; ++h) {
            const float4 v = reinterpret_cast<const float4*>(row)[h];
            c[2 * h] = make_float2(v.x, v.y);
            c[2 * h + 1] = make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) c[h] = reinterpret_cast<const float2*>(row)[h];
    }
}

// two consecutive rows (for DP = 2 they share one 16-byte load)
template <int DP>
__device__ __forceinline__ void load_two_rows(const float* __restrict__ row, float2 (&a)[DP / 2], float2 (&b)[DP / 2]) {
    if constexpr (DP == 2) {
        const float4 v = *reinterpret_cast<const float4*>(row);
        a[0] = make_float2(v.x, v.y);
        b[0] = make_float2(v.z, v.w);
    } else {
        load_row<DP>(row, a);
        load_row<DP>(row + DP, b);
    }
}

constexpr int PLAN_MASK_WORDS = 1024;  // survivor bitmask of the planning pass: up to 32768 tiles (16.7 M points)

// ---- planning pass of the windowed engine.  One CTA per query block: seeds every query's
// threshold from its neighbours in the prepared order, derives the block's pruning radius, tests
// EVERY candidate tile's bounding box against the block's query box and writes the survivors
// (tile, gap) in walk order.  The filter kernel then streams exactly that list (re-checking each
// gap against its shrinking radius) with no in-kernel search rounds, and the host side of the ABI
// launches the blocks longest-list-first so that no heavy block is left for the tail.
template <int DP, int Q, int KC>  // KC > 0: seed lists of KC entries in registers (kcap <= KC); 0: in shared memory
static __global__ void __launch_bounds__(FAST_THREADS)
knn_plan_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box, int64_t n_tiles,
                int64_t q_begin, int64_t nq, int kcap, int kseed, int seed_w, const double* __restrict__ maxabs, int dim,
                const float* __restrict__ qshadow, const int32_t* __restrict__ qhome,
                float* __restrict__ seed_thr, int2* __restrict__ plan, int64_t plan_capacity,
                int32_t* __restrict__ plan_count, int32_t* __restrict__ plan_first, int32_t* __restrict__ meta,
                int32_t* __restrict__ unit_done) {
    // kseed: rank of the seed distance.  kseed == kcap (legacy): the seed is the kcap-th smallest
    // distance of the window, one ulp up -- the window's own rows fill the list.  kseed == k' <
    // kcap (tight): the seed is the k'-th smallest window distance s plus the width of the refine
    // kernel's error band at s (tight_seed below), so the filter keeps exactly the candidates the
    // refine needs -- ~k' insertions per query instead of ~kcap-and-replacements, and a block
    // radius that is the k'-th, not the kcap-th, neighbour distance.
    // qshadow / qhome (optional, cross-set search): the queries are rows [q_begin, q_begin + nq) of
    // qshadow -- foreign points in the candidate set's centred frame, ordered along the candidates'
    // curve -- and qhome[row] is the candidate row next to each of them on the curve: the seeds come
    // from the candidates around it (the home row included: it is not the query itself).
    const float* qrows = qshadow ? qshadow : shadow;
    const int self_skip = qshadow ? 0 : 1;
    constexpr int QB = FAST_THREADS * Q;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* lst_d = reinterpret_cast<float*>(smem_raw);  // kcap * QB (values only)
    __shared__ TileQueueShared sh;
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t qb0 = (int64_t)blockIdx.x * QB;
    (void)maxabs; (void)dim;
    const int64_t n_rows = n_tiles * FAST_TILE;
    // finished work units of this block: [0] for the CTA-level kernels, one counter per 32-query quarter for
    // the warp-level kernel
    if (tid < 4 && unit_done) unit_done[4 * blockIdx.x + tid] = 0;

    // ---- pass 1: seed every query's threshold from its neighbours in the prepared order
    float blo[DP], bhi[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
    float sum = 0.0f, cnt = 0.0f;
    // Within-set searches with register-resident seed lists read their window from shared memory: the
    // rows [first query - W, last query + W] of the block are staged by one coalesced sweep, and the
    // 2 W candidate rows of every query come from there instead of 2 W dependent L1 / L2 round trips.
    constexpr int WIN_ROWS = FAST_THREADS + 2 * KNN_SEED_W;
    __shared__ __align__(16) float s_win[KC > 0 ? WIN_ROWS * DP : 1];
    // seed_w: rows on each side of a query that seed its threshold.  Any width gives a valid upper
    // bound; from ~5 coordinates up consecutive rows of the curve are a poor sample of a query's
    // neighbourhood (seed / true radius 1.5 at +-24 rows in 6-D and 8-D, 1.2 at +-384:
    // profiles/r1aj_seed_window_simulation.txt), so those sets use a wider window (global loads).
    const bool staged = KC > 0 && !qhome && !qshadow && seed_w == KNN_SEED_W;
#pragma unroll 1
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if constexpr (KC > 0) {
            if (staged) {
                if (j > 0) __syncthreads();  // the previous slot's reads are done
                const int64_t base_row = q_begin + qb0 + (int64_t)j * FAST_THREADS - KNN_SEED_W;
                for (int i = tid; i < WIN_ROWS * (DP / 2); i += FAST_THREADS) {
                    const int64_t r2 = base_row + i / (DP / 2);
                    float2 v = make_float2(CUDART_INF_F, CUDART_INF_F);  // rows that do not exist: never a seed
                    if (r2 >= 0 && r2 < n_rows) v = reinterpret_cast<const float2*>(shadow)[r2 * (DP / 2) + i % (DP / 2)];
                    reinterpret_cast<float2*>(s_win)[i] = v;
                }
                __syncthreads();
            }
        }
        if (qi >= nq) continue;
        float2 q[DP / 2];
        load_row<DP>(qrows + (q_begin + qi) * DP, q);
        const int64_t row = qhome ? (int64_t)qhome[q_begin + qi] : q_begin + qi;  // seeds: candidates around it
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            blo[2 * h] = fminf(blo[2 * h], q[h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[h].x);
            blo[2 * h + 1] = fminf(blo[2 * h + 1], q[h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[h].y);
        }
        float thr = CUDART_INF_F;
        if constexpr (KC > 0) {
            // the KC smallest distances, ascending, kept by a branch-free min/max insertion network
            float dl[KC];
#pragma unroll
            for (int r = 0; r < KC; ++r) dl[r] = CUDART_INF_F;
            if (staged) {
                const float* mine = s_win + (tid + KNN_SEED_W) * DP;  // this query's own row
#pragma unroll 4
                for (int off = 1; off <= KNN_SEED_W; ++off)
#pragma unroll
                    for (int sgn = -1; sgn <= 1; sgn += 2) {
                        float2 cc[DP / 2];
                        load_row<DP>(mine + sgn * off * DP, cc);
                        const float m = cheb32<DP>(q, cc);  // +inf for the rows that do not exist
#pragma unroll
                        for (int r = KC - 1; r > 0; --r) dl[r] = fminf(dl[r], fmaxf(dl[r - 1], m));
                        dl[0] = fminf(dl[0], m);
                    }
            } else {
#pragma unroll 2
            for (int off = 1; off <= seed_w; ++off)
#pragma unroll
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    const int64_t r2 = sgn > 0 ? row + off - 1 + self_skip : row - off;
                    const bool inside = r2 >= 0 && r2 < n_rows;
                    float2 cc[DP / 2];
                    load_row<DP>(shadow + (inside ? r2 : row) * DP, cc);
                    const float m = inside ? cheb32<DP>(q, cc) : CUDART_INF_F;
#pragma unroll
                    for (int r = KC - 1; r > 0; --r) dl[r] = fminf(dl[r], fmaxf(dl[r - 1], m));
                    dl[0] = fminf(dl[0], m);
                }
            }
#pragma unroll
            for (int r = 0; r < KC; ++r)
                if (r == kseed - 1) thr = dl[r];
        } else {
            const int col = j * FAST_THREADS + tid;
            for (int r = 0; r < kseed; ++r) lst_d[(size_t)r * QB + col] = CUDART_INF_F;
#pragma unroll 1
            for (int off = 1; off <= seed_w; ++off)
#pragma unroll 1
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    const int64_t r2 = sgn > 0 ? row + off - 1 + self_skip : row - off;
                    if (r2 < 0 || r2 >= n_rows) continue;
                    float2 cc[DP / 2];
                    load_row<DP>(shadow + r2 * DP, cc);
                    const float m = cheb32<DP>(q, cc);
                    if (m < thr) {
                        int r = kseed - 1;
                        while (r > 0 && lst_d[(size_t)(r - 1) * QB + col] > m) {
                            lst_d[(size_t)r * QB + col] = lst_d[(size_t)(r - 1) * QB + col];
                            --r;
                        }
                        lst_d[(size_t)r * QB + col] = m;
                        thr = lst_d[(size_t)(kseed - 1) * QB + col];
                    }
                }
        }
        if (kseed < kcap) {
            float mq = 0.0f;
#pragma unroll
            for (int h = 0; h < DP / 2; ++h) mq = fmaxf(mq, fmaxf(fabsf(q[h].x), fabsf(q[h].y)));
            thr = tight_seed(thr, mq);
        }
        // the seed is itself an fp32 distance to a real candidate: one ulp up and that candidate
        // (and everything tied with it) passes the strict test when its tile is scanned
        thr = nextafterf(thr, CUDART_INF_F);  // inf stays inf
        seed_thr[qi] = thr;
        if (thr < CUDART_INF_F) { sum += thr; cnt += 1.0f; }
    }
    block_query_box<DP>(blo, bhi, sh);

    // ---- outlier deferral: one isolated query next to a dense region would blow the block's
    // pruning radius up to its own huge neighbour distance and make all 128*Q queries walk most
    // of the point set.  Queries whose seeded radius is far above the block's mean AND above the
    // block's own extent are taken out (seed = KNN_DEFERRED) and recomputed by a dense launch
    // over the compacted list of such queries.
    sum = block_sum(sum, sh);
    cnt = block_sum(cnt, sh);
    float extent = 0.0f;
#pragma unroll
    for (int d = 0; d < DP; ++d) extent = fmaxf(extent, sh.qhi[d] - sh.qlo[d]);
    const float cut = fmaxf(KNN_DEFER_CUT * (sum / fmaxf(cnt, 1.0f)), 2.0f * extent);
    __syncthreads();  // everyone has read the full box before it is replaced

    // ---- pass 2: box and radius of the queries that stay
#pragma unroll
    for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
    float rmax = -CUDART_INF_F;
#pragma unroll 1
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if (qi >= nq) continue;
        const float thr = seed_thr[qi];
        if (thr < CUDART_INF_F && thr > cut) { seed_thr[qi] = KNN_DEFERRED; continue; }
        float2 q[DP / 2];
        load_row<DP>(qrows + (q_begin + qi) * DP, q);
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            blo[2 * h] = fminf(blo[2 * h], q[h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[h].x);
            blo[2 * h + 1] = fminf(blo[2 * h + 1], q[h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[h].y);
        }
        rmax = fmaxf(rmax, thr);
    }
    block_query_box<DP>(blo, bhi, sh);
    const float R = block_max(rmax, sh);

    // ---- every tile whose box is within R of the query box, in walk order (near first).
    // The lists live in one pool (meta[4..5] = 64-bit bump cursor): a first sweep counts the survivors,
    // the block reserves that many entries, a second sweep writes them.
    const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
    const int64_t own = (qhome ? (int64_t)qhome[mid] : mid) / FAST_TILE;
    auto gap_of = [&](int64_t pos, int64_t& t) -> float {
        t = walk_tile(pos, own, n_tiles);
        float gap[DP];
        box_gaps<DP>(tile_box, t, sh, gap);
        float dist = gap[0];
#pragma unroll
        for (int d = 1; d < DP; ++d) dist = fmaxf(dist, gap[d]);
        return dist;
    };
    // (up to PLAN_MASK_WORDS * 32 tiles the first sweep leaves a survivor bitmask in shared memory --
    // word i = walk positions [32 i, 32 i + 32) -- and the second one only visits the set bits)
    __shared__ unsigned s_mask[PLAN_MASK_WORDS];
    const bool masked = n_tiles <= (int64_t)PLAN_MASK_WORDS * 32;
    const int n_words = (int)((n_tiles + 31) / 32);
    int mine = 0;
    if (masked) {
        for (int wi = warp; wi < n_words; wi += FAST_THREADS / 32) {
            const int64_t pos = (int64_t)wi * 32 + lane;
            int64_t t;
            const bool keep = pos < n_tiles && !(gap_of(pos, t) > R);
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) {
                s_mask[wi] = ballot;
                mine += __popc(ballot);
            }
        }
    } else {
        for (int64_t pos = tid; pos < n_tiles; pos += FAST_THREADS) {
            int64_t t;
            mine += !(gap_of(pos, t) > R);
        }
    }
    const int total = (int)(block_sum((float)mine, sh) + 0.5f);  // exact: < 2^24 tiles (and publishes s_mask)
    // A block that keeps most of the tiles gets no list at all: plan_first = -1 means "every
    // tile, in walk order" (the warp-level sub-tile culling of the filter kernel still applies).
    // That also bounds the pool: explicit lists are at most half the tiles long.
    if (tid == 0) {
        int base = -1;
        if (total > 0 && 2 * (int64_t)total <= n_tiles) {
            // 64-bit bump cursor at meta[4..5] (zeroed before this kernel)
            const unsigned long long at = atomicAdd(reinterpret_cast<unsigned long long*>(meta + 4),
                                                    (unsigned long long)total);
            base = at + (unsigned long long)total <= (unsigned long long)plan_capacity ? (int)at : -1;  // else: pool exhausted
        }
        s_base = base;
    }
    __syncthreads();
    const int base = s_base;
    if (total > 0 && base < 0) {
        if (tid == 0) { plan_count[blockIdx.x] = (int32_t)n_tiles; plan_first[blockIdx.x] = -1; }
        return;
    }
    int2* out = plan + base;
    if (masked) {
        // thread t owns the words [t W, (t + 1) W): block-wide exclusive scan of their bit counts,
        // then one (tile, gap) entry per set bit, in walk order
        const int W = (n_words + FAST_THREADS - 1) / FAST_THREADS;
        const int w0 = min(tid * W, n_words), w1 = min(w0 + W, n_words);
        int cnt = 0;
        for (int wi = w0; wi < w1; ++wi) cnt += __popc(s_mask[wi]);
        int incl = cnt;
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) sh.warp_count[warp] = incl;
        __syncthreads();
        int offset = incl - cnt;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w)
            if (w < warp) offset += sh.warp_count[w];
        for (int wi = w0; wi < w1; ++wi) {
            unsigned bits = s_mask[wi];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                int64_t t;
                const float dist = gap_of((int64_t)wi * 32 + b, t);
                out[offset++] = make_int2((int)t, __float_as_int(dist));
            }
        }
        if (tid == 0) { plan_count[blockIdx.x] = total; plan_first[blockIdx.x] = total > 0 ? base : 0; }
        return;
    }
    int done = 0;
    for (int64_t pos = 0; pos < n_tiles; pos += FAST_THREADS) {
        bool keep = false;
        int64_t t = 0;
        float dist = 0.0f;
        if (pos + tid < n_tiles) {
            dist = gap_of(pos + tid, t);
            keep = !(dist > R);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) sh.warp_count[warp] = __popc(ballot);
        __syncthreads();
        int offset = done, round_total = 0;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w) {
            if (w < warp) offset += sh.warp_count[w];
            round_total += sh.warp_count[w];
        }
        if (keep) out[offset + __popc(ballot & ((1u << lane) - 1u))] = make_int2((int)t, __float_as_int(dist));
        done += round_total;
    }
    if (tid == 0) { plan_count[blockIdx.x] = total; plan_first[blockIdx.x] = total > 0 ? base : 0; }
}

// part_d32 / part_idx: [split][rank][nq]
template <int DP, int Q, int STAGES>
static __global__ void __launch_bounds__(FAST_THREADS)
knn_f32_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box, int64_t n_tiles,
               int64_t q_begin, int64_t nq, const int32_t* __restrict__ q_index, int kcap,
               const double* __restrict__ maxabs, int dim, float* __restrict__ part_d32,
               int32_t* __restrict__ part_idx, unsigned long long* __restrict__ tile_counter) {
    constexpr int QB = FAST_THREADS * Q;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // STAGES * TILE_FLOATS
    float* lst_d = tiles + STAGES * TILE_FLOATS;                             // kcap * QB
    int32_t* lst_i = reinterpret_cast<int32_t*>(lst_d + (size_t)kcap * QB);  // kcap * QB
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ TileQueueShared sh;

    const int tid = threadIdx.x;
    const int64_t qb0 = (int64_t)blockIdx.x * QB;  // first query of the block (relative to q_begin)
    (void)maxabs; (void)dim;

    // queries: slot j of thread t is query qb0 + j*THREADS + t
    float2 q[Q][DP / 2];
    float thr[Q];     // kcap-th smallest so far
    bool live[Q];
    int64_t qrow[Q];  // row of the query in the prepared order
    float blo[DP], bhi[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        live[j] = qi < nq;
        const int64_t row = q_index ? (int64_t)q_index[live[j] ? qi : 0] : q_begin + (live[j] ? qi : 0);
        qrow[j] = row;
        load_row<DP>(shadow + row * DP, q[j]);
        thr[j] = live[j] ? CUDART_INF_F : -1.0f;  // dead slots never insert
        if (live[j]) {
#pragma unroll
            for (int h = 0; h < DP / 2; ++h) {
                blo[2 * h] = fminf(blo[2 * h], q[j][h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[j][h].x);
                blo[2 * h + 1] = fminf(blo[2 * h + 1], q[j][h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[j][h].y);
            }
        }
    }
    auto reset_lists = [&]() {
        for (int r = 0; r < kcap; ++r)
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                lst_d[(size_t)r * QB + j * FAST_THREADS + tid] = CUDART_INF_F;
                lst_i[(size_t)r * QB + j * FAST_THREADS + tid] = -1;
            }
    };
    // sorted insertion of (m, idx) into slot j's list; returns the new kcap-th value
    auto insert = [&](int j, float m, int32_t idx) -> float {
        const int col = j * FAST_THREADS + tid;
        int r = kcap - 1;
        while (r > 0 && lst_d[(size_t)(r - 1) * QB + col] > m) {
            lst_d[(size_t)r * QB + col] = lst_d[(size_t)(r - 1) * QB + col];
            lst_i[(size_t)r * QB + col] = lst_i[(size_t)(r - 1) * QB + col];
            --r;
        }
        lst_d[(size_t)r * QB + col] = m;
        lst_i[(size_t)r * QB + col] = idx;
        return lst_d[(size_t)(kcap - 1) * QB + col];
    };
    reset_lists();

    // ---- seed the thresholds: the kcap-th smallest fp32 distance among the 2*KNN_SEED_W rows
    // around each query in the prepared order is an upper bound of its final kcap-th distance.
    // Starting from it (one ulp up, so that those very rows re-enter the list when their tile is
    // scanned) removes the flood of insertions a threshold of +inf causes and gives the tile
    // queue a finite pruning radius from its first round.
    {
        const int64_t n_rows = n_tiles * FAST_TILE;
#pragma unroll 1
        for (int off = 1; off <= KNN_SEED_W; ++off) {
#pragma unroll 1
            for (int sgn = -1; sgn <= 1; sgn += 2) {
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    const int64_t row = qrow[j] + sgn * off;
                    if (live[j] && row >= 0 && row < n_rows) {
                        float2 cc[DP / 2];
                        load_row<DP>(shadow + row * DP, cc);
                        const float m = cheb32<DP>(q[j], cc);
                        if (m < thr[j]) thr[j] = insert(j, m, 0);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j]) thr[j] = nextafterf(thr[j], CUDART_INF_F);  // inf stays inf
        reset_lists();
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    block_query_box<DP>(blo, bhi, sh);  // (barriers inside: also publishes the mbarrier init)

    float R = CUDART_INF_F;  // CTA radius: max over its queries' thresholds (unused by the dense walk)
    auto test = [&](int64_t t) -> bool {
        float gap[DP];
        box_gaps<DP>(tile_box, t, sh, gap);
        float dist = gap[0];
#pragma unroll
        for (int d = 1; d < DP; ++d) dist = fmaxf(dist, gap[d]);
        return !(dist > R);
    };
    auto refresh = [&]() {
        float v = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j]) v = fmaxf(v, thr[j]);
        R = block_max(v, sh);
    };
    auto compute = [&](const float* tile, int64_t t) {
        const int64_t tile_base = t * FAST_TILE;
        for (int b = 0; b < FAST_TILE; b += KNN_BATCH) {
            float mrun[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) mrun[j] = CUDART_INF_F;
#pragma unroll
            for (int c = 0; c < KNN_BATCH; c += 2) {
                float2 ca[DP / 2], cb[DP / 2];
                load_row<DP>(tile + (b + c) * DP, ca);
                load_row<DP>(tile + (b + c + 1) * DP, cb);
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    const float ma = cheb32<DP>(q[j], ca);
                    const float mb = cheb32<DP>(q[j], cb);
                    mrun[j] = fminf(fminf(mrun[j], ma), mb);
                }
            }
            bool hit = false;
#pragma unroll
            for (int j = 0; j < Q; ++j) hit = hit || (mrun[j] < thr[j]);
            if (hit) {
                // rare: walk the batch again for the slots that saw a closer candidate, and insert
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    if (mrun[j] < thr[j]) {
#pragma unroll 1
                        for (int c = 0; c < KNN_BATCH; ++c) {
                            float2 cc[DP / 2];
                            load_row<DP>(tile + (b + c) * DP, cc);
                            const float m = cheb32<DP>(q[j], cc);
                            if (m < thr[j]) thr[j] = fminf(thr[j], insert(j, m, (int32_t)(tile_base + b + c)));
                        }
                    }
                }
            }
        }
    };

    // the block's own tile: the one holding its middle query
    const int64_t mid = q_index ? 0 : q_begin + min(qb0 + QB / 2, nq - 1);
    const int tiles_done = run_tile_queue<DP, STAGES>(shadow, tiles, full_bar, sh, mid / FAST_TILE, n_tiles,
                                                      (int)blockIdx.y, (int)gridDim.y, false, test, compute, refresh);
    if (tid == 0 && tile_counter) atomicAdd(tile_counter, (unsigned long long)tiles_done * (FAST_TILE * QB));

    // emit the lists (ascending), one per split
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if (qi < nq) {
            const int col = j * FAST_THREADS + tid;
            for (int r = 0; r < kcap; ++r) {
                const int64_t o = ((int64_t)blockIdx.y * kcap + r) * nq + qi;
                part_d32[o] = lst_d[(size_t)r * QB + col];
                part_idx[o] = lst_i[(size_t)r * QB + col];
            }
        }
    }
}

// ---- work units of the windowed engine.  The planned tile lists are very uneven (a block of
// queries in a sparse region keeps hundreds of tiles, a typical one ten), so a launch with one
// CTA per block ends in a long tail of a few heavy CTAs.  The lists are therefore cut into
// units of at most `chunk` tiles; every unit is an independent (block, tile-range) job that
// starts from the block's seeded thresholds and emits its own partial lists, which the refine
// kernel merges exactly like the split lists of the dense launch.
//   meta[0] = number of units, meta[1] = chunk, meta[2] = work-queue cursor (zeroed here),
//   meta[4..5] = tile-list pool cursor, 64 bit (zeroed before the planning pass)
static __global__ void __launch_bounds__(1024)
knn_unit_table_kernel(const int32_t* __restrict__ plan_count, int64_t n_blocks, int64_t max_units, int min_tiles,
                      int32_t* __restrict__ unit_first, int32_t* __restrict__ unit_block,
                      int32_t* __restrict__ meta) {
    __shared__ long long s_red[32];
    __shared__ int s_chunk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (n_blocks + 1023) / 1024;
    const int64_t b0 = (int64_t)tid * per, b1 = b0 + per < n_blocks ? b0 + per : n_blocks;
    long long tiles = 0;
    for (int64_t b = b0; b < b1; ++b) tiles += plan_count[b] > 0 ? plan_count[b] : 0;
    for (int off = 16; off > 0; off >>= 1) tiles += __shfl_xor_sync(0xffffffffu, tiles, off);
    if (lane == 0) s_red[warp] = tiles;
    __syncthreads();
    if (tid == 0) {
        long long total = 0;
        for (int w = 0; w < 32; ++w) total += s_red[w];
        long long chunk = (total + KNN_UNIT_FACTOR * n_blocks - 1) / (KNN_UNIT_FACTOR * n_blocks);
        if (chunk < min_tiles) chunk = min_tiles;
        s_chunk = (int)chunk;
    }
    __syncthreads();
    const int chunk = s_chunk;
    // Units of the blocks that need SEVERAL come first in the queue: their partial lists are merged by
    // whoever finishes the block's last unit, and that merge must not be the tail of the launch
    // (longest work first).  unit_first[b] = first unit of block b, unit_first[n_blocks + 1 + b] = how many.
    int32_t* unit_count = unit_first + (n_blocks + 1);
    int mine_multi = 0, mine_single = 0;
    for (int64_t b = b0; b < b1; ++b) {
        const int c = plan_count[b];
        const int nu = c > 0 ? (c + chunk - 1) / chunk : 0;
        if (nu > 1) mine_multi += nu; else mine_single += nu;
    }
    // one scan over (multi, single) packed into 64 bits (each < 2^31)
    __shared__ unsigned long long s_scan2[1024];
    unsigned long long mine2 = ((unsigned long long)mine_multi << 32) | (unsigned long long)mine_single;
    s_scan2[tid] = mine2;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // inclusive scan
        const unsigned long long v = tid >= off ? s_scan2[tid - off] : 0ull;
        __syncthreads();
        s_scan2[tid] += v;
        __syncthreads();
    }
    const unsigned long long incl = s_scan2[tid], total2 = s_scan2[1023];
    const int total_multi = (int)(total2 >> 32), total_single = (int)(total2 & 0xffffffffull);
    int um = (int)(incl >> 32) - mine_multi;                                   // next multi-unit slot
    int us = total_multi + (int)(incl & 0xffffffffull) - mine_single;          // next single-unit slot
    for (int64_t b = b0; b < b1; ++b) {
        const int c = plan_count[b];
        const int nu = c > 0 ? (c + chunk - 1) / chunk : 0;
        int& u = nu > 1 ? um : us;
        unit_first[b] = u;
        unit_count[b] = nu;
        for (int i = 0; i < nu; ++i)
            if (u + i < max_units) unit_block[u + i] = (int32_t)b;
        u += nu;
    }
    if (tid == 1023) {
        const int total = total_multi + total_single;
        unit_first[n_blocks] = total;
        meta[0] = total < max_units ? total : (int32_t)max_units;  // (bound holds by construction)
        meta[1] = chunk;
        meta[2] = 0;
    }
}

// (mask << 1) | sign(t): one funnel shift appends a candidate's "closer than the threshold" bit
__device__ __forceinline__ uint32_t push_sign(uint32_t mask, float t) {
    uint32_t r;
    asm("shf.l.clamp.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(__float_as_uint(t)), "r"(mask));
    return r;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int UNITS_THREADS = FAST_THREADS + 32;  // 4 filter warps + 1 copy warp

static __device__ void refine_single_list(const float* col_d, const int32_t* col_i, int stride, int kcap, float seed,
                                   int64_t row_q, int64_t qi, const RefineArgs& ra);
static __device__ void refine_merge_lists(const float* part_d32, const int32_t* part_idx, int kcap, int s_begin,
                                          int s_end, int64_t stride, int64_t off, bool has_seed, float seed,
                                          int64_t row_q, int64_t qi, const RefineArgs& ra);
__device__ __forceinline__ void filter_warps_barrier() {  // the four filter warps only (the copy warp is elsewhere)
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Persistent filter kernel of the windowed engine: CTAs pull work units from a global cursor.
//
// Roles.  Warps 0..3 filter; warp 4 only feeds the shared-memory ring (one lane issues the
// bulk copies).  Ring slots are handed over with a full / empty mbarrier pair, so the filter
// warps never meet at a block barrier inside a unit and may run up to STAGES tiles apart --
// they do very different amounts of work per tile, see below.
//
// Pruning goes one level finer than the planned tile list: every filter warp owns 32*Q
// CONSECUTIVE queries of the block (a compact blob in the prepared order) and tests the
// bounding box of each 32-row sub-tile (staged beside the tile by a second bulk copy) against
// its own query box and its own largest threshold -- a warp-uniform branch.  Surviving
// sub-tiles are scanned with per-candidate hit bits instead of the dense kernel's "running
// minimum + re-walk" (which pays off only when hits are rare).  Per (query, candidate):
//     DP/2 FADD2 + ~DP/2 FMNMX   the distance,
//     1 FADD (m - thr)           its sign is the hit bit,
//     1 SHF                      funnel-shifts the bit into the slot's 32-bit mask;
// then the lanes pop their hit bits in lock step.  The per-query lists (shared memory) are
// UNSORTED: a hit overwrites the current maximum and one sweep over the kcap entries finds the
// new maximum -- constant cost, whereas a sorted insert shifts most of the list because new
// hits are usually closer than what the list holds.  Lists are sorted once, when emitted.
// part_d32 / part_idx: [unit][rank][QB], QB index = position of the query in its block.
// Register cap = what the filter loop alone needs (so many CTAs fit beside its shared memory); the
// fused refine epilogue (a noinline call, once per query) spills instead of costing residency.
constexpr int units_min_blocks(int dp) { return dp <= 2 ? 9 : dp <= 6 ? 6 : 5; }

template <int DP, int Q, int STAGES, bool HITMASK>
static __global__ void __launch_bounds__(UNITS_THREADS, units_min_blocks(DP))
knn_units_kernel(const float* __restrict__ shadow, const float* __restrict__ sub_box, int64_t n_tiles,
                 int64_t q_begin, int64_t nq, int kcap, const float* __restrict__ qshadow,
                 const int32_t* __restrict__ qhome, const float* __restrict__ seed_thr,
                 const int2* __restrict__ plan, const int32_t* __restrict__ plan_count,
                 const int32_t* __restrict__ plan_first, const int32_t* __restrict__ unit_first,
                 const int32_t* __restrict__ unit_block, int32_t* __restrict__ meta, float* __restrict__ part_d32,
                 int32_t* __restrict__ part_idx, unsigned long long* __restrict__ pair_counter,
                 int32_t* __restrict__ unit_done, const __grid_constant__ RefineArgs ra) {
    static_assert(FAST_THREADS == 128, "filter_warps_barrier() counts 128 threads");
    constexpr int QB = FAST_THREADS * Q;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    constexpr int BOX_FLOATS = FAST_NSUB * DP * 2;
    constexpr int SLOT_FLOATS = TILE_FLOATS + BOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // STAGES * SLOT_FLOATS
    float* lst_d = tiles + STAGES * SLOT_FLOATS;                             // kcap * QB
    int32_t* lst_i = reinterpret_cast<int32_t*>(lst_d + (size_t)kcap * QB);  // kcap * QB
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int s_unit;
    __shared__ int s_last;                // this CTA finished the last open unit of its block
    __shared__ int s_tile[STAGES];        // tile staged in each ring slot (-1: end of the unit)
    __shared__ float s_wthr[FAST_THREADS / 32];  // largest threshold of each filter warp

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == FAST_THREADS / 32;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);                   // the copy lane's arrive (+ transaction bytes)
            mbar_init(&empty_bar[s], FAST_THREADS / 32);  // one arrive per filter warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_units = meta[0], chunk = meta[1];
    const int32_t* unit_count = unit_first + ((nq + QB - 1) / QB + 1);  // (see knn_unit_table_kernel)
    // ring state (kernel lifetime, advanced identically by both roles): next slot and the
    // phase parity to wait for on it.  A fresh mbarrier passes a wait on parity 1, which is
    // how the copy lane finds every slot free at the start.
    int slot = 0;
    uint32_t phase_bits = producer ? 0xffffffffu : 0u;
    auto advance = [&]() {
        phase_bits ^= 1u << slot;
        slot = slot + 1 == STAGES ? 0 : slot + 1;
    };
    unsigned long long subtiles_total = 0;  // (warp, sub-tile) scans of this warp

    for (;;) {
        if (tid == 0) s_unit = atomicAdd(&meta[2], 1);
        __syncthreads();  // (A) the unit is published; the previous one is finished by everybody
        const int u = s_unit;
        if (u >= n_units) break;
        const int64_t qblock = unit_block[u];
        const int first = (u - unit_first[qblock]) * chunk;
        const int count = plan_count[qblock];
        const int last = first + chunk < count ? first + chunk : count;
        const int64_t qb0 = qblock * QB;
        const int pfirst = plan_first[qblock];
        // a block whose whole tile list is this one unit is refined right here, from shared memory;
        // the units of a longer list emit partial lists, and whoever finishes the LAST of them merges
        const int ufirst = unit_first[qblock];
        const int nunits = unit_count[qblock];
        const bool refine_here = ra.fused && nunits == 1;

        // ---------------- filter warps: queries, thresholds, warp box, empty lists
        // slot j of this thread = query qb0 + warp*32*Q + j*32 + lane
        float2 q[Q][DP / 2];
        float thr[Q];   // insertion threshold: min(seed, current list maximum)
        float seed[Q];
        int pmax[Q];    // position of the list's maximum (the entry a hit replaces)
        int nfill[Q];   // entries written so far (kcap: the list is full)
        bool alive[Q];  // the slot holds a query of this launch
        float wlo[DP], whi[DP];
        float wthr = -1.0f;
        auto warp_refresh = [&]() {
            float v = -1.0f;
#pragma unroll
            for (int j = 0; j < Q; ++j) v = fmaxf(v, thr[j]);
            for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
            wthr = v;
            if (lane == 0) *reinterpret_cast<volatile float*>(&s_wthr[warp]) = v;
        };
        if (!producer) {
#pragma unroll
            for (int d = 0; d < DP; ++d) { wlo[d] = CUDART_INF_F; whi[d] = -CUDART_INF_F; }
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const int64_t qi = qb0 + warp * (32 * Q) + j * 32 + lane;
                const bool in = qi < nq;
                load_row<DP>((qshadow ? qshadow : shadow) + (q_begin + (in ? qi : 0)) * DP, q[j]);
                const float s = in ? seed_thr[qi] : -1.0f;
                seed[j] = s < 0.0f ? -1.0f : s;  // deferred / dead slots never insert
                thr[j] = seed[j];
                pmax[j] = 0;
                nfill[j] = 0;
                alive[j] = in;
                if (thr[j] >= 0.0f) {
#pragma unroll
                    for (int h = 0; h < DP / 2; ++h) {
                        wlo[2 * h] = fminf(wlo[2 * h], q[j][h].x);         whi[2 * h] = fmaxf(whi[2 * h], q[j][h].x);
                        wlo[2 * h + 1] = fminf(wlo[2 * h + 1], q[j][h].y); whi[2 * h + 1] = fmaxf(whi[2 * h + 1], q[j][h].y);
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < DP; ++d)
                for (int off = 16; off > 0; off >>= 1) {
                    wlo[d] = fminf(wlo[d], __shfl_xor_sync(0xffffffffu, wlo[d], off));
                    whi[d] = fmaxf(whi[d], __shfl_xor_sync(0xffffffffu, whi[d], off));
                }
            for (int r = 0; r < kcap; ++r)
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    lst_d[(size_t)r * QB + j * FAST_THREADS + tid] = CUDART_INF_F;
                    lst_i[(size_t)r * QB + j * FAST_THREADS + tid] = -1;
                }
            warp_refresh();
        }
        __syncthreads();  // (B) s_wthr is set; s_unit has been read by everybody

        if (producer) {
            // ---------------- copy warp: stream the unit's tiles, then an end marker.  All 32
            // lanes walk the list together (no divergence, no spinning lanes); lane 0 issues.
            const int2* lst = plan + (pfirst < 0 ? 0 : pfirst);
            const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
            const int64_t own = (qhome ? (int64_t)qhome[mid] : mid) / FAST_TILE;  // as in the planning pass
            for (int next = first; next <= last; ++next) {
                int tile = -1;
                if (next < last) {
                    if (pfirst < 0) {  // implicit list: every tile, in walk order
                        tile = (int)walk_tile(next, own, n_tiles);
                    } else {
                        const int2 e = lst[next];
                        float R = -1.0f;  // the block's current radius (thresholds only shrink)
#pragma unroll
                        for (int w = 0; w < FAST_THREADS / 32; ++w)
                            R = fmaxf(R, *reinterpret_cast<volatile float*>(&s_wthr[w]));
                        if (!(__int_as_float(e.y) < R)) continue;
                        tile = e.x;
                    }
                }
                mbar_wait_relaxed(&empty_bar[slot], (phase_bits >> slot) & 1u);
                if (lane == 0) {
                    *reinterpret_cast<volatile int*>(&s_tile[slot]) = tile;  // published by the mbarrier below
                    if (tile >= 0) {
                        float* dst = tiles + (size_t)slot * SLOT_FLOATS;
                        mbar_expect_tx(&full_bar[slot], SLOT_FLOATS * 4);
                        bulk_g2s(dst, shadow + (int64_t)tile * TILE_FLOATS, TILE_FLOATS * 4, &full_bar[slot]);
                        bulk_g2s(dst + TILE_FLOATS, sub_box + (int64_t)tile * BOX_FLOATS, BOX_FLOATS * 4, &full_bar[slot]);
                    } else {
                        mbar_arrive(&full_bar[slot]);
                    }
                }
                __syncwarp();
                advance();
            }
            continue;
        }

        // ---------------- filter warps
        auto insert = [&](int j, float m, int32_t idx) -> bool {
            const int col = j * FAST_THREADS + tid;
            lst_d[(size_t)pmax[j] * QB + col] = m;
            lst_i[(size_t)pmax[j] * QB + col] = idx;
            // while the list still has free (+inf) slots the entry to replace is simply the next free
            // one and the threshold stays at the seed: no sweep (it would find just that)
            if (nfill[j] + 1 < kcap) {
                pmax[j] = ++nfill[j];
                return false;
            }
            nfill[j] = kcap;
            float mx = -1.0f;
            int p = 0;
#pragma unroll 3
            for (int r = 0; r < kcap; ++r) {
                const float v = lst_d[(size_t)r * QB + col];
                if (v > mx) { mx = v; p = r; }
            }
            pmax[j] = p;
            thr[j] = fminf(seed[j], mx);
            return true;  // the threshold may have shrunk
        };
        for (;;) {
            mbar_wait(&full_bar[slot], (phase_bits >> slot) & 1u);
            const int t = *reinterpret_cast<volatile int*>(&s_tile[slot]);
            if (t >= 0) {
                const float* tile = tiles + (size_t)slot * SLOT_FLOATS;
                const float2* boxes = reinterpret_cast<const float2*>(tile + TILE_FLOATS);
                const int64_t tile_base = (int64_t)t * FAST_TILE;
                // lane s (mod 16) tests sub-tile s: one round of box arithmetic per tile instead of
                // sixteen warp-uniform ones; a survivor's gap is re-read by shuffle and tested again
                // against the (shrinking) warp threshold when its turn comes
                static_assert(FAST_NSUB == 16, "lane-parallel sub-tile test assumes 16 sub-tiles per tile");
                float my_gap = 0.0f;
                {
                    const int sbl = lane & (FAST_NSUB - 1);
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        const float2 bx = boxes[sbl * DP + d];
                        my_gap = fmaxf(my_gap, fmaxf(bx.x - whi[d], wlo[d] - bx.y));
                    }
                }
                uint32_t live = __ballot_sync(0xffffffffu, my_gap < wthr) & ((1u << FAST_NSUB) - 1u);
#pragma unroll 1
                while (live) {
                    const int sb = __ffs(live) - 1;
                    live &= live - 1u;
                    const float gap = __shfl_sync(0xffffffffu, my_gap, sb);
                    // insertion needs m < thr and the box gap is a lower bound of every computed m
                    if (!(gap < wthr)) continue;  // warp-uniform
                    ++subtiles_total;
                    const float* rows = tile + sb * FAST_SUB * DP;
                    bool changed = false;
                    if constexpr (HITMASK) {
                        // few dimensions: the culling is sharp, most scanned sub-tiles hold hits
                        uint32_t mask[Q];
#pragma unroll
                        for (int j = 0; j < Q; ++j) mask[j] = 0u;
#pragma unroll 4
                        for (int c = 0; c < FAST_SUB; c += 2) {
                            float2 ca[DP / 2], cb[DP / 2];
                            load_two_rows<DP>(rows + c * DP, ca, cb);
#pragma unroll
                            for (int j = 0; j < Q; ++j) {
                                mask[j] = push_sign(mask[j], cheb32<DP>(q[j], ca) - thr[j]);
                                mask[j] = push_sign(mask[j], cheb32<DP>(q[j], cb) - thr[j]);
                            }
                        }
                        // candidate c of the sub-tile sits at bit 31 - c
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            uint32_t mk = mask[j];
                            while (mk) {
                                const int c = __clz(mk);
                                mk &= ~(0x80000000u >> c);
                                float2 cc[DP / 2];
                                load_row<DP>(rows + c * DP, cc);
                                const float m = cheb32<DP>(q[j], cc);
                                if (m < thr[j]) changed |= insert(j, m, (int32_t)(tile_base + sb * FAST_SUB + c));
                            }
                        }
                    } else {
                        // many dimensions: boxes overlap, hits are rare among the scanned pairs --
                        // keep only the running minimum (half a FMNMX3 per pair) and walk the
                        // sub-tile again for the slots that saw something closer
                        float mrun[Q];
#pragma unroll
                        for (int j = 0; j < Q; ++j) mrun[j] = CUDART_INF_F;
#pragma unroll 4
                        for (int c = 0; c < FAST_SUB; c += 2) {
                            float2 ca[DP / 2], cb[DP / 2];
                            load_two_rows<DP>(rows + c * DP, ca, cb);
#pragma unroll
                            for (int j = 0; j < Q; ++j)
                                mrun[j] = fminf(fminf(mrun[j], cheb32<DP>(q[j], ca)), cheb32<DP>(q[j], cb));
                        }
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            if (mrun[j] < thr[j]) {
#pragma unroll 1
                                for (int c = 0; c < FAST_SUB; ++c) {
                                    float2 cc[DP / 2];
                                    load_row<DP>(rows + c * DP, cc);
                                    const float m = cheb32<DP>(q[j], cc);
                                    if (m < thr[j]) changed |= insert(j, m, (int32_t)(tile_base + sb * FAST_SUB + c));
                                }
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, changed)) warp_refresh();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);  // this warp is done with the slot
            advance();
            if (t < 0) break;
        }

        // ---------------- sort each list (ascending); refine it here or emit it
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const int col = j * FAST_THREADS + tid;
            const int filled = nfill[j];  // the rest of the column is +inf / -1
            for (int a = 1; a < filled; ++a) {
                const float d = lst_d[(size_t)a * QB + col];
                const int32_t i = lst_i[(size_t)a * QB + col];
                int b = a - 1;
                while (b >= 0 && lst_d[(size_t)b * QB + col] > d) {
                    lst_d[(size_t)(b + 1) * QB + col] = lst_d[(size_t)b * QB + col];
                    lst_i[(size_t)(b + 1) * QB + col] = lst_i[(size_t)b * QB + col];
                    --b;
                }
                lst_d[(size_t)(b + 1) * QB + col] = d;
                lst_i[(size_t)(b + 1) * QB + col] = i;
            }
            const int lq = warp * (32 * Q) + j * 32 + lane;
            if (refine_here) {
                if (alive[j])
                    refine_single_list(lst_d + col, lst_i + col, QB, kcap, seed[j], q_begin + qb0 + lq, qb0 + lq, ra);
                continue;
            }
            for (int r = 0; r < kcap; ++r) {
                const int64_t o = ((int64_t)u * kcap + r) * QB + lq;
                part_d32[o] = lst_d[(size_t)r * QB + col];
                part_idx[o] = lst_i[(size_t)r * QB + col];
            }
        }
        if (ra.fused && !refine_here) {
            // last-unit-done merge: the lists above are published (fence), the block's counter is
            // bumped once per unit, and the CTA that brings it to nunits reads everybody's lists
            __threadfence();
            filter_warps_barrier();
            if (tid == 0) s_last = atomicAdd(&unit_done[4 * qblock], 1) == nunits - 1;
            filter_warps_barrier();
            if (s_last) {
                __threadfence();
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    const int lq = warp * (32 * Q) + j * 32 + lane;
                    if (alive[j])
                        refine_merge_lists(part_d32, part_idx, kcap, ufirst, ufirst + nunits, QB, lq, true,
                                           seed[j] < 0.0f ? -1.0f : seed_thr[qb0 + lq], q_begin + qb0 + lq,
                                           qb0 + lq, ra);
                }
            }
        }
    }
    if (!producer && lane == 0 && pair_counter && subtiles_total)
        atomicAdd(pair_counter, subtiles_total * (unsigned long long)(FAST_SUB * 32 * Q));
}

// Exact refine: see the file header.  flags[qi] = 1 marks a query for the fp64 fallback.
constexpr int REFINE_CAP = 96;  // union entries examined per query
constexpr int REFINE_BATCH = 8;  // list heads / tails fetched together

// loads of the partial lists: L2 only.  Inside knn_units_kernel they were written by OTHER CTAs of
// the same launch (last-unit-done merge), which the per-SM L1 knows nothing about.
#ifdef __CUDA_ARCH__
#define KSG_LD_PART(ptr) __ldcg(ptr)
#else
#define KSG_LD_PART(ptr) (*(ptr))
#endif

// Merge the ascending lists [s_begin, s_end) of one query (entry (s, r) at ((s * kcap + r) * stride +
// off)) and decide its exact radius; see the file header.  has_seed: `seed` is the threshold the
// windowed filter ran under (< 0: the planning pass set the query aside).
static __device__ __noinline__ void refine_merge_lists(const float* part_d32, const int32_t* part_idx, int kcap,
                                                       int s_begin, int s_end, int64_t stride, int64_t off,
                                                       bool has_seed, float seed, int64_t row_q, int64_t qi,
                                                       const RefineArgs& ra) {
    const int kp = ra.kp, dim = ra.dim;
    const double* sorted = ra.sorted;
    const int64_t ld = ra.ld;
    if (has_seed && seed < 0.0f) {  // taken out by the planning pass: needs the dense launch
        ra.flags[qi] = 2;
        atomicAdd(ra.n_flagged, 1);
        ra.eps_out[qi] = CUDART_NAN;
        return;
    }
    if (ra.maxabs && !fp32_range_ok(ra.maxabs, dim)) {  // shadow rows may have overflowed: float64 kernels
        ra.flags[qi] = 1;
        atomicAdd(ra.n_flagged, 1);
        ra.eps_out[qi] = CUDART_NAN;
        return;
    }
    // the query's own magnitude in the centred frame bounds the fp32 error of all its near pairs
    double qc[KSG_MAX_DIM];
    double mq = 0.0;
    for (int d = 0; d < dim; ++d) {
        qc[d] = ra.qsorted ? ra.qsorted[(int64_t)d * ra.qld + row_q] : sorted[(int64_t)d * ld + row_q];
        mq = fmax(mq, fabs(qc[d] - ra.center[d]));
    }
    mq = mq * (1.0 + 1e-6) + 1e-300;

    // k'-th smallest fp32 distance over the union of the split lists
    // (a block in a sparse region can own a hundred work units, most of whose lists hold nothing
    // below the running k'-th value: the list heads are fetched REFINE_BATCH at a time, so those
    // lists cost one overlapped load each instead of one dependent L2 round trip each)
    float best[KNN_MAX_KCAP];
    int have = 0;
    for (int s0 = s_begin; s0 < s_end; s0 += REFINE_BATCH) {
        float head[REFINE_BATCH];
#pragma unroll
        for (int u = 0; u < REFINE_BATCH; ++u)
            head[u] = s0 + u < s_end ? KSG_LD_PART(&part_d32[((int64_t)(s0 + u) * kcap) * stride + off]) : CUDART_INF_F;
#pragma unroll
        for (int u = 0; u < REFINE_BATCH; ++u) {
            const int s = s0 + u;
            if (s >= s_end) break;
            if (have == kp && !(head[u] < best[kp - 1])) continue;  // ascending list: nothing in it counts
            for (int r = 0; r < kcap; ++r) {
                const float v = r == 0 ? head[u] : KSG_LD_PART(&part_d32[((int64_t)s * kcap + r) * stride + off]);
                if (have == kp && !(v < best[kp - 1])) break;
                int pos = have < kp ? have : kp - 1;
                while (pos > 0 && best[pos - 1] > v) { best[pos] = best[pos - 1]; --pos; }
                best[pos] = v;
                if (have < kp) ++have;
            }
        }
    }
    const double r32 = (have == kp) ? (double)best[kp - 1] : CUDART_INF;
    if (!(r32 < CUDART_INF)) {  // fewer than k' points exist (k+1 > n): cKDTree pads with inf
        ra.eps_out[qi] = CUDART_INF;
        ra.flags[qi] = ra.idx_out ? 1 : 0;  // neighbour identities: leave it to the dense float64 kernel
        if (ra.idx_out) atomicAdd(ra.n_flagged, 1);
        return;
    }
    if (ra.exact32 && *ra.exact32 && !ra.idx_out) {  // exact fp32 copy: see refine_single_list
        ra.flags[qi] = 0;
        ra.eps_out[qi] = r32;
        return;
    }
    const double delta = band_query_f64(mq, 2.0 * r32);
    const double T = r32 + 2.0 * delta;

    // a list that is not full holds everything below the seeded threshold -- and nothing above it
    bool overflow = has_seed && !(T < (double)seed);
    double exact[KNN_MAX_KCAP];
    int32_t exact_j[KNN_MAX_KCAP], exact_o[KNN_MAX_KCAP];  // prepared position / original index (idx_out only)
    int got = 0, examined = 0;
    for (int s0 = s_begin; s0 < s_end && !overflow; s0 += REFINE_BATCH) {
      float head[REFINE_BATCH], tails[REFINE_BATCH];
#pragma unroll
      for (int u = 0; u < REFINE_BATCH; ++u) {
          const bool in = s0 + u < s_end;
          head[u] = in ? KSG_LD_PART(&part_d32[((int64_t)(s0 + u) * kcap) * stride + off]) : CUDART_INF_F;
          tails[u] = in ? KSG_LD_PART(&part_d32[((int64_t)(s0 + u) * kcap + (kcap - 1)) * stride + off]) : CUDART_INF_F;
      }
#pragma unroll
      for (int u = 0; u < REFINE_BATCH; ++u) {
        const int s = s0 + u;
        if (s >= s_end || overflow) break;
        // a full list whose tail is still inside the band may have dropped in-band candidates
        if ((double)tails[u] <= T) { overflow = true; break; }
        if (!((double)head[u] <= T)) continue;  // ascending list: nothing of it is inside the band
        for (int r = 0; r < kcap; ++r) {
            const int64_t o = ((int64_t)s * kcap + r) * stride + off;
            const float v = r == 0 ? head[u] : KSG_LD_PART(&part_d32[o]);
            if (!((double)v <= T)) break;
            if (++examined > REFINE_CAP) { overflow = true; break; }
            const int64_t j = KSG_LD_PART(&part_idx[o]);
            double m = 0.0;
            for (int d = 0; d < dim; ++d) m = fmax(m, fabs(qc[d] - sorted[(int64_t)d * ld + j]));
            int pos = got < kp ? got : kp - 1;
            if (ra.idx_out) {
                const int32_t o_j = ra.perm[j];
                if (got == kp && !(m < exact[kp - 1] || (m == exact[kp - 1] && o_j < exact_o[kp - 1]))) continue;
                while (pos > 0 && (exact[pos - 1] > m || (exact[pos - 1] == m && exact_o[pos - 1] > o_j))) {
                    exact[pos] = exact[pos - 1]; exact_j[pos] = exact_j[pos - 1]; exact_o[pos] = exact_o[pos - 1];
                    --pos;
                }
                exact[pos] = m; exact_j[pos] = (int32_t)j; exact_o[pos] = o_j;
            } else {
                if (got == kp && !(m < exact[kp - 1])) continue;
                while (pos > 0 && exact[pos - 1] > m) { exact[pos] = exact[pos - 1]; --pos; }
                exact[pos] = m;
            }
            if (got < kp) ++got;
        }
      }
    }
    if (overflow || got < kp) {
        ra.flags[qi] = 1;
        atomicAdd(ra.n_flagged, 1);
        ra.eps_out[qi] = CUDART_NAN;
    } else {
        ra.flags[qi] = 0;
        ra.eps_out[qi] = exact[kp - 1];
        if (ra.idx_out)
            for (int r = 0; r < kp; ++r) ra.idx_out[qi * kp + r] = exact_j[r];
    }
}

// The refine as a launch of its own: the dense filter's split lists, and the windowed filter's unit
// lists when the in-kernel refine is switched off (RefineArgs::fused == 0).
static __global__ void __launch_bounds__(128)
knn_refine_kernel(const float* __restrict__ part_d32, const int32_t* __restrict__ part_idx, int kcap, int splits,
                  int kp, const double* __restrict__ sorted, int64_t ld, int dim, int64_t n, int64_t q_begin,
                  int64_t nq, const int32_t* __restrict__ q_index, const double* __restrict__ maxabs,
                  const float* __restrict__ seed_thr, const int32_t* __restrict__ unit_first, int qb,
                  const int32_t* __restrict__ perm, int64_t* __restrict__ idx_out,
                  const double* __restrict__ qsorted, int64_t qld, const double* __restrict__ center,
                  double* __restrict__ eps_out, int32_t* __restrict__ flags, int32_t* __restrict__ n_flagged,
                  const int32_t* __restrict__ exact32) {
    // qsorted (optional, cross-set search): the float64 query coordinates come from this [dim][qld]
    // array (rows q_begin + qi) instead of the candidate set itself.
    // idx_out (optional, [nq][kp] int64): the k' nearest points as positions in the prepared
    // order, ascending by (exact distance, original sample index) -- the order the dense float64
    // kernel's merge produces.  All points tied with the k'-th distance are among the examined
    // candidates whenever the query is not flagged, so the tie-break is well defined.
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    (void)n;
    // list s, rank r of this query sits at ((s * kcap + r) * stride + off):
    //   dense launch   : s = candidate split,  [split][rank][nq]
    //   windowed launch: s = work unit of the query's block,  [unit][rank][QB]
    int s_begin = 0, s_end = splits;
    int64_t stride = nq, off = qi;
    if (unit_first) {
        const int64_t b = qi / qb;
        s_begin = unit_first[b];
        s_end = s_begin + unit_first[(nq + qb - 1) / qb + 1 + b];  // (counts sit behind the starts)
        stride = qb;
        off = qi - b * qb;
    }
    RefineArgs ra;
    ra.fused = 0; ra.kp = kp; ra.dim = dim; ra.sorted = sorted; ra.ld = ld; ra.qsorted = qsorted; ra.qld = qld;
    ra.center = center; ra.maxabs = maxabs; ra.perm = perm; ra.idx_out = idx_out; ra.eps_out = eps_out;
    ra.flags = flags; ra.n_flagged = n_flagged; ra.exact32 = exact32;
    refine_merge_lists(part_d32, part_idx, kcap, s_begin, s_end, stride, off, seed_thr != nullptr,
                       seed_thr ? seed_thr[qi] : 0.0f, q_index ? (int64_t)q_index[qi] : q_begin + qi, qi, ra);
}

// The refine of ONE ascending list that still sits in shared memory (called by knn_units_kernel for
// blocks whose tile list is a single work unit -- nearly all of them): same decision rule as
// knn_refine_kernel, without the merge.  col_d / col_i: the query's column (entry r at r * stride;
// unused entries +inf / -1); seed: the threshold the list was filled under (< 0: the planning pass
// set the query aside).  All in-band float64 rows are fetched as one batch of independent loads.
static __device__ __noinline__ void refine_single_list(const float* col_d, const int32_t* col_i, int stride, int kcap,
                                                float seed, int64_t row_q, int64_t qi, const RefineArgs& ra) {
    if (seed < 0.0f) {
        ra.flags[qi] = 2;
        atomicAdd(ra.n_flagged, 1);
        ra.eps_out[qi] = CUDART_NAN;
        return;
    }
    const int kp = ra.kp, dim = ra.dim;
    if (!fp32_range_ok(ra.maxabs, dim)) {  // shadow rows may have overflowed: float64 kernels
        ra.flags[qi] = 1;
        atomicAdd(ra.n_flagged, 1);
        ra.eps_out[qi] = CUDART_NAN;
        return;
    }
    const float v_kp = col_d[(size_t)(kp - 1) * stride];
    if (!(v_kp < CUDART_INF_F)) {  // fewer than k' points exist (k+1 > n): cKDTree pads with inf
        ra.eps_out[qi] = CUDART_INF;
        ra.flags[qi] = ra.idx_out ? 1 : 0;
        if (ra.idx_out) atomicAdd(ra.n_flagged, 1);
        return;
    }
    if (ra.exact32 && *ra.exact32 && !ra.idx_out) {
        // quantised / integer data: every fp32 distance is the float64 distance itself (prepare.cuh,
        // exact_flag_kernel) -- the k'-th smallest of the list is the exact radius, ties and all
        ra.flags[qi] = 0;
        ra.eps_out[qi] = (double)v_kp;
        return;
    }
    double qc[KSG_MAX_DIM];
    double mq = 0.0;
    for (int d = 0; d < dim; ++d) {
        qc[d] = ra.qsorted ? ra.qsorted[(int64_t)d * ra.qld + row_q] : ra.sorted[(int64_t)d * ra.ld + row_q];
        mq = fmax(mq, fabs(qc[d] - ra.center[d]));
    }
    mq = mq * (1.0 + 1e-6) + 1e-300;
    const double r32 = (double)v_kp;
    const double T = r32 + 2.0 * band_query_f64(mq, 2.0 * r32);
    const double tail = (double)col_d[(size_t)(kcap - 1) * stride];
    // full list with its tail inside the band, or a band that reaches the seeded threshold:
    // in-band candidates may be missing
    bool undecided = tail <= T || !(T < (double)seed);
    if (!undecided && !ra.idx_out) {
        // Radius only.  With |d64 - d32| <= delta the exact radius r64 lies in [r32 - delta, r32 + delta]:
        // an entry with d32 < r32 - 2 delta is strictly below r64, one with d32 > r32 + 2 delta strictly
        // above.  So r64 is the (k' - a)-th smallest EXACT distance among the entries inside
        // [r32 - 2 delta, r32 + 2 delta], a = the number of entries below that band -- on continuous
        // data the band holds the k'-th entry alone and the refine is ONE float64 distance.
        const double lo = r32 - 2.0 * band_query_f64(mq, 2.0 * r32);  // (the band's 1 % slack covers the rounding)
        int a = 0;
        while (a < kp - 1 && (double)col_d[(size_t)a * stride] < lo) ++a;
        int e = kp;  // entry k' - 1 (= r32) is inside
        while (e < kcap && (double)col_d[(size_t)e * stride] <= T) ++e;
        const int b = e - a, want = kp - a;  // 1 <= want <= b
        constexpr int BAND_REGS = 4;
        if (b <= BAND_REGS) {
            double mb[BAND_REGS];
#pragma unroll
            for (int r = 0; r < BAND_REGS; ++r) {
                mb[r] = CUDART_INF;
                if (r < b) {
                    const int64_t j = col_i[(size_t)(a + r) * stride];
                    double v = 0.0;
                    for (int d = 0; d < dim; ++d) v = fmax(v, fabs(qc[d] - ra.sorted[(int64_t)d * ra.ld + j]));
                    mb[r] = v;
                }
            }
            // the want-th smallest of mb[0..b): the value with fewer than `want` values below it and at
            // least `want` values not above it
            double ans = mb[0];
#pragma unroll
            for (int r = 0; r < BAND_REGS; ++r) {
                int less = 0, le = 0;
#pragma unroll
                for (int t = 0; t < BAND_REGS; ++t) {
                    less += (t < b && mb[t] < mb[r]) ? 1 : 0;
                    le += (t < b && mb[t] <= mb[r]) ? 1 : 0;
                }
                if (r < b && less < want && want <= le) ans = mb[r];
            }
            ra.flags[qi] = 0;
            ra.eps_out[qi] = ans;
            return;
        }
    }
    double m[KNN_MAX_KCAP];
    int32_t mj[KNN_MAX_KCAP], mo[KNN_MAX_KCAP];
    int nb = 0;
    if (!undecided) {
        while (nb < kcap && (double)col_d[(size_t)nb * stride] <= T) ++nb;  // >= k' entries
        for (int r = 0; r < nb; ++r) mj[r] = col_i[(size_t)r * stride];
        for (int r = 0; r < nb; ++r) {
            double v = 0.0;
            for (int d = 0; d < dim; ++d) v = fmax(v, fabs(qc[d] - ra.sorted[(int64_t)d * ra.ld + mj[r]]));
            m[r] = v;
        }
        if (ra.idx_out)
            for (int r = 0; r < nb; ++r) mo[r] = ra.perm[mj[r]];
        // ascending by (exact distance, original sample index): insertion sort of <= kcap entries
        for (int a = 1; a < nb; ++a) {
            const double v = m[a];
            const int32_t j = mj[a], o = ra.idx_out ? mo[a] : 0;
            int b = a - 1;
            while (b >= 0 && (m[b] > v || (ra.idx_out && m[b] == v && mo[b] > o))) {
                m[b + 1] = m[b]; mj[b + 1] = mj[b];
                if (ra.idx_out) mo[b + 1] = mo[b];
                --b;
            }
            m[b + 1] = v; mj[b + 1] = j;
            if (ra.idx_out) mo[b + 1] = o;
        }
        undecided = nb < kp;
    }
    if (undecided) {
        ra.flags[qi] = 1;
        atomicAdd(ra.n_flagged, 1);
        ra.eps_out[qi] = CUDART_NAN;
        return;
    }
    ra.flags[qi] = 0;
    ra.eps_out[qi] = m[kp - 1];
    if (ra.idx_out)
        for (int r = 0; r < kp; ++r) ra.idx_out[qi * kp + r] = mj[r];
}

}  // namespace ksg

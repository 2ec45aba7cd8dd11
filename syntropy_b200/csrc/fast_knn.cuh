// K1 fast path: k-NN radius with an fp32 tile filter and an exact fp64 refine.
//
// Filter (knn_f32_kernel).  Works on the prepared set (prepare.cuh): fp32 AoS shadow
// rows, sorted by the primary coordinate, +inf padded.  One thread owns Q queries whose
// coordinates live in registers as packed pairs; candidate tiles are bulk-copied
// (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) into a STAGES-deep ring of
// shared-memory buffers; all lanes read the same candidate (LDS.128 broadcasts).
// Per (query, candidate) the hot loop is   DP/2 FADD2 (packed subtract, two dimensions
// per issue slot) + ~DP/2 FMNMX/FMNMX3 (|.| folded into the operands) + 1/2 FMNMX3
// (running minimum over two candidates) -- no compare, no branch.  Every BATCH
// candidates the running minima are tested against the per-query thresholds; only on a
// hit is the batch re-walked to insert (d32, index) into the query's sorted list in
// shared memory.  Tiles are visited outward from the query block's own tile, so the
// thresholds are tight after the first tile; in windowed mode a side is closed as soon
// as the primary-coordinate gap to its next tile exceeds every threshold (+ band).
//
// Refine (knn_refine_kernel).  The list keeps the KCAP = k' + margin smallest fp32
// distances.  With r32 the k'-th smallest of them and delta the fp32 error bound,
// every point whose exact distance is <= the exact radius has d32 <= r32 + 2*delta;
// if the list provably contains all of those (its tail lies above that bound) the exact
// radius is the k'-th smallest float64 distance among them.  Otherwise the query is
// flagged and recomputed by the dense fp64 kernel.
#pragma once

#include "common.cuh"
#include "prepare.cuh"
#include "tile_queue.cuh"

namespace ksg {

constexpr int KNN_BATCH = 16;   // candidates between threshold checks
constexpr int KNN_MARGIN = 3;   // list entries kept beyond k'
constexpr int KNN_SEED_W = 24;  // rows on each side of a query used to seed its threshold
constexpr float KNN_DEFER_CUT = 4.0f;   // defer a query whose seeded radius exceeds CUT x the CTA mean
constexpr float KNN_DEFERRED = -2.0f;   // sentinel distance marking a deferred query's list
constexpr int KNN_PLAN_CAP = 1024;      // tile-list entries per query block (more: in-kernel walk)

// ---- fp32 error band of the shadow copy (see DESIGN.md "Exactness")
// |d32 - d64| <= 2^-24 * (2*M + |d|) * (1 + tiny) per dimension; kept as a pair of
// coefficients so that band(T) = c0 + c1 * T, evaluated with upward slack.
struct Band {
    float c0, c1;
};
// band(T) = c0 + c1*T  >=  2 * delta(T), with ~1% upward slack for its own fp32 evaluation
__device__ __forceinline__ Band make_band(const double* __restrict__ maxabs, int dim) {
    double mmax = 0.0;
    for (int d = 0; d < dim; ++d) mmax = fmax(mmax, maxabs[d]);
    Band b;
    b.c0 = __double2float_ru(2.0 * 5.9604644775390625e-08 * 2.0 * mmax * 1.02 + 1e-37);
    b.c1 = __double2float_ru(2.0 * 5.9604644775390625e-08 * 1.02);
    return b;
}
__host__ __device__ inline double band_f64(double maxabs, double t) {
    return 5.9604644775390625e-08 * (2.0 * maxabs + t) * 1.001 + 1e-37;
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

// ---- planning pass of the windowed engine.  One CTA per query block: seeds every query's
// threshold from its neighbours in the prepared order, derives the block's pruning radius, tests
// EVERY candidate tile's bounding box against the block's query box and writes the survivors
// (tile, gap) in walk order.  The filter kernel then streams exactly that list (re-checking each
// gap against its shrinking radius) with no in-kernel search rounds, and the host side of the ABI
// launches the blocks longest-list-first so that no heavy block is left for the tail.
template <int DP, int Q>
static __global__ void __launch_bounds__(FAST_THREADS)
knn_plan_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box, int64_t n_tiles,
                int64_t q_begin, int64_t nq, int kcap, const double* __restrict__ maxabs, int dim,
                float* __restrict__ seed_thr, int2* __restrict__ plan, int32_t* __restrict__ plan_count) {
    constexpr int QB = FAST_THREADS * Q;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* lst_d = reinterpret_cast<float*>(smem_raw);  // kcap * QB (values only)
    __shared__ TileQueueShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t qb0 = (int64_t)blockIdx.x * QB;
    const Band band = make_band(maxabs, dim);
    const int64_t n_rows = n_tiles * FAST_TILE;

    float blo[DP], bhi[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
    float rmax = -CUDART_INF_F;
#pragma unroll 1
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if (qi >= nq) continue;
        const int64_t row = q_begin + qi;
        float2 q[DP / 2];
        load_row<DP>(shadow + row * DP, q);
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            blo[2 * h] = fminf(blo[2 * h], q[h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[h].x);
            blo[2 * h + 1] = fminf(blo[2 * h + 1], q[h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[h].y);
        }
        const int col = j * FAST_THREADS + tid;
        for (int r = 0; r < kcap; ++r) lst_d[(size_t)r * QB + col] = CUDART_INF_F;
        float thr = CUDART_INF_F;
#pragma unroll 1
        for (int off = 1; off <= KNN_SEED_W; ++off)
#pragma unroll 1
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                const int64_t r2 = row + sgn * off;
                if (r2 < 0 || r2 >= n_rows) continue;
                float2 cc[DP / 2];
                load_row<DP>(shadow + r2 * DP, cc);
                const float m = cheb32<DP>(q, cc);
                if (m < thr) {
                    int r = kcap - 1;
                    while (r > 0 && lst_d[(size_t)(r - 1) * QB + col] > m) {
                        lst_d[(size_t)r * QB + col] = lst_d[(size_t)(r - 1) * QB + col];
                        --r;
                    }
                    lst_d[(size_t)r * QB + col] = m;
                    thr = lst_d[(size_t)(kcap - 1) * QB + col];
                }
            }
        thr = thr + (band.c0 + band.c1 * thr);  // inf stays inf
        seed_thr[qi] = thr;
        rmax = fmaxf(rmax, thr);
    }
    block_query_box<DP>(blo, bhi, sh);
    const float R = block_max(rmax, sh);

    const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
    const int64_t own = mid / FAST_TILE;
    int2* out = plan + (int64_t)blockIdx.x * KNN_PLAN_CAP;
    int total = 0;
    for (int64_t pos = 0; pos < n_tiles; pos += FAST_THREADS) {
        bool keep = false;
        int64_t t = 0;
        float dist = 0.0f;
        if (pos + tid < n_tiles) {
            t = walk_tile(pos + tid, own, n_tiles);
            float gap[DP];
            box_gaps<DP>(tile_box, t, sh, gap);
            dist = gap[0];
#pragma unroll
            for (int d = 1; d < DP; ++d) dist = fmaxf(dist, gap[d]);
            keep = !(dist > R);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) sh.warp_count[warp] = __popc(ballot);
        __syncthreads();
        int offset = total, round_total = 0;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w) {
            if (w < warp) offset += sh.warp_count[w];
            round_total += sh.warp_count[w];
        }
        if (keep) {
            const int k = offset + __popc(ballot & ((1u << lane) - 1u));
            if (k < KNN_PLAN_CAP) out[k] = make_int2((int)t, __float_as_int(dist));
        }
        total += round_total;
    }
    if (tid == 0) plan_count[blockIdx.x] = total <= KNN_PLAN_CAP ? total : -1;
}

// part_d32 / part_idx: [split][rank][nq]
template <int DP, int Q, int STAGES>
static __global__ void __launch_bounds__(FAST_THREADS)
knn_f32_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box, int64_t n_tiles,
               int64_t q_begin, int64_t nq, const int32_t* __restrict__ q_index, int kcap, int windowed,
               const double* __restrict__ maxabs, int dim, const float* __restrict__ seed_thr,
               const int2* __restrict__ plan, const int32_t* __restrict__ plan_count,
               const int32_t* __restrict__ block_order, float* __restrict__ part_d32,
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
    // planned launches run the query blocks longest-tile-list-first
    const int64_t qblock = block_order ? (int64_t)block_order[blockIdx.x] : (int64_t)blockIdx.x;
    const int64_t qb0 = qblock * QB;  // first query of the block (relative to q_begin)
    const Band band = make_band(maxabs, dim);

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
    // Starting from it (+ band, so that those very rows re-enter the list when their tile is
    // scanned) removes the flood of insertions a threshold of +inf causes and gives the tile
    // queue a finite pruning radius from its first round.
    if (seed_thr) {
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j]) thr[j] = seed_thr[qb0 + j * FAST_THREADS + tid];  // from the planning pass
    } else {
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
            if (live[j]) thr[j] = thr[j] + (band.c0 + band.c1 * thr[j]);  // inf stays inf
        reset_lists();
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    block_query_box<DP>(blo, bhi, sh);  // (barriers inside: also publishes the mbarrier init)

    float R = CUDART_INF_F;  // CTA pruning radius: max over its queries of threshold + band
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
            if (live[j]) v = fmaxf(v, thr[j] + (band.c0 + band.c1 * thr[j]));
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
                            if (m < thr[j]) thr[j] = insert(j, m, (int32_t)(tile_base + b + c));
                        }
                    }
                }
            }
        }
    };

    // ---- outlier deferral (windowed engine): one isolated query next to a dense region would
    // blow the CTA's pruning radius up to its own huge neighbour distance and make all 128*Q
    // queries walk most of the point set.  Queries whose seeded radius is far above the CTA's
    // mean AND above the CTA's own extent are taken out here (sentinel in the output list) and
    // recomputed by a dense launch over the compacted list of such queries.
    bool deferred[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) deferred[j] = false;
    if (windowed) {
        float sum = 0.0f, cnt = 0.0f;
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j] && thr[j] < CUDART_INF_F) { sum += thr[j]; cnt += 1.0f; }
        sum = block_sum(sum, sh);
        cnt = block_sum(cnt, sh);
        float extent = 0.0f;
#pragma unroll
        for (int d = 0; d < DP; ++d) extent = fmaxf(extent, sh.qhi[d] - sh.qlo[d]);
        const float cut = fmaxf(KNN_DEFER_CUT * (sum / fmaxf(cnt, 1.0f)), 2.0f * extent);
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j] && thr[j] < CUDART_INF_F && thr[j] > cut) {
                deferred[j] = true;
                live[j] = false;   // no longer part of the pruning radius
                thr[j] = -1.0f;    // never inserts
            }
        refresh();  // seeded thresholds -> finite pruning radius from the first round
    }

    int tiles_done = 0;
    const int planned = (plan && windowed) ? plan_count[qblock] : -1;
    if (planned >= 0) {
        // stream the planned tile list; entries whose gap exceeds the current radius are skipped
        const int2* lst = plan + qblock * KNN_PLAN_CAP;
        int64_t issue_cnt = 0, consume_cnt = 0;
        int64_t ring_tile[STAGES];
        int next = 0;
        for (;;) {
            while (issue_cnt - consume_cnt < STAGES && next < planned) {
                const int2 e = lst[next++];
                if (__int_as_float(e.y) > R) continue;
                const int slot = (int)(issue_cnt % STAGES);
#pragma unroll
                for (int s = 0; s < STAGES; ++s)
                    if (s == slot) ring_tile[s] = e.x;
                if (tid == 0) {
                    mbar_expect_tx(&full_bar[slot], TILE_FLOATS * 4);
                    bulk_g2s(tiles + (size_t)slot * TILE_FLOATS, shadow + (int64_t)e.x * TILE_FLOATS, TILE_FLOATS * 4,
                             &full_bar[slot]);
                }
                ++issue_cnt;
            }
            if (consume_cnt == issue_cnt) break;
            const int slot = (int)(consume_cnt % STAGES);
            int64_t t = 0;
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (s == slot) t = ring_tile[s];
            mbar_wait(&full_bar[slot], (uint32_t)((consume_cnt / STAGES) & 1));
            compute(tiles + (size_t)slot * TILE_FLOATS, t);
            __syncthreads();  // tile consumed: its slot may be refilled
            ++consume_cnt;
            ++tiles_done;
            if ((tiles_done & 3) == 0) refresh();
        }
    } else {
        // the block's own tile: the one holding its middle query
        const int64_t mid = q_index ? 0 : q_begin + min(qb0 + QB / 2, nq - 1);
        tiles_done = run_tile_queue<DP, STAGES>(shadow, tiles, full_bar, sh, mid / FAST_TILE, n_tiles,
                                                (int)blockIdx.y, (int)gridDim.y, windowed != 0, test, compute,
                                                refresh);
    }
    if (tid == 0 && tile_counter) atomicAdd(tile_counter, (unsigned long long)tiles_done);

    // emit the lists (ascending), one per split
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if (qi < nq) {
            const int col = j * FAST_THREADS + tid;
            for (int r = 0; r < kcap; ++r) {
                const int64_t o = ((int64_t)blockIdx.y * kcap + r) * nq + qi;
                part_d32[o] = deferred[j] ? KNN_DEFERRED : lst_d[(size_t)r * QB + col];
                part_idx[o] = lst_i[(size_t)r * QB + col];
            }
        }
    }
}

// cost[b] = planned tile count (overflowed lists count as the heaviest); iota[b] = b
static __global__ void plan_cost_kernel(const int32_t* __restrict__ plan_count, int64_t n_blocks,
                                        int32_t* __restrict__ cost, int32_t* __restrict__ iota) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int32_t c = plan_count[b];
    cost[b] = c < 0 ? 0x7fffffff : c;
    iota[b] = (int32_t)b;
}

// Exact refine: see the file header.  flags[qi] = 1 marks a query for the fp64 fallback.
constexpr int REFINE_CAP = 96;  // union entries examined per query

static __global__ void __launch_bounds__(128)
knn_refine_kernel(const float* __restrict__ part_d32, const int32_t* __restrict__ part_idx, int kcap, int splits,
                  int kp, const double* __restrict__ sorted, int64_t ld, int dim, int64_t n, int64_t q_begin,
                  int64_t nq, const int32_t* __restrict__ q_index, const double* __restrict__ maxabs,
                  double* __restrict__ eps_out, int32_t* __restrict__ flags, int32_t* __restrict__ n_flagged) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    if (part_d32[qi] == KNN_DEFERRED) {  // taken out by the windowed filter: needs the dense launch
        flags[qi] = 2;
        atomicAdd(n_flagged, 1);
        eps_out[qi] = CUDART_NAN;
        return;
    }
    const int64_t row_q = q_index ? (int64_t)q_index[qi] : q_begin + qi;
    double mmax = 0.0;
    for (int d = 0; d < dim; ++d) mmax = fmax(mmax, maxabs[d]);

    // k'-th smallest fp32 distance over the union of the split lists
    float best[KSG_MAX_KPRIME];
    int have = 0;
    for (int s = 0; s < splits; ++s)
        for (int r = 0; r < kcap; ++r) {
            const float v = part_d32[((int64_t)s * kcap + r) * nq + qi];
            if (have == kp && !(v < best[kp - 1])) break;
            int pos = have < kp ? have : kp - 1;
            while (pos > 0 && best[pos - 1] > v) { best[pos] = best[pos - 1]; --pos; }
            best[pos] = v;
            if (have < kp) ++have;
        }
    const double r32 = (have == kp) ? (double)best[kp - 1] : CUDART_INF;
    if (!(r32 < CUDART_INF)) {  // fewer than k' points exist (k+1 > n): cKDTree pads with inf
        eps_out[qi] = CUDART_INF;
        flags[qi] = 0;
        return;
    }
    const double delta = band_f64(mmax, 2.0 * r32);
    const double T = r32 + 2.0 * delta;

    bool overflow = false;
    double exact[KSG_MAX_KPRIME];
    int got = 0, examined = 0;
    double qc[KSG_MAX_DIM];
    for (int d = 0; d < dim; ++d) qc[d] = sorted[(int64_t)d * ld + row_q];
    for (int s = 0; s < splits && !overflow; ++s) {
        // a full list whose tail is still inside the band may have dropped in-band candidates
        const float tail = part_d32[((int64_t)s * kcap + (kcap - 1)) * nq + qi];
        if ((double)tail <= T) { overflow = true; break; }
        for (int r = 0; r < kcap; ++r) {
            const int64_t o = ((int64_t)s * kcap + r) * nq + qi;
            const float v = part_d32[o];
            if (!((double)v <= T)) break;
            if (++examined > REFINE_CAP) { overflow = true; break; }
            const int64_t j = part_idx[o];
            double m = 0.0;
            for (int d = 0; d < dim; ++d) m = fmax(m, fabs(qc[d] - sorted[(int64_t)d * ld + j]));
            int pos = got < kp ? got : kp - 1;
            if (got == kp && !(m < exact[kp - 1])) continue;
            while (pos > 0 && exact[pos - 1] > m) { exact[pos] = exact[pos - 1]; --pos; }
            exact[pos] = m;
            if (got < kp) ++got;
        }
    }
    if (overflow || got < kp) {
        flags[qi] = 1;
        atomicAdd(n_flagged, 1);
        eps_out[qi] = CUDART_NAN;
    } else {
        flags[qi] = 0;
        eps_out[qi] = exact[kp - 1];
    }
}

}  // namespace ksg

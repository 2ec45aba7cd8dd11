// K2, single-subspace windowed pass: strict-radius neighbour counts in the FULL space of a
// prepared set (the host prepares every multi-dimensional marginal in its own space-filling
// order, so "the subspace" is simply all coordinates of that set).
//
// Same machinery as the windowed k-NN filter (fast_knn.cuh): a planning pass lists, per
// block of 128*Q consecutive queries, the candidate tiles whose bounding box comes closer
// than the block's largest threshold; the lists are cut into work units; a persistent kernel
// (4 filter warps + 1 copy warp, full / empty mbarrier ring, no block barrier inside a unit)
// pulls the units.  Every filter warp owns 32*Q consecutive queries and classifies each
// 32-row sub-tile against its own query box:
//     outside   nearest box distance >= the warp's largest threshold   -> skipped
//     inside    farthest box distance <  the warp's smallest threshold -> +32 for every query
//     boundary  scanned: DP/2 FADD2 + ~DP/2 FMNMX + 1 FSET + 1 FADD per (query, candidate)
// fp32 subtraction is monotone, so both box distances bound every COMPUTED pair distance and
// the result is exactly #{j : d32(q, j) < T_q} -- the quantity count_fix_kernel then corrects
// to the float64 truth inside the error band.  Large marginal counts (thousands of points
// inside the joint radius) therefore cost only their boundary.
#pragma once

#include "fast_count.cuh"
#include "fast_knn.cuh"

namespace ksg {

template <int DP, int Q>
static __global__ void __launch_bounds__(FAST_THREADS)
count_plan_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box, int64_t n_tiles,
                  int64_t q_begin, int64_t nq, const float* __restrict__ thresholds, int2* __restrict__ plan,
                  int64_t plan_capacity, int32_t* __restrict__ plan_count, int32_t* __restrict__ plan_first,
                  int32_t* __restrict__ meta, int32_t* __restrict__ counts) {
    // Tiles are classified against the whole block here: OUTSIDE (nearest box distance >= the
    // block's largest threshold) -> not listed; INSIDE (farthest box distance < the block's smallest
    // threshold: every pair of the block with the tile is inside) -> not listed either, its 512
    // candidates are added to every query's count right here (the padded last tile has an infinite
    // box and is never inside); the rest is listed for the filter kernel.  With radii that hold
    // thousands of points (low-dimensional marginals of a high-dimensional joint radius, sample
    // entropy) most relevant tiles are INSIDE and are never even fetched.
    constexpr int QB = FAST_THREADS * Q;
    __shared__ TileQueueShared sh;
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t qb0 = (int64_t)blockIdx.x * QB;

    float blo[DP], bhi[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
    float tmax = -CUDART_INF_F, tmin = CUDART_INF_F;
#pragma unroll 1
    for (int j = 0; j < Q; ++j) {
        const int64_t qi = qb0 + j * FAST_THREADS + tid;
        if (qi >= nq) continue;
        float2 q[DP / 2];
        load_row<DP>(shadow + (q_begin + qi) * DP, q);
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            blo[2 * h] = fminf(blo[2 * h], q[h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[h].x);
            blo[2 * h + 1] = fminf(blo[2 * h + 1], q[h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[h].y);
        }
        tmax = fmaxf(tmax, thresholds[qi]);
        tmin = fminf(tmin, thresholds[qi]);
    }
    block_query_box<DP>(blo, bhi, sh);
    const float R = block_max(tmax, sh);
    const float Tmin = -block_max(-tmin, sh);

    const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
    const int64_t own = mid / FAST_TILE;
    // nearest box distance; `inside` = the farthest one is below every threshold of the block
    auto gap_of = [&](int64_t pos, int64_t& t, bool& inside) -> float {
        t = walk_tile(pos, own, n_tiles);
        float near = 0.0f, far = 0.0f;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const float2 b = reinterpret_cast<const float2*>(tile_box)[t * DP + d];
            const float a = b.x - sh.qhi[d], c = sh.qlo[d] - b.y;  // > 0 when the boxes are apart
            near = fmaxf(near, fmaxf(a, c));
            far = fmaxf(far, fmaxf(-a, -c));
        }
        inside = far < Tmin;
        return near;
    };
    int mine = 0, mine_inside = 0;
    for (int64_t pos = tid; pos < n_tiles; pos += FAST_THREADS) {
        int64_t t;
        bool inside;
        const bool needed = gap_of(pos, t, inside) < R;
        mine += needed && !inside;
        mine_inside += needed && inside;
    }
    const int total = (int)(block_sum((float)mine, sh) + 0.5f);
    const int n_inside = (int)(block_sum((float)mine_inside, sh) + 0.5f);

    if (tid == 0) {
        int base = -1;
        if (total > 0 && 2 * (int64_t)total <= n_tiles) {
            const unsigned long long at = atomicAdd(reinterpret_cast<unsigned long long*>(meta + 4),
                                                    (unsigned long long)total);
            base = at + (unsigned long long)total <= (unsigned long long)plan_capacity ? (int)at : -1;
        }
        s_base = base;
    }
    __syncthreads();
    const int base = s_base;
    if (total > 0 && base < 0) {  // most tiles are needed (or the pool is full): implicit list
        // (every tile, the INSIDE ones included: the filter kernel classifies them again per warp)
        if (tid == 0) { plan_count[blockIdx.x] = (int32_t)n_tiles; plan_first[blockIdx.x] = -1; }
        return;
    }
    if (n_inside > 0)
        for (int j = 0; j < Q; ++j) {
            const int64_t qi = qb0 + j * FAST_THREADS + tid;
            if (qi < nq) counts[qi] += n_inside * FAST_TILE;  // (the filter kernel's atomics come later)
        }
    int2* out = plan + base;
    int done = 0;
    for (int64_t pos = 0; pos < n_tiles; pos += FAST_THREADS) {
        bool keep = false;
        int64_t t = 0;
        float dist = 0.0f;
        if (pos + tid < n_tiles) {
            bool inside;
            dist = gap_of(pos + tid, t, inside);
            keep = dist < R && !inside;
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

// counts[nq] (int32) are ADDED to with atomicAdd (several units serve one block).
template <int DP, int Q, int STAGES>
static __global__ void __launch_bounds__(UNITS_THREADS)
count_units_kernel(const float* __restrict__ shadow, const float* __restrict__ sub_box, int64_t n_tiles,
                   int64_t q_begin, int64_t nq, const float* __restrict__ thresholds,
                   const int2* __restrict__ plan, const int32_t* __restrict__ plan_count,
                   const int32_t* __restrict__ plan_first, const int32_t* __restrict__ unit_first,
                   const int32_t* __restrict__ unit_block, int32_t* __restrict__ meta,
                   int32_t* __restrict__ counts, unsigned long long* __restrict__ pair_counter) {
    constexpr int QB = FAST_THREADS * Q;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    constexpr int BOX_FLOATS = FAST_NSUB * DP * 2;
    constexpr int SLOT_FLOATS = TILE_FLOATS + BOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);  // STAGES * SLOT_FLOATS
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int s_unit;
    __shared__ int s_tile[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == FAST_THREADS / 32;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], FAST_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_units = meta[0], chunk = meta[1];
    int slot = 0;
    uint32_t phase_bits = producer ? 0xffffffffu : 0u;
    auto advance = [&]() {
        phase_bits ^= 1u << slot;
        slot = slot + 1 == STAGES ? 0 : slot + 1;
    };
    unsigned long long scans_total = 0;

    for (;;) {
        if (tid == 0) s_unit = atomicAdd(&meta[2], 1);
        __syncthreads();
        const int u = s_unit;
        if (u >= n_units) break;
        const int64_t qblock = unit_block[u];
        const int first = (u - unit_first[qblock]) * chunk;
        const int count = plan_count[qblock];
        const int last = first + chunk < count ? first + chunk : count;
        const int64_t qb0 = qblock * QB;
        const int pfirst = plan_first[qblock];
        __syncthreads();  // s_unit has been read by everybody

        if (producer) {
            const int2* lst = plan + (pfirst < 0 ? 0 : pfirst);
            const int64_t own = (q_begin + min(qb0 + QB / 2, nq - 1)) / FAST_TILE;
            for (int next = first; next <= last; ++next) {  // all 32 lanes together; lane 0 issues
                int tile = -1;
                if (next < last) tile = pfirst < 0 ? (int)walk_tile(next, own, n_tiles) : lst[next].x;
                mbar_wait_relaxed(&empty_bar[slot], (phase_bits >> slot) & 1u);
                if (lane == 0) {
                    *reinterpret_cast<volatile int*>(&s_tile[slot]) = tile;
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

        // ---------------- filter warps: slot j of this thread = query qb0 + warp*32*Q + j*32 + lane
        float2 q[Q][DP / 2];
        float T[Q];
        int cnt[Q];
        float wlo[DP], whi[DP];
        float tmax = -CUDART_INF_F, tmin = CUDART_INF_F;
#pragma unroll
        for (int d = 0; d < DP; ++d) { wlo[d] = CUDART_INF_F; whi[d] = -CUDART_INF_F; }
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const int64_t qi = qb0 + warp * (32 * Q) + j * 32 + lane;
            const bool in = qi < nq;
            load_row<DP>(shadow + (q_begin + (in ? qi : 0)) * DP, q[j]);
            T[j] = in ? thresholds[qi] : -1.0f;  // dead slots count nothing
            cnt[j] = 0;
            if (in) {
                tmax = fmaxf(tmax, T[j]);
                tmin = fminf(tmin, T[j]);
#pragma unroll
                for (int h = 0; h < DP / 2; ++h) {
                    wlo[2 * h] = fminf(wlo[2 * h], q[j][h].x);         whi[2 * h] = fmaxf(whi[2 * h], q[j][h].x);
                    wlo[2 * h + 1] = fminf(wlo[2 * h + 1], q[j][h].y); whi[2 * h + 1] = fmaxf(whi[2 * h + 1], q[j][h].y);
                }
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
            tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, off));
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                wlo[d] = fminf(wlo[d], __shfl_xor_sync(0xffffffffu, wlo[d], off));
                whi[d] = fmaxf(whi[d], __shfl_xor_sync(0xffffffffu, whi[d], off));
            }
        }
        int inside_subtiles = 0;  // sub-tiles wholly inside every live query's ball (warp-uniform)

        for (;;) {
            mbar_wait(&full_bar[slot], (phase_bits >> slot) & 1u);
            const int t = *reinterpret_cast<volatile int*>(&s_tile[slot]);
            if (t >= 0) {
                const float* tile = tiles + (size_t)slot * SLOT_FLOATS;
                const float2* boxes = reinterpret_cast<const float2*>(tile + TILE_FLOATS);
#pragma unroll 1
                for (int sb = 0; sb < FAST_NSUB; ++sb) {
                    float near = 0.0f, far = 0.0f;
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        const float2 bx = boxes[sb * DP + d];
                        const float a = bx.x - whi[d], b = wlo[d] - bx.y;  // > 0 when the boxes are apart
                        near = fmaxf(near, fmaxf(a, b));
                        far = fmaxf(far, fmaxf(-a, -b));  // whi - lo, hi - wlo
                    }
                    if (!(near < tmax)) continue;                  // no pair can be inside
                    if (far < tmin) { ++inside_subtiles; continue; }  // every pair is inside
                    ++scans_total;
                    const float* rows = tile + sb * FAST_SUB * DP;
                    float acc[Q];
#pragma unroll
                    for (int j = 0; j < Q; ++j) acc[j] = 0.0f;
#pragma unroll 4
                    for (int c = 0; c < FAST_SUB; c += 2) {
                        float2 ca[DP / 2], cb[DP / 2];
                        load_two_rows<DP>(rows + c * DP, ca, cb);
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            // (FSET here: with one subspace the ALU pipe has room, the FMA pipe
                            //  -- FADD2 differences + the accumulate -- does not; measured)
                            acc[j] += (cheb32<DP>(q[j], ca) < T[j]) ? 1.0f : 0.0f;
                            acc[j] += (cheb32<DP>(q[j], cb) < T[j]) ? 1.0f : 0.0f;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < Q; ++j) cnt[j] += (int)acc[j];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            advance();
            if (t < 0) break;
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const int64_t qi = qb0 + warp * (32 * Q) + j * 32 + lane;
            const int c = cnt[j] + inside_subtiles * FAST_SUB;
            if (qi < nq && c) atomicAdd(&counts[qi], c);
        }
    }
    if (!producer && lane == 0 && pair_counter && scans_total)
        atomicAdd(pair_counter, scans_total * (unsigned long long)(FAST_SUB * 32 * Q));
}

}  // namespace ksg

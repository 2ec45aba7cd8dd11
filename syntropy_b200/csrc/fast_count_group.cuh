// K2, single-subspace windowed pass with 8-QUERY GROUPS inside the filter warps.
//
// count_units_kernel (fast_count_units.cuh) classifies every 32-row sub-tile against the box of a warp's
// 32 consecutive queries and the range [tmin, tmax] of their thresholds: "outside" needs the NEAREST
// box distance above the LARGEST threshold, "inside" the FARTHEST below the SMALLEST, everything else is
// scanned by all 32 queries.  In 4-D the ball of a marginal count is a few sub-tiles across, so nearly every
// sub-tile a 32-query warp touches is boundary: C3 scans 22 % of all n^2 pairs in its three marginal passes.
// Same idea as the grouped k-NN kernel (fast_knn_group.cuh): the warp keeps its 32 queries as FOUR groups of
// 8 consecutive ones, each with its own box and threshold range, so far more (group, sub-tile) pairs are
// settled by their class alone.
//
//   lane l = (stream s = l / 8, member m = l % 8).  The lane holds member m's query of ALL four groups in
//   registers (its own query is the one of group s) and a partial count per group.
//   classification: the 16 sub-tiles x 4 groups of a staged tile are tested by the 32 lanes at once, two
//   (sub-tile, group) pairs per lane; four ballots give per group a 16-bit "scan" and a 16-bit "inside" mask.
//   scan of (group g, sub-tile): all 32 lanes, lane (s, m) compares member m's query of group g with the
//   candidates c = 4 i + s, i = 0..7 (four neighbouring rows per shared-memory read, conflict-free): 8 steps,
//   256 pairs, the same pair rate as the whole-warp kernel.
//   end of a unit: the four streams' partial counts of a query are summed by two shuffles, its owner adds
//   32 per "inside" sub-tile of its group and issues ONE atomicAdd.
// Same planning pass (count_plan_kernel), unit table, tile ring and copy warp; identical counts.
#pragma once

#include <stdlib.h>

#include "fast_count_units.cuh"

namespace ksg {

template <int DP, int STAGES, int CG_GROUPS>
static __global__ void __launch_bounds__(UNITS_THREADS)
count_gunits_kernel(const float* __restrict__ shadow, const float* __restrict__ sub_box, int64_t n_tiles,
                    int64_t q_begin, int64_t nq, const float* __restrict__ thresholds,
                    const int2* __restrict__ plan, const int32_t* __restrict__ plan_count,
                    const int32_t* __restrict__ plan_first, const int32_t* __restrict__ unit_first,
                    const int32_t* __restrict__ unit_block, int32_t* __restrict__ meta,
                    int32_t* __restrict__ counts, unsigned long long* __restrict__ pair_counter) {
    static_assert(FAST_THREADS == 128 && FAST_NSUB == 16 && FAST_SUB == 32, "geometry of the lane / group mapping");
    static_assert(CG_GROUPS == 4 || CG_GROUPS == 8, "four groups of 8 queries or eight groups of 4");
    constexpr int CG_GQ = 32 / CG_GROUPS;  // queries per group = lanes per candidate stream
    constexpr int QB = FAST_THREADS;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    constexpr int BOX_FLOATS = FAST_NSUB * DP * 2;
    constexpr int SLOT_FLOATS = TILE_FLOATS + BOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);  // STAGES * SLOT_FLOATS
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int s_unit;
    __shared__ int s_tile[STAGES];
    __shared__ float2 s_gbox[FAST_THREADS / 32][CG_GROUPS][DP];  // (lo, hi) of every group's live queries
    __shared__ float2 s_gthr[FAST_THREADS / 32][CG_GROUPS];      // (smallest, largest) threshold of the group

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strm = lane / CG_GQ, mem = lane % CG_GQ;
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

        // ---------------- filter warps: member `mem` of every group, thresholds, group boxes
        float2 qm[CG_GROUPS][DP / 2];
        float Tm[CG_GROUPS];
        int acc[CG_GROUPS], inside_g[CG_GROUPS];
        __syncwarp();  // (the previous unit's readers of s_gbox / s_gthr are done)
#pragma unroll
        for (int g = 0; g < CG_GROUPS; ++g) {
            const int64_t qi = qb0 + warp * 32 + g * CG_GQ + mem;
            const bool in = qi < nq;
            load_row<DP>(shadow + (q_begin + (in ? qi : 0)) * DP, qm[g]);
            Tm[g] = in ? thresholds[qi] : -1.0f;  // dead slots count nothing
            acc[g] = 0;
            inside_g[g] = 0;
            float lo[DP], hi[DP];
            float tmax = in ? Tm[g] : -CUDART_INF_F, tmin = in ? Tm[g] : CUDART_INF_F;
#pragma unroll
            for (int h = 0; h < DP / 2; ++h) {
                lo[2 * h] = in ? qm[g][h].x : CUDART_INF_F;      hi[2 * h] = in ? qm[g][h].x : -CUDART_INF_F;
                lo[2 * h + 1] = in ? qm[g][h].y : CUDART_INF_F;  hi[2 * h + 1] = in ? qm[g][h].y : -CUDART_INF_F;
            }
#pragma unroll
            for (int off = 1; off < CG_GQ; off <<= 1) {  // over the eight members (every stream does the same)
                tmax = fmaxf(tmax, __shfl_xor_sync(FULL, tmax, off));
                tmin = fminf(tmin, __shfl_xor_sync(FULL, tmin, off));
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], off));
                    hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], off));
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int d = 0; d < DP; ++d) s_gbox[warp][g][d] = make_float2(lo[d], hi[d]);
                s_gthr[warp][g] = make_float2(tmin, tmax);
            }
        }
        __syncwarp();

        for (;;) {
            mbar_wait(&full_bar[slot], (phase_bits >> slot) & 1u);
            const int t = *reinterpret_cast<volatile int*>(&s_tile[slot]);
            if (t >= 0) {
                const float* tile = tiles + (size_t)slot * SLOT_FLOATS;
                const float2* boxes = reinterpret_cast<const float2*>(tile + TILE_FLOATS);
                // ---- classify 16 sub-tiles x CG_GROUPS groups: lane = (half, sub-tile), groups half * CG_GROUPS / 2 + k
                unsigned scan_m[CG_GROUPS], in_m[CG_GROUPS];
                const int sb_l = lane & 15, half = lane >> 4;
#pragma unroll
                for (int k = 0; k < CG_GROUPS / 2; ++k) {
                    const int g = half * (CG_GROUPS / 2) + k;
                    float near = 0.0f, far = 0.0f;
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        const float2 bx = boxes[sb_l * DP + d];
                        const float2 gb = s_gbox[warp][g][d];
                        const float a = bx.x - gb.y, b = gb.x - bx.y;  // > 0 when the boxes are apart
                        near = fmaxf(near, fmaxf(a, b));
                        far = fmaxf(far, fmaxf(-a, -b));
                    }
                    const float2 th = s_gthr[warp][g];
                    const bool reach = near < th.y;            // some pair may be inside
                    const bool all_in = reach && far < th.x;   // every pair is inside
                    const unsigned b_scan = __ballot_sync(FULL, reach && !all_in);
                    const unsigned b_in = __ballot_sync(FULL, all_in);
                    scan_m[k] = b_scan & 0xffffu;      scan_m[CG_GROUPS / 2 + k] = b_scan >> 16;
                    in_m[k] = b_in & 0xffffu;          in_m[CG_GROUPS / 2 + k] = b_in >> 16;
                }
#pragma unroll
                for (int g = 0; g < CG_GROUPS; ++g) {
                    inside_g[g] += __popc(in_m[g]);
                    unsigned m = scan_m[g];
#pragma unroll 1
                    while (m) {
                        const int sb = __ffs(m) - 1;
                        m &= m - 1u;
                        ++scans_total;
                        const float* rows = tile + (sb * FAST_SUB + strm) * DP;
                        float a = 0.0f;
#pragma unroll
                        for (int i = 0; i < FAST_SUB / CG_GROUPS; ++i) {  // candidates CG_GROUPS * i + strm
                            float2 cc[DP / 2];
                            load_row<DP>(rows + CG_GROUPS * i * DP, cc);
                            a += (cheb32<DP>(qm[g], cc) < Tm[g]) ? 1.0f : 0.0f;
                        }
                        acc[g] += (int)a;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            advance();
            if (t < 0) break;
        }
        // ---- the four streams' shares of every query -> its owner (lane = group * 8 + member)
        int mine = 0;
#pragma unroll
        for (int g = 0; g < CG_GROUPS; ++g) {
            int v = acc[g];
#pragma unroll
            for (int off = CG_GQ; off < 32; off <<= 1) v += __shfl_xor_sync(FULL, v, off);
            if (strm == g) mine = v + inside_g[g] * FAST_SUB;
        }
        const int64_t qi = qb0 + warp * 32 + lane;
        if (qi < nq && mine) atomicAdd(&counts[qi], mine);
    }
    if (!producer && lane == 0 && pair_counter && scans_total)
        atomicAdd(pair_counter, scans_total * (unsigned long long)(FAST_SUB * CG_GQ));
}

template <int DP>
static int launch_count_full_dp(const CountF32Args& a, const CountUnitsArgs& w) {
    constexpr int Q = COUNT_UNITS_Q, ST = FastGeom<DP>::STAGES;
    const int64_t qblocks = ceil_div(a.nq, FAST_THREADS * Q);
    if (cudaMemsetAsync(w.unit_meta, 0, 32, a.st) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return KSG_ECUDA; }
    count_plan_kernel<DP, Q><<<(unsigned)qblocks, FAST_THREADS, 0, a.st>>>(
        a.shadow, a.tile_box, a.n_tiles, a.q_begin, a.nq, a.thresholds, (int2*)w.plan, w.plan_capacity, w.plan_count,
        w.plan_first, w.unit_meta, a.counts);
    note_launch();
    knn_unit_table_kernel<<<1, 1024, 0, a.st>>>(w.plan_count, qblocks, w.max_units, COUNT_UNIT_MIN_TILES, w.unit_first,
                                                 w.unit_block, w.unit_meta);
    note_launch();
    // 8-query groups inside the filter warps (this file); KSG_B200_COUNT_GROUPS=0: whole warps (A/B runs, identity tests)
    static_assert(COUNT_UNITS_Q == 1, "the grouped kernel keeps one query per lane");
    static const int groups = [] { const char* e = getenv("KSG_B200_COUNT_GROUPS"); return e ? atoi(e) : 4; }();
    auto kern = groups == 8 ? count_gunits_kernel<DP, ST, 8> : groups == 4 ? count_gunits_kernel<DP, ST, 4> : count_units_kernel<DP, Q, ST>;
    const size_t smem = (size_t)ST * (FAST_TILE * DP + FAST_NSUB * DP * 2) * sizeof(float);
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    // (the occupancy query costs ~10 us of host time: remember the last answer per instantiation)
    static int cached_dev = -1, cached_per_sm = 0;
    int dev = 0;
    (void)cudaGetDevice(&dev);
    int per_sm = cached_per_sm;
    cudaError_t e = cudaSuccess;
    if (dev != cached_dev || per_sm < 1) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, UNITS_THREADS, smem);
        if (e != cudaSuccess || per_sm < 1) { (void)cudaGetLastError(); per_sm = 1; }
        cached_dev = dev; cached_per_sm = per_sm;
    }
    int64_t grid = (int64_t)per_sm * sm_count();
    if (grid > w.max_units) grid = w.max_units;
    kern<<<(unsigned)grid, UNITS_THREADS, smem, a.st>>>(a.shadow, w.sub_box, a.n_tiles, a.q_begin, a.nq, a.thresholds,
                                                       (const int2*)w.plan, w.plan_count, w.plan_first, w.unit_first,
                                                       w.unit_block, w.unit_meta, a.counts, a.tile_counter);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("count_units_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

}  // namespace ksg

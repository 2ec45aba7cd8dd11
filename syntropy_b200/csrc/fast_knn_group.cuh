// Windowed k-NN filter for many dimensions (DP >= 6): 8-query groups inside the filter warps.
//
// In 6 to 8 coordinates the work of knn_units_kernel is set by the QUERY GROUP a sub-tile is culled
// against -- the box of the warp's 32 queries inflated by the LARGEST of their thresholds -- not by the
// candidate boxes: even with the true radii a 32-query warp must scan 14 % (6-D) / 43 % (8-D) of all
// pairs at n = 2e5, a group of 8 consecutive queries 5 % / 19 % (tools/pruning_sim.py,
// profiles/r1aj_pruning_simulation.txt).  Here every filter warp keeps its 32 consecutive queries as
// FOUR groups of 8, each with its own box, threshold and surviving sub-tiles; a lane is (stream s =
// lane / 8, member m = lane % 8) and holds member m of every group.  A surviving (group, sub-tile) is
// scanned by all 32 lanes at once: stream s takes the candidates c = 4 i + s (four neighbouring rows per
// shared-memory read, conflict-free), 8 steps instead of 32, then the four streams' running minima are
// folded (two shuffles) and only a group that saw something closer walks the sub-tile again (stream 0)
// to insert.  Same tile ring, copy warp, planning pass, unit table, lists and refine as knn_units_kernel;
// results are identical.
#pragma once

#include "fast_knn.cuh"

namespace ksg {

constexpr int GU_GROUPS = 4;  // query groups per filter warp
constexpr int GU_GQ = 8;      // queries per group
constexpr int gunits_min_blocks(int dp) { return dp <= 6 ? 4 : 4; }

template <int DP, int STAGES>
static __global__ void __launch_bounds__(UNITS_THREADS, gunits_min_blocks(DP))
knn_gunits_kernel(const float* __restrict__ shadow, const float* __restrict__ sub_box, int64_t n_tiles,
                  int64_t q_begin, int64_t nq, int kcap, const float* __restrict__ qshadow,
                  const int32_t* __restrict__ qhome, const float* __restrict__ seed_thr,
                  const int2* __restrict__ plan, const int32_t* __restrict__ plan_count,
                  const int32_t* __restrict__ plan_first, const int32_t* __restrict__ unit_first,
                  const int32_t* __restrict__ unit_block, int32_t* __restrict__ meta, float* __restrict__ part_d32,
                  int32_t* __restrict__ part_idx, unsigned long long* __restrict__ pair_counter,
                  int32_t* __restrict__ unit_done, const __grid_constant__ RefineArgs ra) {
    static_assert(FAST_THREADS == 128 && FAST_NSUB == 16 && FAST_SUB == 32, "geometry of the lane / group mapping");
    constexpr int QB = FAST_THREADS;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    constexpr int BOX_FLOATS = FAST_NSUB * DP * 2;
    constexpr int SLOT_FLOATS = TILE_FLOATS + BOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // STAGES * SLOT_FLOATS
    float* lst_d = tiles + STAGES * SLOT_FLOATS;                             // kcap * QB
    int32_t* lst_i = reinterpret_cast<int32_t*>(lst_d + (size_t)kcap * QB);  // kcap * QB
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int s_unit;
    __shared__ int s_last;
    __shared__ int s_tile[STAGES];
    __shared__ float s_wthr[FAST_THREADS / 32];                              // largest threshold of each filter warp
    __shared__ float2 s_gbox[FAST_THREADS / 32][GU_GROUPS][DP];              // (lo, hi) of every group's queries
    __shared__ float s_gthr[FAST_THREADS / 32][GU_GROUPS];                   // largest threshold of every group

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strm = lane >> 3, mem = lane & 7;
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
    const int32_t* unit_count = unit_first + ((nq + QB - 1) / QB + 1);
    int slot = 0;
    uint32_t phase_bits = producer ? 0xffffffffu : 0u;
    auto advance = [&]() {
        phase_bits ^= 1u << slot;
        slot = slot + 1 == STAGES ? 0 : slot + 1;
    };
    unsigned int scans_total = 0;  // (group, sub-tile) scans of this warp

    for (;;) {
        if (tid == 0) s_unit = atomicAdd(&meta[2], 1);
        __syncthreads();  // (A)
        const int u = s_unit;
        if (u >= n_units) break;
        const int64_t qblock = unit_block[u];
        const int first = (u - unit_first[qblock]) * chunk;
        const int count = plan_count[qblock];
        const int last = first + chunk < count ? first + chunk : count;
        const int64_t qb0 = qblock * QB;
        const int pfirst = plan_first[qblock];
        const int ufirst = unit_first[qblock];
        const int nunits = unit_count[qblock];
        const bool refine_here = ra.fused && nunits == 1;

        // ---------------- filter warps: member `mem` of every group, thresholds, group boxes, empty lists
        float2 q[GU_GROUPS][DP / 2];
        float thr[GU_GROUPS];    // insertion threshold of query (g, mem) -- the same in all four streams
        float seed[GU_GROUPS];
        int pmax[GU_GROUPS], nfill[GU_GROUPS];  // list state (meaningful in stream 0, which inserts)
        float gthr[GU_GROUPS];   // largest threshold of group g (warp-uniform)
        if (!producer) {
#pragma unroll
            for (int g = 0; g < GU_GROUPS; ++g) {
                const int64_t qi = qb0 + warp * 32 + g * GU_GQ + mem;
                const bool in = qi < nq;
                load_row<DP>((qshadow ? qshadow : shadow) + (q_begin + (in ? qi : 0)) * DP, q[g]);
                const float s0 = in ? seed_thr[qi] : -1.0f;
                seed[g] = s0 < 0.0f ? -1.0f : s0;  // deferred / dead slots never insert
                thr[g] = seed[g];
                pmax[g] = 0;
                nfill[g] = 0;
            }
            // stream s folds group s: box and largest threshold over the members (xor 1, 2, 4 stay inside a stream)
            float2 qs[DP / 2];
            float ts = -1.0f;
#pragma unroll
            for (int g = 0; g < GU_GROUPS; ++g)
                if (g == strm) {
#pragma unroll
                    for (int h = 0; h < DP / 2; ++h) qs[h] = q[g][h];
                    ts = thr[g];
                }
            const bool in_box = ts >= 0.0f;
#pragma unroll
            for (int h = 0; h < DP / 2; ++h) {
                float lx = in_box ? qs[h].x : CUDART_INF_F, hx = in_box ? qs[h].x : -CUDART_INF_F;
                float ly = in_box ? qs[h].y : CUDART_INF_F, hy = in_box ? qs[h].y : -CUDART_INF_F;
#pragma unroll
                for (int off = 1; off < GU_GQ; off <<= 1) {
                    lx = fminf(lx, __shfl_xor_sync(FULL, lx, off)); hx = fmaxf(hx, __shfl_xor_sync(FULL, hx, off));
                    ly = fminf(ly, __shfl_xor_sync(FULL, ly, off)); hy = fmaxf(hy, __shfl_xor_sync(FULL, hy, off));
                }
                if (mem == 0) {
                    s_gbox[warp][strm][2 * h] = make_float2(lx, hx);
                    s_gbox[warp][strm][2 * h + 1] = make_float2(ly, hy);
                }
            }
#pragma unroll
            for (int off = 1; off < GU_GQ; off <<= 1) ts = fmaxf(ts, __shfl_xor_sync(FULL, ts, off));
            if (mem == 0) s_gthr[warp][strm] = ts;
            float wt = ts;
            wt = fmaxf(wt, __shfl_xor_sync(FULL, wt, 8));
            wt = fmaxf(wt, __shfl_xor_sync(FULL, wt, 16));
            if (lane == 0) *reinterpret_cast<volatile float*>(&s_wthr[warp]) = wt;
            __syncwarp();
#pragma unroll
            for (int g = 0; g < GU_GROUPS; ++g) gthr[g] = s_gthr[warp][g];
            for (int r = 0; r < kcap; ++r) {
                lst_d[(size_t)r * QB + tid] = CUDART_INF_F;
                lst_i[(size_t)r * QB + tid] = -1;
            }
        }
        __syncthreads();  // (B)

        if (producer) {
            // ---------------- copy warp: as in knn_units_kernel
            const int2* lst = plan + (pfirst < 0 ? 0 : pfirst);
            const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
            const int64_t own = (qhome ? (int64_t)qhome[mid] : mid) / FAST_TILE;
            for (int next = first; next <= last; ++next) {
                int tile = -1;
                if (next < last) {
                    if (pfirst < 0) {
                        tile = (int)walk_tile(next, own, n_tiles);
                    } else {
                        const int2 e = lst[next];
                        float R = -1.0f;
#pragma unroll
                        for (int w = 0; w < FAST_THREADS / 32; ++w)
                            R = fmaxf(R, *reinterpret_cast<volatile float*>(&s_wthr[w]));
                        if (!(__int_as_float(e.y) < R)) continue;
                        tile = e.x;
                    }
                }
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

        // ---------------- filter warps
        for (;;) {
            mbar_wait(&full_bar[slot], (phase_bits >> slot) & 1u);
            const int t = *reinterpret_cast<volatile int*>(&s_tile[slot]);
            if (t >= 0) {
                const float* tile = tiles + (size_t)slot * SLOT_FLOATS;
                const float2* boxes = reinterpret_cast<const float2*>(tile + TILE_FLOATS);
                const int64_t tile_base = (int64_t)t * FAST_TILE;
                // round r: lanes 0..15 test the 16 sub-tiles against group 2 r, lanes 16..31 against group 2 r + 1
                float my_gap[2];
                uint32_t live[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int g = 2 * r + (lane >> 4), sbl = lane & (FAST_NSUB - 1);
                    float gap = 0.0f;
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        const float2 bx = boxes[sbl * DP + d];
                        const float2 gb = s_gbox[warp][g][d];
                        gap = fmaxf(gap, fmaxf(bx.x - gb.y, gb.x - bx.y));
                    }
                    my_gap[r] = gap;
                    live[r] = __ballot_sync(FULL, gap < ((lane >> 4) ? gthr[2 * r + 1] : gthr[2 * r]));
                }
#pragma unroll
                for (int g = 0; g < GU_GROUPS; ++g) {
                    uint32_t todo = (g & 1) ? (live[g >> 1] >> 16) : (live[g >> 1] & 0xffffu);
#pragma unroll 1
                    while (todo) {
                        const int sb = __ffs(todo) - 1;
                        todo &= todo - 1u;
                        const float gap = __shfl_sync(FULL, my_gap[g >> 1], (g & 1) * 16 + sb);
                        if (!(gap < gthr[g])) continue;  // warp-uniform: the group's threshold has shrunk since
                        ++scans_total;
                        const float* rows = tile + sb * FAST_SUB * DP;
                        float mrun = CUDART_INF_F;
#pragma unroll
                        for (int i = 0; i < FAST_SUB / 4; ++i) {
                            float2 cc[DP / 2];
                            load_row<DP>(rows + (4 * i + strm) * DP, cc);
                            mrun = fminf(mrun, cheb32<DP>(q[g], cc));
                        }
                        mrun = fminf(mrun, __shfl_xor_sync(FULL, mrun, 8));
                        mrun = fminf(mrun, __shfl_xor_sync(FULL, mrun, 16));
                        const bool hit = mrun < thr[g];
                        if (!__any_sync(FULL, hit)) continue;
                        bool changed = false;
                        if (hit && strm == 0) {
                            const int col = warp * 32 + g * GU_GQ + mem;
#pragma unroll 1
                            for (int c = 0; c < FAST_SUB; ++c) {
                                float2 cc[DP / 2];
                                load_row<DP>(rows + c * DP, cc);
                                const float m = cheb32<DP>(q[g], cc);
                                if (m < thr[g]) {
                                    lst_d[(size_t)pmax[g] * QB + col] = m;
                                    lst_i[(size_t)pmax[g] * QB + col] = (int32_t)(tile_base + sb * FAST_SUB + c);
                                    if (nfill[g] + 1 < kcap) {
                                        pmax[g] = ++nfill[g];  // free slots left: append, the threshold stays at the seed
                                    } else {
                                        nfill[g] = kcap;
                                        float mx = -1.0f;
                                        int p = 0;
#pragma unroll 3
                                        for (int r = 0; r < kcap; ++r) {
                                            const float v = lst_d[(size_t)r * QB + col];
                                            if (v > mx) { mx = v; p = r; }
                                        }
                                        pmax[g] = p;
                                        thr[g] = fminf(seed[g], mx);
                                        changed = true;
                                    }
                                }
                            }
                        }
                        if (__any_sync(FULL, changed)) {
                            thr[g] = __shfl_sync(FULL, thr[g], mem);  // stream 0 holds the new thresholds
                            float gt = thr[g];
#pragma unroll
                            for (int off = 1; off < GU_GQ; off <<= 1) gt = fmaxf(gt, __shfl_xor_sync(FULL, gt, off));
                            gthr[g] = gt;
                            float wt = fmaxf(fmaxf(gthr[0], gthr[1]), fmaxf(gthr[2], gthr[3]));
                            if (lane == 0) *reinterpret_cast<volatile float*>(&s_wthr[warp]) = wt;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            advance();
            if (t < 0) break;
        }

        // ---------------- lane l now owns the list of query warp * 32 + l (filled by the stream-0 lane of its group)
        __syncwarp();
        int filled = 0;
#pragma unroll
        for (int g = 0; g < GU_GROUPS; ++g) {
            const int v = __shfl_sync(FULL, nfill[g], lane & (GU_GQ - 1));
            if (g == (lane >> 3)) filled = v;
        }
        const int lq = warp * 32 + lane;
        const int64_t qi_l = qb0 + lq;
        const bool alive = qi_l < nq;
        const float s_l = alive ? seed_thr[qi_l] : -1.0f;
        const float seed_l = s_l < 0.0f ? -1.0f : s_l;
        {
            const int col = tid;
            for (int a = 1; a < filled; ++a) {  // ascending; the rest of the column is +inf / -1
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
            if (refine_here) {
                if (alive) refine_single_list(lst_d + col, lst_i + col, QB, kcap, seed_l, q_begin + qi_l, qi_l, ra);
            } else {
                for (int r = 0; r < kcap; ++r) {
                    const int64_t o = ((int64_t)u * kcap + r) * QB + lq;
                    part_d32[o] = lst_d[(size_t)r * QB + col];
                    part_idx[o] = lst_i[(size_t)r * QB + col];
                }
            }
        }
        if (ra.fused && !refine_here) {
            __threadfence();
            filter_warps_barrier();
            if (tid == 0) s_last = atomicAdd(&unit_done[qblock], 1) == nunits - 1;
            filter_warps_barrier();
            if (s_last) {
                __threadfence();
                if (alive)
                    refine_merge_lists(part_d32, part_idx, kcap, ufirst, ufirst + nunits, QB, lq, true,
                                       seed_l < 0.0f ? -1.0f : seed_thr[qi_l], q_begin + qi_l, qi_l, ra);
            }
        }
    }
    if (!producer && lane == 0 && pair_counter && scans_total)
        atomicAdd(pair_counter, (unsigned long long)scans_total * (unsigned long long)(FAST_SUB * GU_GQ));
}

}  // namespace ksg

// Windowed k-NN filter for many dimensions (DP >= 6): 8-query groups inside the filter warps.
//
// In 6 to 8 coordinates the work of knn_units_kernel is set by the QUERY GROUP a sub-tile is culled
// against -- the box of the warp's 32 queries inflated by the LARGEST of their thresholds -- not by the
// candidate boxes: even with the true radii a 32-query warp must scan 14 % (6-D) / 43 % (8-D) of all
// pairs at n = 2e5, a group of 8 consecutive queries 5 % / 19 % (tools/pruning_sim.py,
// profiles/r1aj_pruning_simulation.txt).  Here every filter warp keeps its 32 consecutive queries as
// FOUR groups of 8, each with its own box, threshold and surviving sub-tiles.  Lane l owns query l (its
// list, its threshold) and is (stream s = l / 8, member m = l % 8): the owners of group g are the lanes of
// stream g.  A surviving (group, sub-tile) is scanned by all 32 lanes at once: every lane fetches member
// m's query of that group from its owner (shuffles, once per group and tile), stream s takes the candidates
// c = 4 i + s (four neighbouring rows per shared-memory read, conflict-free) -- 8 steps instead of 32 --
// and keeps one hit bit per candidate; only when somebody saw a candidate below its threshold do the
// owners collect the four streams' bits of their query (four shuffles) and insert exactly those candidates.
// Same tile ring, copy warp, planning pass, unit table, lists and refine as knn_units_kernel; results are
// identical.
#pragma once

#include "fast_knn.cuh"

namespace ksg {

constexpr int GU_GROUPS = 4;  // query groups per filter warp
constexpr int GU_GQ = 8;      // queries per group
constexpr int gunits_min_blocks(int dp) { return dp <= 6 ? 6 : 5; }

template <int DP, int STAGES>
static __global__ void __launch_bounds__(UNITS_THREADS, gunits_min_blocks(DP))
knn_gunits_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box,
                  const float* __restrict__ sub_box, int64_t n_tiles,
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

        // ---------------- filter warps: lane l owns query warp * 32 + l (group l / 8, member l % 8)
        float2 q[DP / 2];
        float thr = -1.0f, seed = -1.0f;  // insertion threshold: min(seed, current list maximum)
        int pmax = 0, nfill = 0;
        bool alive = false;
        float gthr[GU_GROUPS];  // largest threshold of every group (warp-uniform)
        if (!producer) {
            const int64_t qi = qb0 + warp * 32 + lane;
            alive = qi < nq;
            load_row<DP>((qshadow ? qshadow : shadow) + (q_begin + (alive ? qi : 0)) * DP, q);
            const float s0 = alive ? seed_thr[qi] : -1.0f;
            seed = s0 < 0.0f ? -1.0f : s0;  // deferred / dead slots never insert
            thr = seed;
            // a group's members are the 8 lanes of one stream: xor 1, 2, 4 fold inside it
            const bool in_box = thr >= 0.0f;
#pragma unroll
            for (int h = 0; h < DP / 2; ++h) {
                float lx = in_box ? q[h].x : CUDART_INF_F, hx = in_box ? q[h].x : -CUDART_INF_F;
                float ly = in_box ? q[h].y : CUDART_INF_F, hy = in_box ? q[h].y : -CUDART_INF_F;
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
            float ts = thr;
#pragma unroll
            for (int off = 1; off < GU_GQ; off <<= 1) ts = fmaxf(ts, __shfl_xor_sync(FULL, ts, off));
            if (mem == 0) *reinterpret_cast<volatile float*>(&s_gthr[warp][strm]) = ts;
#pragma unroll
            for (int g = 0; g < GU_GROUPS; ++g) gthr[g] = __shfl_sync(FULL, ts, g * GU_GQ);
            if (lane == 0)
                *reinterpret_cast<volatile float*>(&s_wthr[warp]) = fmaxf(fmaxf(gthr[0], gthr[1]), fmaxf(gthr[2], gthr[3]));
            __syncwarp();
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
                    // Does ANY of the CTA's 16 query groups reach into the tile?  (lane l < 16: group l % 4 of
                    // filter warp l / 4; thresholds only shrink, a stale one is merely conservative.)  A tile
                    // nobody needs is neither copied nor waited for: with 8-query groups that is most of a
                    // planned list, and the tile feed -- not the arithmetic -- was what the filter warps waited on.
                    bool reach = false;
                    if (lane < (FAST_THREADS / 32) * GU_GROUPS) {
                        const int w = lane >> 2, g = lane & 3;
                        float gap = 0.0f;
#pragma unroll
                        for (int d = 0; d < DP; ++d) {
                            const float2 bx = reinterpret_cast<const float2*>(tile_box)[(int64_t)tile * DP + d];
                            const float2 gb = s_gbox[w][g][d];
                            gap = fmaxf(gap, fmaxf(bx.x - gb.y, gb.x - bx.y));
                        }
                        reach = gap < *reinterpret_cast<volatile float*>(&s_gthr[w][g]);
                    }
                    if (!__any_sync(FULL, reach)) continue;
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
        auto insert = [&](float m, int32_t idx) -> bool {
            lst_d[(size_t)pmax * QB + tid] = m;
            lst_i[(size_t)pmax * QB + tid] = idx;
            if (nfill + 1 < kcap) {  // free slots left: append, the threshold stays at the seed
                pmax = ++nfill;
                return false;
            }
            nfill = kcap;
            float mx = -1.0f;
            int p = 0;
#pragma unroll 3
            for (int r = 0; r < kcap; ++r) {
                const float v = lst_d[(size_t)r * QB + tid];
                if (v > mx) { mx = v; p = r; }
            }
            pmax = p;
            thr = fminf(seed, mx);
            return true;  // the threshold may have shrunk
        };
        for (;;) {
            mbar_wait(&full_bar[slot], (phase_bits >> slot) & 1u);
            const int t = *reinterpret_cast<volatile int*>(&s_tile[slot]);
            if (t >= 0) {
                const float* tile = tiles + (size_t)slot * SLOT_FLOATS;
                const float2* boxes = reinterpret_cast<const float2*>(tile + TILE_FLOATS);
                const int32_t tile_base = (int32_t)((int64_t)t * FAST_TILE);
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
                    if (!todo) continue;  // warp-uniform
                    // this lane scans for member `mem` of group g: its query and threshold come from their owner
                    float2 qg[DP / 2];
#pragma unroll
                    for (int h = 0; h < DP / 2; ++h) {
                        qg[h].x = __shfl_sync(FULL, q[h].x, g * GU_GQ + mem);
                        qg[h].y = __shfl_sync(FULL, q[h].y, g * GU_GQ + mem);
                    }
                    float tg = __shfl_sync(FULL, thr, g * GU_GQ + mem);
#pragma unroll 1
                    while (todo) {
                        const int sb = __ffs(todo) - 1;
                        todo &= todo - 1u;
                        const float gap = __shfl_sync(FULL, my_gap[g >> 1], (g & 1) * 16 + sb);
                        if (!(gap < gthr[g])) continue;  // warp-uniform: the group's threshold has shrunk since
                        ++scans_total;
                        const float* rows = tile + sb * FAST_SUB * DP;
                        // stream s takes the candidates 4 i + s; bit 7 - i of hm: candidate 4 i + s is closer than tg
                        uint32_t hm = 0u;
#pragma unroll
                        for (int i = 0; i < FAST_SUB / 4; ++i) {
                            float2 cc[DP / 2];
                            load_row<DP>(rows + (4 * i + strm) * DP, cc);
                            hm = push_sign(hm, cheb32<DP>(qg, cc) - tg);
                        }
                        if (!__any_sync(FULL, hm != 0u)) continue;
                        // the owners (stream g) collect the four streams' bits of their query and insert
                        uint32_t hs[4];
#pragma unroll
                        for (int s2 = 0; s2 < 4; ++s2) hs[s2] = __shfl_sync(FULL, hm, s2 * GU_GQ + mem);
                        bool changed = false;
                        if (strm == g) {
#pragma unroll
                            for (int s2 = 0; s2 < 4; ++s2) {
                                uint32_t mk = hs[s2];
                                while (mk) {
                                    const int b = 31 - __clz(mk);
                                    mk &= ~(1u << b);
                                    const int c = 4 * (FAST_SUB / 4 - 1 - b) + s2;
                                    float2 cc[DP / 2];
                                    load_row<DP>(rows + c * DP, cc);
                                    const float m = cheb32<DP>(q, cc);
                                    if (m < thr) changed |= insert(m, tile_base + sb * FAST_SUB + c);
                                }
                            }
                        }
                        if (__any_sync(FULL, changed)) {
                            tg = __shfl_sync(FULL, thr, g * GU_GQ + mem);
                            float ts = thr;
#pragma unroll
                            for (int off = 1; off < GU_GQ; off <<= 1) ts = fmaxf(ts, __shfl_xor_sync(FULL, ts, off));
                            gthr[g] = __shfl_sync(FULL, ts, g * GU_GQ);
                            if (lane == 0) {
                                *reinterpret_cast<volatile float*>(&s_gthr[warp][g]) = gthr[g];
                                *reinterpret_cast<volatile float*>(&s_wthr[warp]) =
                                    fmaxf(fmaxf(gthr[0], gthr[1]), fmaxf(gthr[2], gthr[3]));
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            advance();
            if (t < 0) break;
        }

        // ---------------- sort each list (ascending); refine it here or emit it
        const int lq = warp * 32 + lane;
        const int64_t qi_l = qb0 + lq;
        const float seed_l = seed;
        {
            const int col = tid;
            const int filled = nfill;  // the rest of the column is +inf / -1
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
            if (tid == 0) s_last = atomicAdd(&unit_done[4 * qblock], 1) == nunits - 1;
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

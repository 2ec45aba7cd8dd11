// Windowed k-NN filter for few dimensions (DP <= 4): every WARP is its own worker.
//
// knn_units_kernel (fast_knn.cuh) couples the four filter warps of a CTA through a shared tile ring:
// they meet at a block barrier per work unit, spin on the ring's mbarriers and wait for whole
// 512-row tiles of which a warp scans one or two 32-row sub-tiles.  ncu on C2 (2-D, n = 1e6,
// profiles/r2g_c2_knn_units_source_lines.txt): 15 % of the stall samples in the per-unit barrier,
// 10 % + 5 % waiting for a tile, 9 % of all instructions spinning on the full barrier.  In 2-D / 4-D
// a warp needs ~10 sub-tiles of ~6 planned tiles, i.e. 2.5 KB of the 24 KB a unit stages.
//
// Here a work item is (unit, quarter): the 32 consecutive queries `quarter` of the unit's query
// block, against the unit's range of the block's planned tile list.  A warp
//   1. pulls an item from the global cursor (no CTA-level hand-over at all: no __syncthreads in the loop),
//   2. box-tests the sub-tiles of EIGHT planned tiles at once (lane = (tile parity, sub-tile), four
//      independent 16-byte box loads in flight per lane) and compacts the survivors, near tiles
//      first, into a small queue in its own shared memory,
//   3. walks the queue: every lane fetches ONE row of the sub-tile (a coalesced 256 / 512-byte read
//      that hits L1 / L2: neighbouring warps scan the same sub-tiles), the next sub-tile's rows are
//      already in flight while the current one is scanned from shared memory (LDS.128 broadcasts),
//   4. keeps the per-query lists exactly as knn_units_kernel does and ends the same way: single-unit
//      blocks are refined from shared memory, units of longer lists emit partial lists and the warp
//      that finishes the LAST item of a (block, quarter) merges that quarter's 32 queries (unit_done: four
//      counters per block).
// Same planning pass, unit table, list format and refine functions; results are identical.
// Sets of 5 to 8 coordinates keep knn_units_kernel: there a warp scans a large share of every staged
// tile and four warps sharing one staged copy is what keeps the L2 traffic down.
#pragma once

#include "fast_knn.cuh"

namespace ksg {

constexpr int WU_GROUP = 8;                     // planned tiles box-tested together
constexpr int WU_QUEUE = WU_GROUP * FAST_NSUB;  // survivor queue entries per warp
constexpr int WU_MAX_KC = 24;                   // longest list of this kernel (longer ones: knn_units_kernel)
constexpr int WU_BAND_REGS = 4;                 // in-band entries the unsorted refine resolves in registers
constexpr int wunits_min_blocks(int dp) { return dp <= 2 ? 10 : 8; }
__host__ __device__ inline size_t wunits_smem_per_warp(int kc, int dp) {
    return (size_t)kc * 32 * 8 + (size_t)WU_QUEUE * 8 + (size_t)32 * dp * 4;
}
// (env KSG_B200_WU_ROOMY=1, read by the launcher: the next capacity up -- fewer full lists, i.e. fewer
//  replace-the-maximum sweeps, for 1 KB more shared memory per warp)
constexpr int wunits_list_capacity(int kcap) { return kcap <= 8 ? 8 : kcap <= 12 ? 12 : kcap <= 16 ? 16 : WU_MAX_KC; }

// The refine of ONE list that still sits UNSORTED in shared memory (entries [0, nfill) of the lane's
// column; radius only).  Same decision rule as refine_single_list, without the sort in front of it:
//   pass 1  r32 = the k'-th smallest entry, from a branch-free min / max insertion network over KPR
//           registers (every lane does the same work: no divergence, no dependent shared-memory traffic);
//   pass 2  a = #entries below the band [r32 - 2 delta, r32 + 2 delta], the entries inside it as a bit
//           mask (float compares against the bounds rounded towards the band are the float64 compares);
//   then the (k' - a)-th smallest EXACT distance among the <= WU_BAND_REGS in-band entries.
// Returns false when the case needs the general code (undecided, crowded band): the caller then sorts the
// list and calls refine_single_list, which takes the same decisions on the same data.
template <int DP, int KC, int KPR>
static __device__ __noinline__ bool refine_unsorted_list(const float* col_d, const int32_t* col_i, int nfill, float seed,
                                                         int64_t row_q, int64_t qi, bool exact32,
                                                         const RefineArgs& ra) {
    const int kp = ra.kp, dim = ra.dim;
    float dl[KPR];
#pragma unroll
    for (int t = 0; t < KPR; ++t) dl[t] = CUDART_INF_F;
#pragma unroll
    for (int r = 0; r < KC; ++r) {
        const float v = r < nfill ? col_d[r * 32] : CUDART_INF_F;
#pragma unroll
        for (int t = KPR - 1; t > 0; --t) dl[t] = fminf(dl[t], fmaxf(dl[t - 1], v));
        dl[0] = fminf(dl[0], v);
    }
    float v_kp = CUDART_INF_F;
#pragma unroll
    for (int t = 0; t < KPR; ++t)
        if (t == kp - 1) v_kp = dl[t];
    if (!(v_kp < CUDART_INF_F)) {  // fewer than k' points exist (k+1 > n): cKDTree pads with inf
        ra.eps_out[qi] = CUDART_INF;
        ra.flags[qi] = 0;
        return true;
    }
    if (exact32) {  // quantised / integer data: the fp32 distances are the float64 distances
        ra.flags[qi] = 0;
        ra.eps_out[qi] = (double)v_kp;
        return true;
    }
    double qc[DP];
    double mq = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d)
        if (d < dim) {
            qc[d] = ra.qsorted ? ra.qsorted[(int64_t)d * ra.qld + row_q] : ra.sorted[(int64_t)d * ra.ld + row_q];
            mq = fmax(mq, fabs(qc[d] - ra.center[d]));
        }
    mq = mq * (1.0 + 1e-6) + 1e-300;
    const double r32 = (double)v_kp;
    const double band2 = 2.0 * band_query_f64(mq, 2.0 * r32);
    const double T = r32 + band2, lo = r32 - band2;
    if (!(T < (double)seed)) return false;  // the band reaches the seeded threshold: in-band candidates may be missing
    // float v: (double)v < lo  <=>  v < ru(lo);   (double)v <= T  <=>  v <= rd(T)
    const float lo_f = __double2float_ru(lo), T_f = __double2float_rd(T);
    int a = 0;
    unsigned inband = 0u;
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int r = 0; r < KC; ++r) {
        const float v = r < nfill ? col_d[r * 32] : CUDART_INF_F;
        a += v < lo_f ? 1 : 0;
        inband |= (!(v < lo_f) && v <= T_f) ? (1u << r) : 0u;
        mx = r < nfill ? fmaxf(mx, v) : mx;
    }
    if (nfill >= KC && (double)mx <= T) return false;  // full list with its tail inside the band
    const int b = __popc(inband), want = kp - a;  // 1 <= want <= b
    if (b > WU_BAND_REGS) return false;
    double mb[WU_BAND_REGS];
#pragma unroll
    for (int t = 0; t < WU_BAND_REGS; ++t) {
        mb[t] = CUDART_INF;
        if (t < b) {
            const int r = __ffs(inband) - 1;
            inband &= inband - 1u;
            const int64_t j = col_i[r * 32];
            double v = 0.0;
#pragma unroll
            for (int d = 0; d < DP; ++d)
                if (d < dim) v = fmax(v, fabs(qc[d] - ra.sorted[(int64_t)d * ra.ld + j]));
            mb[t] = v;
        }
    }
    // the want-th smallest of mb[0..b): the value with fewer than `want` values below it and at least
    // `want` values not above it
    double ans = mb[0];
#pragma unroll
    for (int r = 0; r < WU_BAND_REGS; ++r) {
        int less = 0, le = 0;
#pragma unroll
        for (int t = 0; t < WU_BAND_REGS; ++t) {
            less += (t < b && mb[t] < mb[r]) ? 1 : 0;
            le += (t < b && mb[t] <= mb[r]) ? 1 : 0;
        }
        if (r < b && less < want && want <= le) ans = mb[r];
    }
    ra.flags[qi] = 0;
    ra.eps_out[qi] = ans;
    return true;
}

template <int DP, int KC>  // KC: list capacity (>= kcap, the length of an EMITTED partial list)
static __global__ void __launch_bounds__(FAST_THREADS, wunits_min_blocks(DP))
knn_wunits_kernel(const float* __restrict__ shadow, const float* __restrict__ sub_box, int64_t n_tiles,
                  int64_t q_begin, int64_t nq, int kcap, const float* __restrict__ qshadow,
                  const int32_t* __restrict__ qhome, const float* __restrict__ seed_thr,
                  const int2* __restrict__ plan, const int32_t* __restrict__ plan_count,
                  const int32_t* __restrict__ plan_first, const int32_t* __restrict__ unit_first,
                  const int32_t* __restrict__ unit_block, int32_t* __restrict__ meta, float* __restrict__ part_d32,
                  int32_t* __restrict__ part_idx, unsigned long long* __restrict__ pair_counter,
                  int32_t* __restrict__ unit_done, const __grid_constant__ RefineArgs ra) {
    static_assert(FAST_NSUB == 16 && FAST_SUB == 32, "lane = (tile parity, sub-tile); one row per lane");
    constexpr int QB = FAST_THREADS;  // queries of a planning block (one per thread of the planning pass)
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem_raw + (size_t)warp * wunits_smem_per_warp(KC, DP);
    float* col_d = reinterpret_cast<float*>(mine) + lane;                       // this lane's list: entry r at [r * 32]
    int32_t* col_i = reinterpret_cast<int32_t*>(mine + (size_t)KC * 32 * 4) + lane;
    int2* queue = reinterpret_cast<int2*>(mine + (size_t)KC * 32 * 8);          // (tile * 16 + sub-tile, gap bits)
    float* rows = reinterpret_cast<float*>(queue + WU_QUEUE);                   // one staged sub-tile: [32][DP]
    // kernel-wide facts of the refine
    const bool range_ok = fp32_range_ok(ra.maxabs, ra.dim);
    const bool exact32 = ra.exact32 && *ra.exact32;
    const bool fast_refine = ra.fused && !ra.idx_out && range_ok && ra.kp <= 8;

    const int n_units = meta[0], chunk = meta[1];
    const int32_t* unit_count = unit_first + ((nq + QB - 1) / QB + 1);  // (see knn_unit_table_kernel)
    const long long n_items = 4ll * n_units;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned int subtiles_total = 0;  // sub-tile scans of this warp

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&meta[2], 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= n_items) break;
        const int u = w >> 2, quarter = w & 3;
        const int64_t qblock = unit_block[u];
        const int ufirst = unit_first[qblock];
        const int nunits = unit_count[qblock];
        const int count = plan_count[qblock];
        const int pfirst = plan_first[qblock];
        const int first = (u - ufirst) * chunk;
        const int last = first + chunk < count ? first + chunk : count;
        const int64_t qb0 = qblock * QB;
        const bool refine_here = ra.fused && nunits == 1;
        const int lq = quarter * 32 + lane;
        const int64_t qi = qb0 + lq;
        const bool alive = qi < nq;

        // ---------------- the warp's queries, thresholds, box and empty lists
        float2 q[DP / 2];
        load_row<DP>((qshadow ? qshadow : shadow) + (q_begin + (alive ? qi : 0)) * DP, q);
        const float s0 = alive ? seed_thr[qi] : -1.0f;
        const float seed = s0 < 0.0f ? -1.0f : s0;  // deferred / dead lanes never insert
        float thr = seed;  // insertion threshold: min(seed, current list maximum)
        int pmax = 0;      // position of the entry a hit replaces
        int nfill = 0;     // entries written so far (kcap: the list is full)
        float wlo[DP], whi[DP];
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            const bool in = thr >= 0.0f;
            wlo[2 * h] = in ? q[h].x : CUDART_INF_F;      whi[2 * h] = in ? q[h].x : -CUDART_INF_F;
            wlo[2 * h + 1] = in ? q[h].y : CUDART_INF_F;  whi[2 * h + 1] = in ? q[h].y : -CUDART_INF_F;
        }
#pragma unroll
        for (int d = 0; d < DP; ++d)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                wlo[d] = fminf(wlo[d], __shfl_xor_sync(FULL, wlo[d], off));
                whi[d] = fmaxf(whi[d], __shfl_xor_sync(FULL, whi[d], off));
            }
        // (the lists are not cleared: entries [nfill, KC) are never read before they are written or padded)
        float wthr;  // the warp's largest threshold
        auto warp_refresh = [&]() {
            float v = thr;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, off));
            wthr = v;
        };
        warp_refresh();

        auto insert = [&](float m, int32_t idx) -> bool {
            col_d[pmax * 32] = m;
            col_i[pmax * 32] = idx;
            // while the list still has free slots the entry to replace is simply the next free one and
            // the threshold stays at the seed: no sweep
            if (nfill + 1 < KC) {
                pmax = ++nfill;
                return false;
            }
            nfill = KC;
            float mx = col_d[0];
            int p = 0;
#pragma unroll
            for (int r = 1; r < KC; ++r) {
                const float v = col_d[r * 32];
                if (v > mx) { mx = v; p = r; }
            }
            pmax = p;
            thr = fminf(seed, mx);
            return true;  // the threshold may have shrunk
        };

        // ---------------- the unit's tiles, WU_GROUP at a time
        const int2* lst = plan + (pfirst < 0 ? 0 : pfirst);
        int64_t own = 0;
        if (pfirst < 0) {  // implicit list "every tile, in walk order" (as in the planning pass)
            const int64_t mid = q_begin + min(qb0 + QB / 2, nq - 1);
            own = (qhome ? (int64_t)qhome[mid] : mid) / FAST_TILE;
        }
#pragma unroll 1
        for (int g = first; g < last; g += WU_GROUP) {
            // lane (h, s) of round k tests sub-tile s of the list's entry g + 2 k + h
            float gp[WU_GROUP / 2];
            int id[WU_GROUP / 2];
#pragma unroll
            for (int k = 0; k < WU_GROUP / 2; ++k) {
                const int e = g + 2 * k + (lane >> 4);
                gp[k] = CUDART_INF_F;
                id[k] = 0;
                if (e < last) {
                    int t;
                    float tg;
                    if (pfirst < 0) {
                        t = (int)walk_tile(e, own, n_tiles);
                        tg = 0.0f;
                    } else {
                        const int2 en = lst[e];
                        t = en.x;
                        tg = __int_as_float(en.y);
                    }
                    if (tg < wthr) {  // (the block's gap is a lower bound of the warp's)
                        const int sid = t * FAST_NSUB + (lane & (FAST_NSUB - 1));
                        const float2* bx = reinterpret_cast<const float2*>(sub_box) + (int64_t)sid * DP;
                        float gg = 0.0f;
                        if constexpr (DP == 2) {
                            const float4 b = *reinterpret_cast<const float4*>(bx);
                            gg = fmaxf(fmaxf(b.x - whi[0], wlo[0] - b.y), fmaxf(b.z - whi[1], wlo[1] - b.w));
                            gg = fmaxf(gg, 0.0f);
                        } else {
#pragma unroll
                            for (int d = 0; d < DP; ++d) {
                                const float2 b = bx[d];
                                gg = fmaxf(gg, fmaxf(b.x - whi[d], wlo[d] - b.y));
                            }
                        }
                        gp[k] = gg;
                        id[k] = sid;
                    }
                }
            }
            int n_live = 0;
#pragma unroll
            for (int k = 0; k < WU_GROUP / 2; ++k) {
                // insertion needs m < thr and the box gap is a lower bound of every computed m
                const bool keep = gp[k] < wthr;
                const unsigned b = __ballot_sync(FULL, keep);
                if (keep) queue[n_live + __popc(b & lt_mask)] = make_int2(id[k], __float_as_int(gp[k]));
                n_live += __popc(b);
            }
            __syncwarp();
            if (n_live == 0) continue;

            // walk the survivors; the next one's rows are in flight while this one is scanned
            int2 ent = queue[0];
            float2 pre[DP / 2];
            load_row<DP>(shadow + ((int64_t)ent.x * FAST_SUB + lane) * DP, pre);
#pragma unroll 1
            for (int i = 0; i < n_live; ++i) {
                const int2 cur = ent;
                float2 mine_row[DP / 2];
#pragma unroll
                for (int h = 0; h < DP / 2; ++h) mine_row[h] = pre[h];
                if (i + 1 < n_live) {
                    ent = queue[i + 1];
                    load_row<DP>(shadow + ((int64_t)ent.x * FAST_SUB + lane) * DP, pre);
                }
                if (!(__int_as_float(cur.y) < wthr)) continue;  // warp-uniform: the threshold has shrunk since
                ++subtiles_total;
                __syncwarp();  // everybody is done with the previous sub-tile
#pragma unroll
                for (int h = 0; h < DP / 2; ++h) reinterpret_cast<float2*>(rows)[lane * (DP / 2) + h] = mine_row[h];
                __syncwarp();
                const int32_t row0 = cur.x * FAST_SUB;  // first row of the sub-tile in the prepared order
                uint32_t mask = 0u;
#pragma unroll 8
                for (int c = 0; c < FAST_SUB; c += 2) {
                    float2 ca[DP / 2], cb[DP / 2];
                    load_two_rows<DP>(rows + c * DP, ca, cb);
                    mask = push_sign(mask, cheb32<DP>(q, ca) - thr);
                    mask = push_sign(mask, cheb32<DP>(q, cb) - thr);
                }
                // candidate c of the sub-tile sits at bit 31 - c
                bool changed = false;
                while (mask) {
                    const int c = __clz(mask);
                    mask &= ~(0x80000000u >> c);
                    float2 cc[DP / 2];
                    load_row<DP>(rows + c * DP, cc);
                    const float m = cheb32<DP>(q, cc);
                    if (m < thr) changed |= insert(m, row0 + c);
                }
                if (__any_sync(FULL, changed)) warp_refresh();
            }
            __syncwarp();  // the queue may be overwritten
        }

        // ---------------- refine the list here (single-unit block) or emit it
        if (refine_here && alive && fast_refine && seed >= 0.0f) {
            bool done;
            if (ra.kp <= 4) done = refine_unsorted_list<DP, KC, 4>(col_d, col_i, nfill, seed, q_begin + qi, qi, exact32, ra);
            else if (ra.kp <= 6) done = refine_unsorted_list<DP, KC, 6>(col_d, col_i, nfill, seed, q_begin + qi, qi, exact32, ra);
            else done = refine_unsorted_list<DP, KC, 8>(col_d, col_i, nfill, seed, q_begin + qi, qi, exact32, ra);
            if (done) continue;
        } else if (refine_here && !alive) {
            continue;
        }
        // general path: ascending order, unused entries +inf / -1
        for (int a = 1; a < nfill; ++a) {
            const float d = col_d[a * 32];
            const int32_t i = col_i[a * 32];
            int b = a - 1;
            while (b >= 0 && col_d[b * 32] > d) {
                col_d[(b + 1) * 32] = col_d[b * 32];
                col_i[(b + 1) * 32] = col_i[b * 32];
                --b;
            }
            col_d[(b + 1) * 32] = d;
            col_i[(b + 1) * 32] = i;
        }
        for (int r = nfill; r < KC; ++r) {
            col_d[r * 32] = CUDART_INF_F;
            col_i[r * 32] = -1;
        }
        if (refine_here) {
            refine_single_list(col_d, col_i, 32, KC, seed, q_begin + qi, qi, ra);
            continue;
        }
        for (int r = 0; r < kcap; ++r) {  // (kcap <= KC: a longer list is cut, its tail then says "full")
            const int64_t o = ((int64_t)u * kcap + r) * QB + lq;
            part_d32[o] = col_d[r * 32];
            part_idx[o] = col_i[r * 32];
        }
        if (ra.fused) {
            // last-item-done merge, PER QUARTER: the lists above are published (fence), the counter of this
            // (block, quarter) is bumped once per item, and the warp that brings it to nunits merges the 32
            // queries of the quarter -- the four merges of a heavy block run side by side instead of one warp
            // walking all 128 queries after the block's very last item (that merge alone was 60-80 us of a
            // 221 us launch when a rank owns 1/8 of C2's queries: per-item trace)
            __threadfence();
            __syncwarp();
            int is_last = 0;
            if (lane == 0) is_last = atomicAdd(&unit_done[4 * qblock + quarter], 1) == nunits - 1;
            is_last = __shfl_sync(FULL, is_last, 0);
            if (is_last) {
                __threadfence();
                if (alive)
                    refine_merge_lists(part_d32, part_idx, kcap, ufirst, ufirst + nunits, QB, lq, true, seed_thr[qi],
                                       q_begin + qi, qi, ra);
            }
        }
    }
    if (lane == 0 && pair_counter && subtiles_total)
        atomicAdd(pair_counter, (unsigned long long)subtiles_total * (unsigned long long)(FAST_SUB * 32));
}

}  // namespace ksg

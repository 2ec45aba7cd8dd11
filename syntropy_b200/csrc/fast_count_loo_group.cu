// Fused leave-one-out count pass (DTC, O-, S-information) with 8-QUERY GROUPS inside the filter warps.
//
// count_f32_kernel<SpecLoo<D>> (fast_count.cuh) culls candidate tiles against the box of a CTA's 256 queries; in
// 8-D a block of queries reaches every tile, so C4 evaluates all n^2 pairs (74.7 ms of its 100 ms, at 0.81 of the
// FP32 roofline: nothing left to gain per pair).  A leave-one-out subspace keeps all coordinates but one, so a
// candidate can only count when AT MOST ONE coordinate is outside the ball: a (query group, sub-tile) pair
// whose SECOND-LARGEST box gap is not below the group's largest threshold is settled without a scan.  Here
//   * a CTA owns 128 consecutive queries (prepared order), 4 filter warps + 1 copy warp; the copy warp tests
//     the tile boxes against the block's box (32 tiles per round, one per lane) and streams the surviving tiles
//     + their sub-tile boxes through a bulk-copy ring;
//   * a filter warp keeps its 32 queries as four groups of 8: lane l = (stream s = l / 8, member m = l % 8)
//     holds member m's query of every group in registers, that query's accumulators in shared memory; the 16 sub-tiles x 4 groups of a
//     tile are classified by the 32 lanes at once (two per lane, two ballots);
//   * a surviving (group, sub-tile) is scanned by all lanes -- stream s takes the candidates 4 i + s -- with the
//     indicator-sum arithmetic of SpecLoo (count_s += [sigma >= DP - 1] - [sigma == DP - 1] i_s): the same
//     saturating-FMA predicate per coordinate, hence the same counts as count_f32_kernel;
//   * at the end two shuffles per (group, subspace) add the four streams' shares; the owner lane writes.
// Counts are float accumulators of small integers (exact below 2^24 per stream: n < 6.7e7).
#include <stdlib.h>

#include "fast_count_launch.cuh"

namespace ksg {

template <int DD, int STAGES>
static __global__ void __launch_bounds__(UNITS_THREADS, 4)
count_loo_group_kernel(const float* __restrict__ shadow, const float* __restrict__ tile_box,
                       const float* __restrict__ sub_box, int64_t n_tiles, int64_t q_begin, int64_t nq,
                       const float* __restrict__ thresholds, int32_t* __restrict__ counts,
                       unsigned long long* __restrict__ pair_counter) {
    static_assert(FAST_THREADS == 128 && FAST_NSUB == 16 && FAST_SUB == 32, "geometry of the lane / group mapping");
    constexpr int DP = (DD + 1) & ~1, H = DP / 2, NG = 4, GQ = 8;
    constexpr int QB = FAST_THREADS;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    constexpr int BOX_FLOATS = FAST_NSUB * DP * 2;
    constexpr int SLOT_FLOATS = TILE_FLOATS + BOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);  // STAGES * SLOT_FLOATS
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int s_tile[STAGES];
    __shared__ float2 s_gbox[FAST_THREADS / 32][NG][DP];  // (lo, hi) of every group's live queries
    __shared__ float s_gthr[FAST_THREADS / 32][NG];       // largest threshold of the group
    // the accumulators of the three groups a lane is NOT scanning live here ([group][word][thread]: conflict-free);
    // in registers they cost 40 per thread and a fourth CTA per SM
    __shared__ float2 s_acc[NG][H][FAST_THREADS];
    __shared__ float s_ge[NG][FAST_THREADS];  // sum of [sigma >= DP - 1] = [all inside] + [exactly one outside]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strm = lane / GQ, mem = lane % GQ;
    const bool producer = warp == FAST_THREADS / 32;
    const int64_t qb0 = (int64_t)blockIdx.x * QB;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], FAST_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int slot = 0;
    uint32_t phase_bits = producer ? 0xffffffffu : 0u;
    auto advance = [&]() {
        phase_bits ^= 1u << slot;
        slot = slot + 1 == STAGES ? 0 : slot + 1;
    };

    if (producer) {
        // ---- the block's query box and largest threshold (four queries per lane)
        float blo[DP], bhi[DP], tmax = -CUDART_INF_F;
#pragma unroll
        for (int d = 0; d < DP; ++d) { blo[d] = CUDART_INF_F; bhi[d] = -CUDART_INF_F; }
#pragma unroll 1
        for (int j = 0; j < QB / 32; ++j) {
            const int64_t qi = qb0 + j * 32 + lane;
            if (qi >= nq) continue;
            float2 q[H];
            load_row<DP>(shadow + (q_begin + qi) * DP, q);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                blo[2 * h] = fminf(blo[2 * h], q[h].x);         bhi[2 * h] = fmaxf(bhi[2 * h], q[h].x);
                blo[2 * h + 1] = fminf(blo[2 * h + 1], q[h].y); bhi[2 * h + 1] = fmaxf(bhi[2 * h + 1], q[h].y);
            }
            tmax = fmaxf(tmax, thresholds[qi]);
        }
        for (int off = 16; off > 0; off >>= 1) {
            tmax = fmaxf(tmax, __shfl_xor_sync(FULL, tmax, off));
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                blo[d] = fminf(blo[d], __shfl_xor_sync(FULL, blo[d], off));
                bhi[d] = fmaxf(bhi[d], __shfl_xor_sync(FULL, bhi[d], off));
            }
        }
        // ---- 32 tiles per round: lane = tile; a tile is needed iff its second-largest box gap is below tmax
        for (int64_t t0 = 0; t0 < n_tiles; t0 += 32) {
            const int64_t t = t0 + lane;
            bool need = false;
            if (t < n_tiles) {
                float g1 = -CUDART_INF_F, g2 = -CUDART_INF_F;
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    const float2 b = reinterpret_cast<const float2*>(tile_box)[t * DP + d];
                    const float g = fmaxf(b.x - bhi[d], blo[d] - b.y);
                    g2 = fmaxf(g2, fminf(g1, g));
                    g1 = fmaxf(g1, g);
                }
                need = g2 < tmax;
            }
            unsigned todo = __ballot_sync(FULL, need);
            while (todo) {
                const int k = __ffs(todo) - 1;
                todo &= todo - 1u;
                const int64_t tile = t0 + k;
                mbar_wait_relaxed(&empty_bar[slot], (phase_bits >> slot) & 1u);
                if (lane == 0) {
                    *reinterpret_cast<volatile int*>(&s_tile[slot]) = (int)tile;
                    float* dst = tiles + (size_t)slot * SLOT_FLOATS;
                    mbar_expect_tx(&full_bar[slot], SLOT_FLOATS * 4);
                    bulk_g2s(dst, shadow + tile * TILE_FLOATS, TILE_FLOATS * 4, &full_bar[slot]);
                    bulk_g2s(dst + TILE_FLOATS, sub_box + tile * BOX_FLOATS, BOX_FLOATS * 4, &full_bar[slot]);
                }
                __syncwarp();
                advance();
            }
        }
        // end marker
        mbar_wait_relaxed(&empty_bar[slot], (phase_bits >> slot) & 1u);
        if (lane == 0) {
            *reinterpret_cast<volatile int*>(&s_tile[slot]) = -1;
            mbar_arrive(&full_bar[slot]);
        }
        return;
    }

    // ---------------- filter warps: member `mem` of every group, scales, group boxes
    float2 qm[NG][H];
    IndScale sc[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int64_t qi = qb0 + warp * 32 + g * GQ + mem;
        const bool in = qi < nq;
        load_row<DP>(shadow + (q_begin + (in ? qi : 0)) * DP, qm[g]);
        const float T = in ? thresholds[qi] : -1.0f;  // dead slots count nothing
        sc[g] = make_ind_scale(T);
        s_ge[g][tid] = 0.0f;
#pragma unroll
        for (int h = 0; h < H; ++h) s_acc[g][h][tid] = make_float2(0.0f, 0.0f);
        float lo[DP], hi[DP];
        float tmax = in ? T : -CUDART_INF_F;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            lo[2 * h] = in ? qm[g][h].x : CUDART_INF_F;      hi[2 * h] = in ? qm[g][h].x : -CUDART_INF_F;
            lo[2 * h + 1] = in ? qm[g][h].y : CUDART_INF_F;  hi[2 * h + 1] = in ? qm[g][h].y : -CUDART_INF_F;
        }
#pragma unroll
        for (int off = 1; off < GQ; off <<= 1) {  // over the eight members (every stream does the same)
            tmax = fmaxf(tmax, __shfl_xor_sync(FULL, tmax, off));
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], off));
                hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], off));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < DP; ++d) s_gbox[warp][g][d] = make_float2(lo[d], hi[d]);
            s_gthr[warp][g] = tmax;
        }
    }
    __syncwarp();
    unsigned long long scans_total = 0;

    for (;;) {
        mbar_wait(&full_bar[slot], (phase_bits >> slot) & 1u);
        const int t = *reinterpret_cast<volatile int*>(&s_tile[slot]);
        if (t >= 0) {
            const float* tile = tiles + (size_t)slot * SLOT_FLOATS;
            const float2* boxes = reinterpret_cast<const float2*>(tile + TILE_FLOATS);
            // ---- classify 16 sub-tiles x 4 groups: lane = (half, sub-tile), groups 2 * half + {0, 1}
            unsigned scan_m[NG];
            const int sb_l = lane & 15, half = lane >> 4;
#pragma unroll
            for (int k = 0; k < NG / 2; ++k) {
                const int g = half * (NG / 2) + k;
                float g1 = -CUDART_INF_F, g2 = -CUDART_INF_F;
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    const float2 bx = boxes[sb_l * DP + d];
                    const float2 gb = s_gbox[warp][g][d];
                    const float gap = fmaxf(bx.x - gb.y, gb.x - bx.y);  // > 0 when the boxes are apart
                    g2 = fmaxf(g2, fminf(g1, gap));
                    g1 = fmaxf(g1, gap);
                }
                // (fp32 subtraction is monotone: a box gap is a lower bound of every computed |difference|, so a
                //  second-largest gap >= T leaves two coordinates with an indicator of 0 for every pair)
                const unsigned b_scan = __ballot_sync(FULL, g2 < s_gthr[warp][g]);
                scan_m[k] = b_scan & 0xffffu;
                scan_m[NG / 2 + k] = b_scan >> 16;
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                unsigned m = scan_m[g];
                if (m == 0u) continue;  // (warp-uniform)
                float sge = s_ge[g][tid];
                float2 acc[H];
#pragma unroll
                for (int h = 0; h < H; ++h) acc[h] = s_acc[g][h][tid];
#pragma unroll 1
                while (m) {
                    const int sb = __ffs(m) - 1;
                    m &= m - 1u;
                    ++scans_total;
                    const float* rows = tile + (sb * FAST_SUB + strm) * DP;
#pragma unroll 2
                    for (int i = 0; i < FAST_SUB / NG; ++i) {  // candidates NG * i + strm
                        float2 cc[H];
                        load_row<DP>(rows + NG * i * DP, cc);
                        float2 ind[H];
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            const float2 d = sub2(qm[g][h], cc[h]);
                            ind[h] = make_float2(ind_lt(fabsf(d.x), sc[g]), ind_lt(fabsf(d.y), sc[g]));
                        }
                        float2 s2 = ind[0];
#pragma unroll
                        for (int h = 1; h < H; ++h) s2 = add2(s2, ind[h]);
                        const float sigma = s2.x + s2.y;
                        // count_s += [sigma == DP] + [sigma == DP - 1] (1 - i_s) = [sigma >= DP - 1] - [sigma == DP - 1] i_s
                        const float ge = add_sat(sigma, -(float)(DP - 2));       // [sigma >= DP - 1]
                        const float e = ge - add_sat(sigma, -(float)(DP - 1));   // [sigma == DP - 1]
                        sge += ge;
                        const float2 ee = make_float2(e, e);
#pragma unroll
                        for (int h = 0; h < H; ++h) acc[h] = fma2(ind[h], ee, acc[h]);
                    }
                }
                s_ge[g][tid] = sge;
#pragma unroll
                for (int h = 0; h < H; ++h) s_acc[g][h][tid] = acc[h];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
        advance();
        if (t < 0) break;
    }
    // ---- the four streams' shares of every (query, subspace) -> the query's owner (lane = group * 8 + member)
    const int64_t qi = qb0 + warp * 32 + lane;
#pragma unroll
    for (int s = 0; s < DD; ++s) {
        int mine = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float2 a2 = s_acc[g][s / 2][tid];
            const float in_s = (s & 1) ? a2.y : a2.x;
            int v = (int)(s_ge[g][tid] - in_s);
            v += __shfl_xor_sync(FULL, v, 8);
            v += __shfl_xor_sync(FULL, v, 16);
            if (strm == g) mine = v;
        }
        if (qi < nq && mine) counts[(int64_t)s * nq + qi] += mine;
    }
    if (lane == 0 && pair_counter && scans_total)
        atomicAdd(pair_counter, scans_total * (unsigned long long)(FAST_SUB * GQ));
}

template <int DD>
static int launch_loo_group(const CountF32Args& a) {
    constexpr int DP = (DD + 1) & ~1;
    // ring depth: 2 (four CTAs per SM at D = 8); KSG_B200_LOO_STAGES=3 for experiments
    static const int st_env = [] { const char* e = getenv("KSG_B200_LOO_STAGES"); return e ? atoi(e) : 2; }();
    const int ST = st_env == 3 ? 3 : 2;
    auto kern = ST == 3 ? count_loo_group_kernel<DD, 3> : count_loo_group_kernel<DD, 2>;
    const size_t smem = (size_t)ST * (FAST_TILE * DP + FAST_NSUB * DP * 2) * sizeof(float);
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    kern<<<(unsigned)ceil_div(a.nq, FAST_THREADS), UNITS_THREADS, smem, a.st>>>(
        a.shadow, a.tile_box, a.sub_box, a.n_tiles, a.q_begin, a.nq, a.thresholds, a.counts, a.tile_counter);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("count_loo_group_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

// KSG_EUNSUPPORTED: not this kernel's case (fewer than 5 coordinates, no sub-tile boxes, too few queries to fill the
// machine with one CTA per 128 queries, or switched off by KSG_B200_LOO_GROUPS=0) -> the caller takes count_f32_kernel
int launch_count_loo_grouped(int d, const CountF32Args& a) {
    // KSG_B200_LOO_GROUPS: 0 = off, "force" = also for small query sets (tests)
    static const int mode = [] { const char* e = getenv("KSG_B200_LOO_GROUPS"); return !e ? 1 : (e[0] == '0' ? 0 : (e[0] == 'f' ? 2 : 1)); }();
    // one CTA per 128 queries, no candidate splits: it needs at least one full wave of CTAs (4 per SM) to use the
    // machine -- below that (small sets, a rank's slice of an 8-GPU run) the block-level kernel with its candidate
    // splits is the better choice (C4 on 8 ranks: 31 250 queries per rank = 244 CTAs for 592 slots)
    const int64_t min_queries = (int64_t)sm_count() * 4 * FAST_THREADS;
    if (mode == 0 || !a.sub_box || !a.windowed || (mode == 1 && a.nq < min_queries) ||
        a.n_tiles * (int64_t)FAST_TILE >= (1ll << 26))
        return KSG_EUNSUPPORTED;
    switch (d) {
        case 5: return launch_loo_group<5>(a);
        case 6: return launch_loo_group<6>(a);
        case 7: return launch_loo_group<7>(a);
        case 8: return launch_loo_group<8>(a);
        default: return KSG_EUNSUPPORTED;
    }
}

}  // namespace ksg

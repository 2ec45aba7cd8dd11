// Candidate-tile scheduling shared by the two filter kernels.
//
// A CTA owns a block of queries that are neighbours in the prepared order.  Candidate
// tiles (512 consecutive rows of the fp32 shadow) are visited outward from the CTA's own
// tile -- position 0 = own, then alternately the next tile to the right and to the left --
// so that near tiles come first and the pruning radius shrinks quickly.
//
// Pruning ("windowed" engine): every tile has a bounding box (min / max of each
// coordinate, prepare.cuh), the CTA has the bounding box of its queries; a tile whose
// box lies farther than the CTA's current radius R from the query box (max-norm gap, in
// the relevant coordinates) cannot contain a neighbour / a counted point of ANY query of
// the CTA and is skipped.  Because fp32 subtraction is monotone, the computed gap is a
// lower bound of every computed pair distance -- no slack needed.  Walk positions are
// tested 128 at a time (one per thread), survivors are compacted in walk order into a
// shared-memory list, and the list is streamed through the STAGES-deep bulk-copy ring.
// R is refreshed between rounds (it only shrinks, so an earlier skip stays valid).
#pragma once

#include "common.cuh"
#include "prepare.cuh"

namespace ksg {

constexpr int FAST_THREADS = 128;

// ---- mbarrier / bulk-copy primitives (shared::cta addresses)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{ .reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// the copy warp's wait: long hardware-suspended tries and a sleep in between, so that it does
// not take issue slots from the filter warps while the ring is full
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{ .reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
        if (done) return;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// tile at walk position p (0 = own)
__device__ __forceinline__ int64_t walk_tile(int64_t p, int64_t own, int64_t n_tiles) {
    if (p == 0) return own;
    const int64_t r = n_tiles - 1 - own, l = own, m = r < l ? r : l;
    if (p <= 2 * m) {
        const int64_t j = (p + 1) >> 1;
        return (p & 1) ? own + j : own - j;
    }
    const int64_t q = p - 2 * m;
    return r > l ? own + m + q : own - m - q;
}

struct TileQueueShared {
    int32_t todo[FAST_THREADS];
    int32_t warp_count[FAST_THREADS / 32];
    float red[FAST_THREADS / 32];
    float qlo[KSG_MAX_DIM > 8 ? 8 : KSG_MAX_DIM], qhi[KSG_MAX_DIM > 8 ? 8 : KSG_MAX_DIM];
    float red_lo[FAST_THREADS / 32][8], red_hi[FAST_THREADS / 32][8];
};

// block-wide max of a per-thread value (all threads get it); contains two barriers
__device__ __forceinline__ float block_max(float v, TileQueueShared& sh) {
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmaxf(fmaxf(sh.red[0], sh.red[1]), fmaxf(sh.red[2], sh.red[3]));
}

// block-wide sum of a per-thread value (all threads get it); contains two barriers
__device__ __forceinline__ float block_sum(float v, TileQueueShared& sh) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
    __syncthreads();
    return (sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]);
}

// bounding box of the CTA's live queries -> sh.qlo / sh.qhi
template <int DP>
__device__ __forceinline__ void block_query_box(const float (&lo)[DP], const float (&hi)[DP], TileQueueShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        float a = lo[d], b = hi[d];
        for (int off = 16; off > 0; off >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, off));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, off));
        }
        if (lane == 0) { sh.red_lo[warp][d] = a; sh.red_hi[warp][d] = b; }
    }
    __syncthreads();
    if (threadIdx.x < DP) {
        const int d = threadIdx.x;
        sh.qlo[d] = fminf(fminf(sh.red_lo[0][d], sh.red_lo[1][d]), fminf(sh.red_lo[2][d], sh.red_lo[3][d]));
        sh.qhi[d] = fmaxf(fmaxf(sh.red_hi[0][d], sh.red_hi[1][d]), fmaxf(sh.red_hi[2][d], sh.red_hi[3][d]));
    }
    __syncthreads();
}

// per-coordinate gaps between the CTA's query box and tile t's box (>= 0)
template <int DP>
__device__ __forceinline__ void box_gaps(const float* __restrict__ tile_box, int64_t t, const TileQueueShared& sh,
                                         float (&gap)[DP]) {
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        const float2 b = reinterpret_cast<const float2*>(tile_box)[t * DP + d];
        gap[d] = fmaxf(0.0f, fmaxf(b.x - sh.qhi[d], sh.qlo[d] - b.y));
    }
}

// Drives the walk.  test(t) -> does tile t survive (per thread, any thread may be asked);
// compute(tile_smem, t) processes one staged tile (all threads); refresh() recomputes the
// pruning radius (block-uniform, may contain barriers).  Returns the number of tiles processed.
template <int DP, int STAGES, class TestFn, class ComputeFn, class RefreshFn>
__device__ __forceinline__ int run_tile_queue(const float* __restrict__ shadow, float* tiles, uint64_t* full_bar,
                                              TileQueueShared& sh, int64_t own, int64_t n_tiles, int split,
                                              int splits, bool prune, TestFn test, ComputeFn compute,
                                              RefreshFn refresh) {
    constexpr int TILE_FLOATS = FAST_TILE * DP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t issue_cnt = 0, consume_cnt = 0;  // kernel-lifetime counters: slot = cnt % STAGES
    int tiles_done = 0;
    int64_t pos = 0;
    int chunk = prune ? 32 : FAST_THREADS;  // thresholds are seeded: R is finite from the first round
    while (pos < n_tiles) {
        const int c = (int)((n_tiles - pos) < chunk ? (n_tiles - pos) : chunk);
        // ---- test this round's walk positions, one per thread
        bool keep = false;
        int64_t t = 0;
        if (tid < c) {
            const int64_t p = pos + tid;
            if ((p % splits) == split) {
                t = walk_tile(p, own, n_tiles);
                keep = !prune || test(t);
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sh.warp_count[warp] = __popc(ballot);
        __syncthreads();
        int offset = 0, total = 0;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w) {
            if (w < warp) offset += sh.warp_count[w];
            total += sh.warp_count[w];
        }
        if (keep) sh.todo[offset + __popc(ballot & ((1u << lane) - 1u))] = (int32_t)t;
        __syncthreads();

        // ---- stream the survivors through the ring
        int issued = 0;
        for (int i = 0; i < total; ++i) {
            while (issue_cnt - consume_cnt < STAGES && issued < total) {
                if (tid == 0) {
                    const int slot = (int)(issue_cnt % STAGES);
                    mbar_expect_tx(&full_bar[slot], TILE_FLOATS * 4);
                    bulk_g2s(tiles + (size_t)slot * TILE_FLOATS, shadow + (int64_t)sh.todo[issued] * TILE_FLOATS,
                             TILE_FLOATS * 4, &full_bar[slot]);
                }
                ++issue_cnt;
                ++issued;
            }
            const int slot = (int)(consume_cnt % STAGES);
            mbar_wait(&full_bar[slot], (uint32_t)((consume_cnt / STAGES) & 1));
            compute(tiles + (size_t)slot * TILE_FLOATS, (int64_t)sh.todo[i]);
            __syncthreads();  // tile consumed: its slot may be refilled
            ++consume_cnt;
            ++tiles_done;
        }
        pos += c;
        if (prune) {
            refresh();
            chunk = FAST_THREADS;
        }
    }
    return tiles_done;
}

}  // namespace ksg

// Key sort of ksg_prepare: (key, index) pairs of up to 2^21 elements, ascending by (key, index).
//
// A million-element radix sort is latency bound on this machine: every 8-bit onesweep pass moves
// 24 MB in ~18 us, a float64 key needs eight of them (profiles/r2a_c2_windowed_launches.csv: 160 us
// per coordinate, 90 us for the 32-bit curve keys).  The prepared set needs each coordinate's
// exact order (binary-search counts, band look-ups) and the order of the curve keys -- three sorts
// on the critical path of every call.  This is a SAMPLE SORT sized for that regime, five short
// launches whatever the key width:
//   1. ss_sample_kernel    one CTA sorts nb * 8 pseudo-randomly drawn (key, index) pairs in shared
//                          memory (bitonic) and keeps every 8th as a splitter;
//   2. ss_classify_kernel  every element finds its bucket by binary search among the splitters
//                          (shared memory) and takes its arrival number in the bucket (one atomic);
//   3. ss_scan_kernel      exclusive scan of the bucket sizes;
//   4. ss_scatter_kernel   element -> bucket_start + arrival number;
//   5. ss_bucket_kernel    one CTA per bucket: bitonic sort in shared memory, written to its final
//                          place.  Buckets hold ~n / nb <= 1024 elements; a bucket that outgrows the
//                          shared-memory capacity (probability ~1e-9 per sort with 8-fold
//                          oversampling; adversarial input) is ranked by counting instead.
// Pairs are compared as (key, index) -- all distinct -- so the buckets split evenly even when every
// key is the same, the result does not depend on the arrival order inside a bucket, and it equals a
// stable sort by key (cub::DeviceRadixSort with an iota payload, which it replaces).
// float64 keys are compared through the usual order-preserving bit transform (negative: all bits
// flipped, else the sign bit set): -inf / negative NaNs first, +inf / positive NaNs last.
#pragma once

#include <string.h>

#include "common.cuh"

namespace ksg {

constexpr int SS_MAX_BUCKETS = 2048;
constexpr int SS_OVERSAMPLE = 8;
constexpr int SS_BUCKET_CAP = 4096;            // elements a bucket CTA can sort in shared memory
constexpr int SS_BUCKET_THREADS = 512;
constexpr int SS_SAMPLE_THREADS = 1024;
constexpr int64_t SS_MAX_N = (int64_t)1 << 21;  // larger arrays stay on the radix sort
constexpr int SS_ITEMS = 4;                    // elements per thread of the classify / scatter kernels

__host__ __device__ inline uint64_t ss_key_of(double v) {
    long long b;
#ifdef __CUDA_ARCH__
    b = __double_as_longlong(v);
#else
    memcpy(&b, &v, 8);
#endif
    const uint64_t u = (uint64_t)b;
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline uint64_t ss_key_of(unsigned long long v) { return (uint64_t)v; }
__device__ __forceinline__ void ss_store_key(double* out, uint64_t k) {
    const uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    *out = __longlong_as_double((long long)u);
}
__device__ __forceinline__ void ss_store_key(unsigned long long* out, uint64_t k) { *out = k; }

__host__ __device__ inline bool ss_less(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}

// number of buckets for n elements: a power of two with ~512 elements per bucket, at most 2048;
// one bucket (no sampling at all) while the whole array fits one CTA
static inline int ss_buckets(int64_t n) {
    if (n <= SS_BUCKET_CAP) return 1;
    int nb = 2;
    while (nb < SS_MAX_BUCKETS && (int64_t)nb * 512 < n) nb <<= 1;
    return nb;
}

struct SsLayout {
    size_t off_tkeys, off_tidx, off_bid, off_rank, off_spk, off_spi, off_hist, off_start, total;
};
static inline SsLayout ss_layout(int64_t n) {
    SsLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) / 256 * 256; return at; };
    L.off_tkeys = take((size_t)n * 8);
    L.off_tidx = take((size_t)n * 4);
    L.off_bid = take((size_t)n * 2);
    L.off_rank = take((size_t)n * 4);
    L.off_spk = take((size_t)SS_MAX_BUCKETS * 8);
    L.off_spi = take((size_t)SS_MAX_BUCKETS * 4);
    L.off_hist = take((size_t)SS_MAX_BUCKETS * 4);
    L.off_start = take((size_t)(SS_MAX_BUCKETS + 1) * 4);
    L.total = o;
    return L;
}

// in-place bitonic sort of P = 2^m (key, index) pairs in shared memory, ascending by (key, index)
__device__ __forceinline__ void ss_bitonic(uint64_t* sk, uint32_t* si, int P, int tid, int nthreads) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const uint64_t ka = sk[i], kb = sk[p];
                const uint32_t ia = si[i], ib = si[p];
                const bool up = (i & k) == 0;
                if (ss_less(kb, ib, ka, ia) == up) {
                    sk[i] = kb; sk[p] = ka;
                    si[i] = ib; si[p] = ia;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint32_t ss_hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ---- 1. splitters.  nb * 8 samples at pseudo-random positions; splitter b = sample (b + 1) * 8.
// Also zeroes the bucket counters.
template <class KeyT>
static __global__ void __launch_bounds__(SS_SAMPLE_THREADS)
ss_sample_kernel(const KeyT* __restrict__ keys, int64_t n, int nb, uint64_t* __restrict__ spk,
                 uint32_t* __restrict__ spi, int32_t* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    const int S = nb * SS_OVERSAMPLE;
    uint64_t* sk = reinterpret_cast<uint64_t*>(ss_smem);
    uint32_t* si = reinterpret_cast<uint32_t*>(sk + S);
    const int tid = threadIdx.x;
    for (int b = tid; b < nb; b += SS_SAMPLE_THREADS) hist[b] = 0;
    for (int s = tid; s < S; s += SS_SAMPLE_THREADS) {
        // one draw per stratum of n / S consecutive elements: representative for sorted, periodic
        // and shuffled input alike
        const uint64_t lo = (uint64_t)s * (uint64_t)n / (uint64_t)S, hi = (uint64_t)(s + 1) * (uint64_t)n / (uint64_t)S;
        const uint64_t e = lo + (hi > lo ? ss_hash((uint32_t)s * 2654435761u + 12345u) % (hi - lo) : 0);
        sk[s] = ss_key_of(keys[e]);
        si[s] = (uint32_t)e;
    }
    __syncthreads();
    ss_bitonic(sk, si, S, tid, SS_SAMPLE_THREADS);
    for (int b = tid; b + 1 < nb; b += SS_SAMPLE_THREADS) {
        spk[b] = sk[(b + 1) * SS_OVERSAMPLE];
        spi[b] = si[(b + 1) * SS_OVERSAMPLE];
    }
}

// ---- 2. bucket of every element (= number of splitters <= it) and its arrival number there
template <class KeyT>
static __global__ void __launch_bounds__(256)
ss_classify_kernel(const KeyT* __restrict__ keys, int64_t n, int nb, const uint64_t* __restrict__ spk,
                   const uint32_t* __restrict__ spi, int32_t* __restrict__ hist, uint16_t* __restrict__ bid,
                   uint32_t* __restrict__ rank) {
    __shared__ uint64_t s_k[SS_MAX_BUCKETS];
    __shared__ uint32_t s_i[SS_MAX_BUCKETS];
    for (int b = threadIdx.x; b + 1 < nb; b += 256) { s_k[b] = spk[b]; s_i[b] = spi[b]; }
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (256 * SS_ITEMS);
#pragma unroll
    for (int it = 0; it < SS_ITEMS; ++it) {
        const int64_t e = base + it * 256 + threadIdx.x;
        if (e >= n) break;
        const uint64_t k = ss_key_of(keys[e]);
        int lo = 0, hi = nb - 1;  // first splitter index with splitter > element
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ss_less(k, (uint32_t)e, s_k[mid], s_i[mid])) hi = mid; else lo = mid + 1;
        }
        bid[e] = (uint16_t)lo;
        rank[e] = (uint32_t)atomicAdd(&hist[lo], 1);
    }
}

// ---- 3. bucket starts (nb <= 2048: one CTA)
static __global__ void __launch_bounds__(1024)
ss_scan_kernel(const int32_t* __restrict__ hist, int nb, int32_t* __restrict__ start) {
    __shared__ int s_scan[1024];
    const int tid = threadIdx.x;
    const int a = 2 * tid < nb ? hist[2 * tid] : 0, b = 2 * tid + 1 < nb ? hist[2 * tid + 1] : 0;
    s_scan[tid] = a + b;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = tid >= off ? s_scan[tid - off] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    const int excl = s_scan[tid] - (a + b);
    if (2 * tid < nb) start[2 * tid] = excl;
    if (2 * tid + 1 < nb) start[2 * tid + 1] = excl + a;
    if (tid == 1023) start[nb] = s_scan[1023];
}

// ---- 4. scatter into the buckets
template <class KeyT>
static __global__ void __launch_bounds__(256)
ss_scatter_kernel(const KeyT* __restrict__ keys, int64_t n, const uint16_t* __restrict__ bid,
                  const uint32_t* __restrict__ rank, const int32_t* __restrict__ start,
                  uint64_t* __restrict__ tkeys, uint32_t* __restrict__ tidx) {
    const int64_t base = (int64_t)blockIdx.x * (256 * SS_ITEMS);
#pragma unroll
    for (int it = 0; it < SS_ITEMS; ++it) {
        const int64_t e = base + it * 256 + threadIdx.x;
        if (e >= n) break;
        const int64_t pos = (int64_t)start[bid[e]] + rank[e];
        tkeys[pos] = ss_key_of(keys[e]);
        tidx[pos] = (uint32_t)e;
    }
}

// ---- 5. sort every bucket and write it to its final place.  nb == 1: the "bucket" is the whole
// input, read straight from `keys` (no sampling / scatter launches before this one).
template <class KeyT>
static __global__ void __launch_bounds__(SS_BUCKET_THREADS)
ss_bucket_kernel(const KeyT* __restrict__ keys, const uint64_t* __restrict__ tkeys, const uint32_t* __restrict__ tidx,
                 const int32_t* __restrict__ start, int64_t n, int nb, KeyT* __restrict__ keys_out,
                 int32_t* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(ss_smem);
    uint32_t* si = reinterpret_cast<uint32_t*>(sk + SS_BUCKET_CAP);
    const int tid = threadIdx.x;
    const int64_t b0 = nb == 1 ? 0 : start[blockIdx.x], b1 = nb == 1 ? n : start[blockIdx.x + 1];
    const int m = (int)(b1 - b0);
    if (m <= 0) return;
    if (m <= SS_BUCKET_CAP) {
        int P = 2;
        while (P < m) P <<= 1;
        for (int i = tid; i < P; i += SS_BUCKET_THREADS) {
            if (i < m) {
                sk[i] = nb == 1 ? ss_key_of(keys[i]) : tkeys[b0 + i];
                si[i] = nb == 1 ? (uint32_t)i : tidx[b0 + i];
            } else {
                sk[i] = ~0ull;  // padding sorts last (index ~0 is no element's)
                si[i] = ~0u;
            }
        }
        __syncthreads();
        ss_bitonic(sk, si, P, tid, SS_BUCKET_THREADS);
        for (int i = tid; i < m; i += SS_BUCKET_THREADS) {
            ss_store_key(&keys_out[b0 + i], sk[i]);
            idx_out[b0 + i] = (int32_t)si[i];
        }
        return;
    }
    // oversized bucket (never with representative samples): rank by counting, straight from global
    for (int i = tid; i < m; i += SS_BUCKET_THREADS) {
        const uint64_t k = tkeys[b0 + i];
        const uint32_t e = tidx[b0 + i];
        int r = 0;
        for (int j = 0; j < m; ++j) r += ss_less(tkeys[b0 + j], tidx[b0 + j], k, e) ? 1 : 0;
        ss_store_key(&keys_out[b0 + r], k);
        idx_out[b0 + r] = (int32_t)e;
    }
}

// Enqueue the sort of n (key, index = position) pairs: keys_out ascending, idx_out the positions.
// `tmp` holds ss_layout(n).total bytes.  Returns a cudaError_t (launch errors).
template <class KeyT>
static cudaError_t sample_sort_pairs(void* tmp, const KeyT* keys_in, KeyT* keys_out, int32_t* idx_out, int64_t n,
                                     cudaStream_t st, int* launches) {
    const SsLayout L = ss_layout(n);
    char* base = (char*)tmp;
    uint64_t* tkeys = (uint64_t*)(base + L.off_tkeys);
    uint32_t* tidx = (uint32_t*)(base + L.off_tidx);
    uint16_t* bid = (uint16_t*)(base + L.off_bid);
    uint32_t* rank = (uint32_t*)(base + L.off_rank);
    uint64_t* spk = (uint64_t*)(base + L.off_spk);
    uint32_t* spi = (uint32_t*)(base + L.off_spi);
    int32_t* hist = (int32_t*)(base + L.off_hist);
    int32_t* start = (int32_t*)(base + L.off_start);
    const int nb = ss_buckets(n);
    const size_t bucket_smem = (size_t)SS_BUCKET_CAP * 12;
    // (function attributes are per device; set once, outside any stream capture that replays this later)
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(ss_bucket_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bucket_smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ss_sample_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 SS_MAX_BUCKETS * SS_OVERSAMPLE * 12);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (nb > 1) {
        const size_t sample_smem = (size_t)nb * SS_OVERSAMPLE * 12;
        const unsigned grid = (unsigned)((n + 256 * SS_ITEMS - 1) / (256 * SS_ITEMS));
        ss_sample_kernel<KeyT><<<1, SS_SAMPLE_THREADS, sample_smem, st>>>(keys_in, n, nb, spk, spi, hist);
        ss_classify_kernel<KeyT><<<grid, 256, 0, st>>>(keys_in, n, nb, spk, spi, hist, bid, rank);
        ss_scan_kernel<<<1, 1024, 0, st>>>(hist, nb, start);
        ss_scatter_kernel<KeyT><<<grid, 256, 0, st>>>(keys_in, n, bid, rank, start, tkeys, tidx);
        *launches += 4;
    }
    ss_bucket_kernel<KeyT><<<(unsigned)nb, SS_BUCKET_THREADS, bucket_smem, st>>>(keys_in, tkeys, tidx, start, n, nb,
                                                                                  keys_out, idx_out);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace ksg

// extern "C" entry points of the fast path (prepare / fast_knn / fast_count / axis_count).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "exact_f64.cuh"
#include "axis_buckets.cuh"
#include "fast_count.cuh"
#include "fast_knn.cuh"
#include "fast_launch.h"
#include "prepare.cuh"
#include "sample_sort.cuh"

namespace ksg {

static inline size_t cub_sort_bytes_query(int64_t n) {
    size_t bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const double*)nullptr, (double*)nullptr,
                                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
    if (e != cudaSuccess || bytes == 0) {
        (void)cudaGetLastError();
        bytes = (size_t)n * 16 + (16u << 20);  // no device visible: conservative estimate
    }
    return bytes;
}

// (the size query walks CUB's whole dispatch on the host: remember the last few answers)
static inline size_t cub_sort_bytes_cached(int64_t n);
// scratch of one key sort: whichever of the two implementations runs (sample_sort.cuh / CUB)
static inline size_t cub_sort_bytes(int64_t n) {
    const size_t a = cub_sort_bytes_cached(n), b = n <= SS_MAX_N ? ss_layout(n).total : 0;
    return a > b ? a : b;
}
static inline size_t cub_sort_bytes_cached(int64_t n) {
    static std::mutex mu;
    static int64_t seen_n[8];
    static size_t seen_bytes[8];
    static int used = 0, next = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < used; ++i)
            if (seen_n[i] == n) return seen_bytes[i];
    }
    const size_t bytes = cub_sort_bytes_query(n);
    std::lock_guard<std::mutex> lock(mu);
    seen_n[next] = n;
    seen_bytes[next] = bytes;
    next = (next + 1) % 8;
    if (used < 8) ++used;
    return bytes;
}

static inline PreparedLayout prepared_layout(int64_t n, int dim) {
    PreparedLayout L;
    L.ld = (int64_t)align_up((size_t)n, 32);
    L.n_padded = (int64_t)align_up((size_t)n, FAST_TILE);
    L.dim_padded = (dim + 1) & ~1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    L.off_perm = take((size_t)L.ld * 4);
    L.off_sorted = take((size_t)dim * L.ld * 8);
    L.off_shadow = take((size_t)L.n_padded * L.dim_padded * 4);
    L.off_axis_vals = take((size_t)dim * L.ld * 8);
    L.off_axis_pos = take((size_t)dim * L.ld * 4);
    L.off_center = take((size_t)KSG_MAX_DIM * 8);
    L.off_maxabs = take((size_t)KSG_MAX_DIM * 8);
    L.off_flags = take(64);
    L.off_tile_box = take((size_t)(L.n_padded / FAST_TILE) * L.dim_padded * 2 * 4);
    L.off_sub_box = take((size_t)(L.n_padded / FAST_TILE) * FAST_NSUB * L.dim_padded * 2 * 4);
    L.off_iota = take((size_t)L.ld * 4);
    L.off_ranks = take((size_t)dim * L.ld * 4);
    L.off_mkeys = take((size_t)L.ld * 8);
    L.off_mkeys_out = take((size_t)L.ld * 8);
    L.off_invperm = take((size_t)L.ld * 4);
    L.off_curve_table = take((size_t)KSG_MAX_DIM * CURVE_TABLE * 4);
    L.off_stats = take((size_t)KSG_MAX_DIM * STATS_BLOCKS * 2 * 8);
    L.off_exact = take((size_t)(L.n_padded / 256 + 1) * EXACT_INFO_INTS * 4);  // one record per gather block
    L.cub_bytes = align_up(cub_sort_bytes(n), 256);
    L.sort_lanes = dim < SORT_LANES ? dim : SORT_LANES;
    // one CUB scratch area per concurrent coordinate sort (early order: area 0 = the key sort on the
    // caller's stream, areas 1.. = the coordinate sorts on the side streams)
    L.off_cub = take(L.cub_bytes * SORT_LANES);
    L.total = o;
    return L;
}

// The per-coordinate sorts of ksg_prepare are independent and, at a million points, latency bound
// (a one-sweep pass moves 24 MB in ~15 us): they run side by side on up to SORT_LANES streams --
// lane 0 is the caller's stream, the others are helper streams of this library (one set per
// device, created on first use) forked from / joined to the caller's stream with events.
constexpr int AXES_EVENTS = 256;  // ring of "exact coordinate sorts of this prepared set are done" events
struct SortLanes {
    cudaStream_t helper[SORT_LANES - 1];
    cudaEvent_t fork, join[SORT_LANES - 1];
    cudaEvent_t perm_ready;               // early order: the curve permutation exists (main stream)
    cudaEvent_t axes_ready[AXES_EVENTS];  // early order: axis_vals / axis_pos are final (side stream)
    unsigned next_axes;
    bool ready;
};
static SortLanes* sort_lanes_for_device() {
    static SortLanes lanes[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    SortLanes& L = lanes[dev];
    if (!L.ready) {
        if (cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int i = 0; i < SORT_LANES - 1; ++i) {
            if (cudaStreamCreateWithFlags(&L.helper[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&L.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&L.perm_ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int i = 0; i < AXES_EVENTS; ++i)
            if (cudaEventCreateWithFlags(&L.axes_ready[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        L.next_axes = 0;
        L.ready = true;
    }
    return &L;
}


// ---- launch-sequence cache.  ksg_prepare is ~60 short launches (three radix sorts of 5-8 passes
// each, every pass a memset + a ~17 us kernel): at a million points the host cannot enqueue two
// concurrent sorts as fast as the device retires them, and every launch leaves a 2-3 us gap.  A
// sequence that is requested again with the SAME arguments (pointers, sizes, options -- what an
// estimator called in a loop, or an allocator that recycles its blocks, produces) is captured once
// into a CUDA graph and replayed with one launch from then on.  First sighting: direct launches
// (a one-off call never pays for a capture); second sighting: capture + instantiate + launch;
// later: replay.  KSG_B200_GRAPHS=0 switches the cache off.
struct GraphKey {
    int32_t dev, kind;
    const void* ptr[5];
    int64_t i64[3];
    int32_t i32[4];
};
enum { GRAPH_AXIS_SORT = 1, GRAPH_PREPARE_TAIL = 2, GRAPH_PREPARE_EARLY = 3 };

// KSG_B200_EARLY_ORDER=0: the curve waits for the exact coordinate sorts again (A/B measurements)
static bool early_order() {
    static const bool v = [] {
        const char* e = getenv("KSG_B200_EARLY_ORDER");
        return !(e && e[0] == '0');
    }();
    return v;
}

// Order `st` after the side-stream work of an early-order prepared set (no-op otherwise).
static int wait_axes(const ksg_prepared* P, cudaStream_t st) {
    if (P->axes_ready) KSG_CUDA_TRY(cudaStreamWaitEvent(st, (cudaEvent_t)P->axes_ready, 0));
    return KSG_OK;
}
constexpr size_t GRAPH_CACHE_ENTRIES = 32;

class GraphCache {
    struct Entry {
        GraphKey key;
        cudaGraphExec_t exec;
        unsigned long long launches;
        uint64_t stamp;
    };
    std::mutex mu_;
    std::vector<Entry> entries_;
    uint64_t clock_ = 0;
    bool disabled_;
    cudaStream_t capture_stream_[64] = {};

    cudaStream_t capture_stream(int dev) {
        if (dev < 0 || dev >= 64) return nullptr;
        if (!capture_stream_[dev] &&
            cudaStreamCreateWithFlags(&capture_stream_[dev], cudaStreamNonBlocking) != cudaSuccess) {
            (void)cudaGetLastError();
            capture_stream_[dev] = nullptr;
        }
        return capture_stream_[dev];
    }

    template <class Body>
    bool capture(Entry& e, Body& body) {
        cudaStream_t cap = capture_stream(e.key.dev);
        if (!cap) return false;
        if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        const unsigned long long l0 = ksg_launch_count();
        const int rc = body(cap);
        const unsigned long long l1 = ksg_launch_count();
        cudaGraph_t graph = nullptr;
        cudaError_t err = cudaStreamEndCapture(cap, &graph);
        cudaGraphExec_t exec = nullptr;
        if (rc == KSG_OK && err == cudaSuccess && graph) err = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) (void)cudaGraphDestroy(graph);
        if (rc != KSG_OK || err != cudaSuccess || !exec) {
            (void)cudaGetLastError();
            note_launches(-(long long)(l1 - l0));  // nothing was launched
            return false;
        }
        e.exec = exec;
        e.launches = l1 - l0;  // (counted once by the capture pass itself)
        return true;
    }

public:
    GraphCache() {
        const char* env = getenv("KSG_B200_GRAPHS");
        disabled_ = env && env[0] == '0';
    }

    // Enqueue body(stream)'s launches on `st`: directly, or as the replay of their captured graph.
    template <class Body>
    int run(const GraphKey& key, cudaStream_t st, Body&& body) {
        if (disabled_) return body(st);
        cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &status) != cudaSuccess || status != cudaStreamCaptureStatusNone) {
            (void)cudaGetLastError();
            return body(st);  // the caller is capturing: stay part of ITS graph
        }
        std::unique_lock<std::mutex> lock(mu_);
        Entry* hit = nullptr;
        for (Entry& e : entries_)
            if (memcmp(&e.key, &key, sizeof(GraphKey)) == 0) { hit = &e; break; }
        if (!hit) {
            if (entries_.size() >= GRAPH_CACHE_ENTRIES) {
                size_t lru = 0;
                for (size_t i = 1; i < entries_.size(); ++i)
                    if (entries_[i].stamp < entries_[lru].stamp) lru = i;
                if (entries_[lru].exec) (void)cudaGraphExecDestroy(entries_[lru].exec);
                entries_.erase(entries_.begin() + (long)lru);
            }
            entries_.push_back(Entry{key, nullptr, 0ull, ++clock_});
            lock.unlock();
            return body(st);
        }
        hit->stamp = ++clock_;
        bool counted = false;
        if (!hit->exec) {
            if (!capture(*hit, body)) {
                disabled_ = true;  // capture is not possible here: direct launches from now on
                lock.unlock();
                return body(st);
            }
            counted = true;
        }
        const cudaError_t err = cudaGraphLaunch(hit->exec, st);
        if (err != cudaSuccess) {
            set_error("cudaGraphLaunch: %s", cudaGetErrorString(err));
            (void)cudaGetLastError();
            return KSG_ECUDA;
        }
        if (!counted) note_launches((long long)hit->launches);
        return KSG_OK;
    }
};
static GraphCache& graph_cache() {
    static GraphCache cache;
    return cache;
}
static inline GraphKey graph_key(int kind, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> i64,
                                 std::initializer_list<int32_t> i32) {
    GraphKey k;
    memset(&k, 0, sizeof(k));
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = -1; }
    k.dev = dev;
    k.kind = kind;
    int i = 0;
    for (const void* p : ptrs) k.ptr[i++] = p;
    i = 0;
    for (int64_t v : i64) k.i64[i++] = v;
    i = 0;
    for (int32_t v : i32) k.i32[i++] = v;
    return k;
}

#define KSG_LAUNCHED2()                   \
    do {                                  \
        note_launch();                    \
        KSG_CUDA_TRY(cudaGetLastError()); \
    } while (0)

static inline cudaStream_t as_stream2(void* s) { return (cudaStream_t)s; }

// KSG_B200_SORT=sample moves the key sorts of ksg_prepare to the sample sort of sample_sort.cuh.
// Measured on C2 (profiles/r2b_*): its single-CTA splitter sort and its shared-memory bitonic bucket
// sorts make it SLOWER than eight latency-bound onesweep passes (prepare 1.12 ms against 0.38 ms), so
// cub::DeviceRadixSort stays the default; the code is kept for ksg_sort_f64(algorithm 0) cross-checks.
static bool use_sample_sort(int64_t n) {
    static const bool on = [] {
        const char* e = getenv("KSG_B200_SORT");
        return e && e[0] == 's';
    }();
    return on && n <= SS_MAX_N;
}

// (key, position) pairs by sample sort; `tmp` holds cub_sort_bytes(n)
template <class KeyT>
static int sample_sort_launch(void* tmp, const KeyT* keys_in, KeyT* keys_out, int32_t* idx_out, int64_t n,
                              cudaStream_t st) {
    int launches = 0;
    const cudaError_t e = sample_sort_pairs<KeyT>(tmp, keys_in, keys_out, idx_out, n, st, &launches);
    note_launches(launches);
    if (e != cudaSuccess) {
        set_error("sample sort: %s", cudaGetErrorString(e));
        return KSG_ECUDA;
    }
    return KSG_OK;
}

// keys + the positions 0..n-1 (`vals_in` = iota) by key, stable
static int sort_pairs(void* cub_tmp, size_t cub_bytes, const double* keys_in, double* keys_out,
                      const int32_t* vals_in, int32_t* vals_out, int64_t n, cudaStream_t st) {
    if (use_sample_sort(n)) return sample_sort_launch<double>(cub_tmp, keys_in, keys_out, vals_out, n, st);
    size_t need = cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, need, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                                    64, st);
    if (e != cudaSuccess) {
        set_error("cub::DeviceRadixSort::SortPairs: %s", cudaGetErrorString(e));
        return KSG_ECUDA;
    }
    note_launch();
    return KSG_OK;
}

// Experiment switches of the windowed k-NN stage (read once): KSG_B200_TIGHT_SEED=0 restores the
// legacy kcap-th-distance seeds, KSG_B200_FUSED_REFINE=0 the separate refine launch for every block.
static bool env_flag(const char* name, bool dflt) {
    const char* e = getenv(name);
    return e && e[0] ? e[0] != '0' : dflt;
}
static bool tight_seeds() { static const bool v = env_flag("KSG_B200_TIGHT_SEED", true); return v; }
static bool fused_refine() { static const bool v = env_flag("KSG_B200_FUSED_REFINE", true); return v; }

static void fill_refine_args(KnnF32Args& a, const ksg_prepared* P, int kprime, const double* qsorted, int64_t qld,
                             const double* maxabs, int64_t* d_idx_out, double* d_eps_out, int32_t* d_flags, int32_t* d_n_flagged) {
    // the window excludes the query itself when it is a member of the candidate set: the (k' - 1)-th
    // smallest distance to the OTHER rows bounds the k'-th smallest with self (0) included
    const int kseed = qsorted ? kprime : (kprime > 1 ? kprime - 1 : 1);
    a.kseed = tight_seeds() && kseed < a.kcap ? kseed : a.kcap;
    a.ra.fused = fused_refine() ? 1 : 0;
    a.ra.kp = kprime;
    a.ra.dim = P->dim;
    a.ra.sorted = P->sorted_f64;
    a.ra.ld = P->ld;
    a.ra.qsorted = qsorted;
    a.ra.qld = qld;
    a.ra.center = P->center;
    a.ra.maxabs = maxabs;
    a.ra.exact32 = qsorted ? nullptr : P->flags + 8;  // (foreign queries: their shadows were never analysed)
    a.ra.perm = P->perm;
    a.ra.idx_out = d_idx_out;
    a.ra.eps_out = d_eps_out;
    a.ra.flags = d_flags;
    a.ra.n_flagged = d_n_flagged;
}

struct FastKnnPlan {
    int kcap, splits, qb;
    int64_t qblocks, max_units, plan_capacity;
    size_t d32_bytes, idx_bytes;
    // windowed engine: planning pass + unit table
    size_t seed_bytes, plan_bytes, count_bytes, first_bytes, ublock_bytes, meta_bytes, done_bytes;
    size_t total() const {
        return d32_bytes + idx_bytes + seed_bytes + plan_bytes + count_bytes + first_bytes + ublock_bytes + meta_bytes +
               done_bytes;
    }
};
static FastKnnPlan fast_knn_plan(int64_t n, int64_t nq, int dim, int kprime, int windowed, int margin = KNN_MARGIN) {
    FastKnnPlan p;
    memset(&p, 0, sizeof(p));
    const int dp = (dim + 1) & ~1;
    p.kcap = kprime + margin;
    p.qb = windowed ? knn_units_queries_per_block(dp) : knn_f32_queries_per_block(dp);
    p.qblocks = ceil_div(nq, p.qb);
    if (windowed) {
        // partial lists per work unit: [unit][rank][qb]
        p.splits = 1;
        p.max_units = (1 + (int64_t)knn_unit_factor()) * p.qblocks;
        p.d32_bytes = align_up((size_t)p.max_units * p.kcap * p.qb * sizeof(float), 256);
        p.idx_bytes = align_up((size_t)p.max_units * p.kcap * p.qb * sizeof(int32_t), 256);
        p.seed_bytes = align_up((size_t)nq * sizeof(float), 256);
        const int64_t n_tiles = ceil_div(n, FAST_TILE);
        p.plan_capacity = p.qblocks * (n_tiles < knn_plan_cap() ? n_tiles : (int64_t)knn_plan_cap());
        if (p.plan_capacity > 0x7fff0000ll) p.plan_capacity = 0x7fff0000ll;  // int32 list offsets
        p.plan_bytes = align_up((size_t)p.plan_capacity * 8, 256);
        p.count_bytes = 2 * align_up((size_t)p.qblocks * sizeof(int32_t), 256);  // list lengths + starts
        p.first_bytes = align_up((size_t)2 * (p.qblocks + 1) * sizeof(int32_t), 256);  // starts + counts
        p.ublock_bytes = align_up((size_t)p.max_units * sizeof(int32_t), 256);
        p.meta_bytes = 256;
        p.done_bytes = align_up((size_t)4 * p.qblocks * sizeof(int32_t), 256);  // (four counters per block)
    } else {
        // partial lists per candidate split: [split][rank][nq]
        p.splits = choose_splits(p.qblocks, n, FAST_TILE);
        p.d32_bytes = align_up((size_t)p.splits * p.kcap * nq * sizeof(float), 256);
        p.idx_bytes = align_up((size_t)p.splits * p.kcap * nq * sizeof(int32_t), 256);
    }
    return p;
}

}  // namespace ksg

using namespace ksg;

extern "C" {

// KSG_B200_FUSED_GATHER=0: the four separate kernels behind the key sort (A/B runs)
static bool fused_gather() {
    static const bool on = [] { const char* e = getenv("KSG_B200_FUSED_GATHER"); return !(e && e[0] == '0'); }();
    return on;
}

size_t ksg_prepare_workspace(int64_t n, int dim) {
    if (n <= 0 || dim < 1 || dim > KSG_MAX_DIM) return 0;
    return prepared_layout(n, dim).total;
}

int ksg_prepare_join(const ksg_prepared* P, void* stream) {
    KSG_REQUIRE(P, "ksg_prepare_join: null pointer");
    return wait_axes(P, as_stream2(stream));
}

int ksg_prepare(const double* d_points, int64_t ld, int64_t n, int dim, int primary, void* d_buffer,
                size_t buffer_bytes, ksg_prepared* h_out, void* stream) {
    return ksg_prepare_rows(d_points, ld, n, dim, primary, nullptr, d_buffer, buffer_bytes, h_out, stream);
}

int ksg_prepare_rows(const double* d_points, int64_t ld, int64_t n, int dim, int primary, void* const* h_row_ready,
                     void* d_buffer, size_t buffer_bytes, ksg_prepared* h_out, void* stream) {
    KSG_REQUIRE(d_points && d_buffer && h_out, "ksg_prepare: null pointer");
    KSG_REQUIRE(n > 0 && n < (1ll << 31) - 1024 && ld >= n, "ksg_prepare: need 0 < n < 2^31 and ld >= n");
    KSG_REQUIRE(dim >= 1 && dim <= KSG_MAX_DIM, "ksg_prepare: dim out of range");
    KSG_REQUIRE(primary >= -3 && primary < dim,
                "ksg_prepare: primary must be -3 (curve order, lazy axes), -2 (kd order), -1 (curve order) or a coordinate");
    // primary == -3: curve order with LAZY axes -- instead of the exact coordinate sorts the values of every
    // coordinate are grouped into value buckets (axis_buckets.cuh), which is all ksg_axis_count_buckets needs;
    // ksg_prepare_axes builds the sorted arrays later if somebody does need them.  Needs the early order
    // (else the set is an ordinary primary == -1 one: check axes_mode).
    bool lazy = primary == -3;
    if (lazy) primary = -1;
    const PreparedLayout L = prepared_layout(n, dim);
    if (buffer_bytes < L.total) {
        set_error("ksg_prepare: buffer of %zu bytes needed, %zu given", L.total, buffer_bytes);
        return KSG_EWORKSPACE;
    }
    char* base = (char*)d_buffer;
    ksg_prepared P;
    memset(&P, 0, sizeof(P));
    P.n = n; P.dim = dim; P.dim_padded = L.dim_padded; P.primary = primary; P.ld = L.ld; P.n_padded = L.n_padded;
    P.perm = (int32_t*)(base + L.off_perm);
    P.sorted_f64 = (double*)(base + L.off_sorted);
    P.shadow_f32 = (float*)(base + L.off_shadow);
    P.axis_vals = (double*)(base + L.off_axis_vals);
    P.axis_pos = (int32_t*)(base + L.off_axis_pos);
    P.center = (double*)(base + L.off_center);
    P.maxabs = (double*)(base + L.off_maxabs);
    P.flags = (int32_t*)(base + L.off_flags);
    int32_t* iota = (int32_t*)(base + L.off_iota);
    int32_t* exact_info = (int32_t*)(base + L.off_exact);
    char* cub_base = base + L.off_cub;
    void* cub_tmp = cub_base;
    cudaStream_t st = as_stream2(stream);
    // KSG_B200_EXACT32=0: never declare the fp32 copy exact (A/B of the quantised-data path)
    static const int exact_off = [] { const char* e = getenv("KSG_B200_EXACT32"); return e && e[0] == '0' ? 1 : 0; }();
    const int64_t gather_blocks = ceil_div(L.n_padded, 256);

    KSG_CUDA_TRY(cudaMemsetAsync(P.flags, 0, 64, st));
    SortLanes* lanes = sort_lanes_for_device();
    if (!(primary == -1 && early_order() && lanes)) {  // (early order: curve_key_kernel writes the identity order too)
        iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(iota, n);
        KSG_LAUNCHED2();
    }
    const unsigned nb = (unsigned)ceil_div(n, 256);
    int rc;
    P.tile_box = (float*)(base + L.off_tile_box);
    P.sub_box = (float*)(base + L.off_sub_box);
    if (primary == -1) P.curve_keys = (uint64_t*)(base + L.off_mkeys_out);

    if (!(primary == -1 && early_order() && lanes)) lazy = false;
    const int bbits = bucket_bits(n);
    const int n_buckets = 1 << bbits;
    // (the chunk sums of the bucket scan borrow the stats area: KSG_MAX_DIM * STATS_BLOCKS * 2 doubles)
    if ((int64_t)dim * ceil_div(n_buckets, BUCKET_CHUNK) * 4 > (int64_t)KSG_MAX_DIM * STATS_BLOCKS * 2 * 8) lazy = false;
    if (primary == -1 && early_order() && lanes) {
        // ================= early order: the curve from approximate ranks on `st`, the exact coordinate
        // sorts (+ their remap into the prepared order) on the side streams, joined by P.axes_ready
        P.curve_table = (float*)(base + L.off_curve_table);
        double* stats = (double*)(base + L.off_stats);
        unsigned long long* mkeys = (unsigned long long*)(base + L.off_mkeys);
        unsigned long long* mkeys_out = (unsigned long long*)(base + L.off_mkeys_out);
        int32_t* invperm = (int32_t*)(base + L.off_invperm);
        const int n_side = SORT_LANES - 1;
        if (h_row_ready)
            for (int d = 0; d < dim; ++d)
                if (h_row_ready[d]) KSG_CUDA_TRY(cudaStreamWaitEvent(st, (cudaEvent_t)h_row_ready[d], 0));
        int rank_bits, bits;
        curve_bits(n, dim, &rank_bits, &bits);
        if ((size_t)dim * CURVE_TABLE * sizeof(float) > 48 * 1024)  // (dim > 11: the run-time-dimension kernel; outside any capture)
            KSG_CUDA_TRY(cudaFuncSetAttribute(curve_key_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)((size_t)dim * CURVE_TABLE * sizeof(float))));
        auto early = [&](cudaStream_t st) -> int {
            stats_stage1_kernel<<<dim3(STATS_BLOCKS, dim), 256, 0, st>>>(d_points, ld, n, stats, P.flags);
            KSG_LAUNCHED2();
            curve_table_kernel<<<dim, 512, 0, st>>>(d_points, ld, n, stats, P.curve_table, P.center, P.maxabs);
            KSG_LAUNCHED2();
            const size_t tsmem = (size_t)dim * CURVE_TABLE * sizeof(float);
            // lazy axes: per coordinate d, axis_pos row d = [start: n_buckets + 1 | cursor: n_buckets] (2^bbits < n / 8),
            // the bucket ids live in the (otherwise unused) rank rows, the grouped values in axis_vals
            int32_t* bstart = P.axis_pos;
            int32_t* bcursor = lazy ? P.axis_pos + (n_buckets + 1) : nullptr;
            uint32_t* bucket_id = lazy ? (uint32_t*)(base + L.off_ranks) : nullptr;
            if (lazy)
                for (int d = 0; d < dim; ++d)
                    KSG_CUDA_TRY(cudaMemsetAsync(bcursor + (int64_t)d * L.ld, 0, (size_t)n_buckets * 4, st));
            switch (dim) {
#define KSG_CKEY(D) case D: curve_key_kernel<D><<<nb, 256, tsmem, st>>>(d_points, ld, n, dim, P.curve_table, P.center, bits, mkeys, iota, bbits, bcursor, L.ld, bucket_id, L.ld); break;
                KSG_CKEY(1) KSG_CKEY(2) KSG_CKEY(3) KSG_CKEY(4) KSG_CKEY(5) KSG_CKEY(6) KSG_CKEY(7) KSG_CKEY(8)
#undef KSG_CKEY
                default: curve_key_kernel<0><<<nb, 256, tsmem, st>>>(d_points, ld, n, dim, P.curve_table, P.center, bits, mkeys, iota, bbits, bcursor, L.ld, bucket_id, L.ld);
            }
            KSG_LAUNCHED2();
            // The bucket build (needed by the 1-D counts only) runs on a side stream BESIDE the key sort and the
            // gather (three latency-bound radix passes that leave most SMs idle): forked here, joined at the end of
            // this sequence -- inside a capture both become branches of the same graph.  KSG_B200_BUCKET_FORK=0: in line.
            static const bool fork_buckets = [] { const char* e = getenv("KSG_B200_BUCKET_FORK"); return !(e && e[0] == '0'); }();
            const bool forked = lazy && fork_buckets;
            cudaStream_t bs = forked ? lanes->helper[0] : st;
            if (lazy) {
                if (forked) {
                    KSG_CUDA_TRY(cudaEventRecord(lanes->fork, st));
                    KSG_CUDA_TRY(cudaStreamWaitEvent(bs, lanes->fork, 0));
                }
                const unsigned chunks = (unsigned)ceil_div(n_buckets, BUCKET_CHUNK);
                int32_t* chunk_sum = (int32_t*)stats;  // (the min / max partials are spent: curve_table_kernel has run)
                bucket_sum_kernel<<<dim3(chunks, dim), BUCKET_CHUNK, 0, bs>>>(bcursor, L.ld, n_buckets, chunk_sum, chunks,
                                                                               P.flags + 10);
                KSG_LAUNCHED2();
                bucket_scan_kernel<<<dim3(chunks, dim), BUCKET_CHUNK, 0, bs>>>(bcursor, L.ld, bstart, L.ld, n_buckets,
                                                                                chunk_sum, chunks);
                KSG_LAUNCHED2();
                bucket_scatter_kernel<<<dim3((unsigned)ceil_div(n, 256 * BUCKET_SCATTER_PER_THREAD), dim), 256, 0, bs>>>(
                    d_points, ld, n, bucket_id, L.ld, bcursor, L.ld, P.axis_vals, L.ld);
                KSG_LAUNCHED2();
                if (forked) KSG_CUDA_TRY(cudaEventRecord(lanes->join[0], bs));
            }
            auto join_buckets = [&]() -> int {
                if (forked) KSG_CUDA_TRY(cudaStreamWaitEvent(st, lanes->join[0], 0));
                return KSG_OK;
            };
            if (use_sample_sort(n)) {
                const int rc2 = sample_sort_launch<unsigned long long>(cub_tmp, mkeys, mkeys_out, P.perm, n, st);
                if (rc2) return rc2;
            } else {
                size_t need = L.cub_bytes;
                cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, need, mkeys, mkeys_out, iota, P.perm, (int)n, 0,
                                                                bits * dim, st);
                if (e != cudaSuccess) { set_error("cub SortPairs (curve): %s", cudaGetErrorString(e)); return KSG_ECUDA; }
                note_launch();
            }
            if (fused_gather()) {
                // (the coordinate sorts of this path run on the prepared copy: no inverse permutation, no remap)
                gather_tile_kernel<<<(unsigned)(L.n_padded / FAST_TILE), 256, 0, st>>>(
                    d_points, ld, n, dim, L.dim_padded, P.perm, P.center, P.sorted_f64, L.ld, P.shadow_f32,
                    nullptr, P.tile_box, P.sub_box, exact_info, exact_off, P.flags);
                KSG_LAUNCHED2();
                return join_buckets();
            }
            gather_sorted_kernel<<<(unsigned)gather_blocks, 256, 0, st>>>(
                d_points, ld, n, dim, L.dim_padded, P.perm, P.center, P.sorted_f64, L.ld, P.shadow_f32, L.n_padded,
                exact_info);
            KSG_LAUNCHED2();
            exact_flag_kernel<<<1, 256, 0, st>>>(exact_info, gather_blocks, dim, exact_off, P.flags);
            KSG_LAUNCHED2();
            invert_perm_kernel<<<nb, 256, 0, st>>>(P.perm, n, invperm);
            KSG_LAUNCHED2();
            tile_box_kernel<<<(unsigned)(L.n_padded / FAST_TILE), 128, 0, st>>>(P.shadow_f32, L.dim_padded, P.tile_box,
                                                                                 P.sub_box);
            KSG_LAUNCHED2();
            return join_buckets();
        };
        rc = graph_cache().run(graph_key(GRAPH_PREPARE_EARLY, {d_points, d_buffer}, {n, ld},
                                         {dim, primary, fused_gather() ? 1 : 0, lazy ? 1 : 0}), st, early);
        if (rc) return rc;
        if (lazy) {  // no coordinate sorts at all (ksg_prepare_axes builds them on demand)
            P.axes_mode = bbits;
            P.axes_ready = nullptr;
            *h_out = P;
            return KSG_OK;
        }
        KSG_CUDA_TRY(cudaEventRecord(lanes->perm_ready, st));
        // The exact coordinate sorts start HERE, behind the curve path: three radix sorts side by side
        // slow each other down (every pass is latency bound and they share the SMs: measured, the curve
        // path took 2.5x as long with the coordinate sorts running beside it), whereas the k-NN kernels
        // the caller enqueues next leave issue slots free.
        // They read the set's OWN float64 copy (prepared order), not the caller's array: the caller may free
        // or overwrite d_points as soon as this function has returned and its stream work is ordered -- the
        // side streams are not -- and the payload is then the prepared position itself (no remap pass; ties
        // come out in prepared order).
        for (int l = 0; l < n_side && l < dim; ++l) KSG_CUDA_TRY(cudaStreamWaitEvent(lanes->helper[l], lanes->perm_ready, 0));
        for (int d = 0; d < dim; ++d) {
            cudaStream_t ls = lanes->helper[d % n_side];
            void* lane_tmp = cub_base + (size_t)(1 + d % n_side) * L.cub_bytes;  // (area 0: the key sort on `st`)
            const double* keys_in = P.sorted_f64 + (int64_t)d * L.ld;
            double* keys_out = P.axis_vals + (int64_t)d * L.ld;
            int32_t* pos_out = P.axis_pos + (int64_t)d * L.ld;
            rc = graph_cache().run(graph_key(GRAPH_AXIS_SORT, {keys_in, keys_out, iota, pos_out, lane_tmp}, {n}, {}), ls,
                                   [&](cudaStream_t s2) {
                                       return sort_pairs(lane_tmp, L.cub_bytes, keys_in, keys_out, iota, pos_out, n, s2);
                                   });
            if (rc) return rc;
        }
        cudaStream_t s0 = lanes->helper[0];
        for (int l = 1; l < n_side && l < dim; ++l) {
            KSG_CUDA_TRY(cudaEventRecord(lanes->join[l], lanes->helper[l]));
            KSG_CUDA_TRY(cudaStreamWaitEvent(s0, lanes->join[l], 0));
        }
        cudaEvent_t done = lanes->axes_ready[lanes->next_axes++ % AXES_EVENTS];
        KSG_CUDA_TRY(cudaEventRecord(done, s0));
        P.axes_ready = (void*)done;
        *h_out = P;
        return KSG_OK;
    }

    // ================= exact order: every coordinate sorted first (kd order, single-coordinate order,
    // KSG_B200_EARLY_ORDER=0)
    // every coordinate sorted on its own: values + ORIGINAL sample index (remapped below);
    // coordinate d runs on lane d % sort_lanes (see SortLanes)
    if (L.sort_lanes <= 1) lanes = nullptr;
    const int n_lanes = lanes ? L.sort_lanes : 1;
    if (n_lanes > 1) {
        KSG_CUDA_TRY(cudaEventRecord(lanes->fork, st));
        for (int l = 1; l < n_lanes; ++l) KSG_CUDA_TRY(cudaStreamWaitEvent(lanes->helper[l - 1], lanes->fork, 0));
    }
    for (int d = 0; d < dim; ++d) {
        const int lane = d % n_lanes;
        if (h_row_ready && h_row_ready[d])
            KSG_CUDA_TRY(cudaStreamWaitEvent(lane == 0 ? st : lanes->helper[lane - 1], (cudaEvent_t)h_row_ready[d], 0));
        void* lane_tmp = cub_base + (size_t)lane * L.cub_bytes;
        const double* keys_in = d_points + (int64_t)d * ld;
        double* keys_out = P.axis_vals + (int64_t)d * L.ld;
        int32_t* pos_out = P.axis_pos + (int64_t)d * L.ld;
        rc = graph_cache().run(graph_key(GRAPH_AXIS_SORT, {keys_in, keys_out, iota, pos_out, lane_tmp}, {n}, {}),
                               lane == 0 ? st : lanes->helper[lane - 1], [&](cudaStream_t s) {
                                   return sort_pairs(lane_tmp, L.cub_bytes, keys_in, keys_out, iota, pos_out, n, s);
                               });
        if (rc) return rc;
    }
    for (int l = 1; l < n_lanes; ++l) {
        KSG_CUDA_TRY(cudaEventRecord(lanes->join[l - 1], lanes->helper[l - 1]));
        KSG_CUDA_TRY(cudaStreamWaitEvent(st, lanes->join[l - 1], 0));
    }
    // everything behind the coordinate sorts, as one cached launch sequence
    auto tail = [&](cudaStream_t st) -> int {
        // centre of the fp32 copy = per-coordinate MEDIAN (the bulk of the points then has small centred
        // magnitudes, hence small fp32 error bands, even when outliers stretch the range); the ends of
        // the sorted arrays also tell whether the input held a NaN or an infinity (flags[0])
        center_median_kernel<<<1, KSG_MAX_DIM, 0, st>>>(P.axis_vals, L.ld, n, dim, P.center, P.maxabs, P.flags);
        KSG_LAUNCHED2();
        if (primary >= 0) {
            // order = one coordinate's sorted order
            KSG_CUDA_TRY(cudaMemcpyAsync(P.perm, P.axis_pos + (int64_t)primary * L.ld, (size_t)n * 4,
                                         cudaMemcpyDeviceToDevice, st));
        } else if (primary == -2) {
            // order = leaves of a balanced kd-tree over the per-coordinate ranks (see kd_key_kernel)
            int32_t* ranks = (int32_t*)(base + L.off_ranks);
            unsigned long long* mkeys = (unsigned long long*)(base + L.off_mkeys);
            unsigned long long* mkeys_out = (unsigned long long*)(base + L.off_mkeys_out);
            rank_scatter_kernel<<<dim3(nb, dim), 256, 0, st>>>(P.axis_pos, L.ld, n, ranks);
            KSG_LAUNCHED2();
            const int64_t n_groups = ceil_div(n, FAST_SUB);
            int levels = 0;
            while ((1ll << levels) < n_groups) ++levels;
            int32_t* pa = iota;    // holds 0..n-1: the identity order
            int32_t* pb = P.perm;
            for (int level = 0; level < levels; ++level) {
                kd_key_kernel<<<nb, 256, 0, st>>>(pa, ranks + (int64_t)(level % dim) * L.ld, n, level, n_groups, mkeys);
                KSG_LAUNCHED2();
                size_t need = L.cub_bytes;
                cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, need, mkeys, mkeys_out, pa, pb, (int)n, 0,
                                                                32 + level, st);
                if (e != cudaSuccess) { set_error("cub SortPairs (kd): %s", cudaGetErrorString(e)); return KSG_ECUDA; }
                note_launch();
                int32_t* t = pa; pa = pb; pb = t;
            }
            if (pa != P.perm)
                KSG_CUDA_TRY(cudaMemcpyAsync(P.perm, pa, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        } else {
            // order = Hilbert curve over the per-coordinate ranks
            int32_t* ranks = (int32_t*)(base + L.off_ranks);
            unsigned long long* mkeys = (unsigned long long*)(base + L.off_mkeys);
            unsigned long long* mkeys_out = (unsigned long long*)(base + L.off_mkeys_out);
            rank_scatter_kernel<<<dim3(nb, dim), 256, 0, st>>>(P.axis_pos, L.ld, n, ranks);
            KSG_LAUNCHED2();
            int rank_bits, bits;
            curve_bits(n, dim, &rank_bits, &bits);
            switch (dim) {
#define KSG_MORTON(D) case D: morton_key_kernel<D><<<nb, 256, 0, st>>>(ranks, L.ld, n, dim, rank_bits, bits, mkeys); break;
                KSG_MORTON(1) KSG_MORTON(2) KSG_MORTON(3) KSG_MORTON(4) KSG_MORTON(5) KSG_MORTON(6) KSG_MORTON(7) KSG_MORTON(8)
#undef KSG_MORTON
                default: morton_key_kernel<0><<<nb, 256, 0, st>>>(ranks, L.ld, n, dim, rank_bits, bits, mkeys);
            }
            KSG_LAUNCHED2();
            if (use_sample_sort(n)) {
                const int rc2 = sample_sort_launch<unsigned long long>(cub_tmp, mkeys, mkeys_out, P.perm, n, st);
                if (rc2) return rc2;
            } else {
                size_t need = L.cub_bytes;
                cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, need, mkeys, mkeys_out, iota, P.perm, (int)n, 0,
                                                                bits * dim, st);
                if (e != cudaSuccess) { set_error("cub SortPairs (morton): %s", cudaGetErrorString(e)); return KSG_ECUDA; }
                note_launch();
            }
        }
        // axis_pos: original sample index -> position in the final order
        int32_t* invperm = (int32_t*)(base + L.off_invperm);
        if (fused_gather()) {
            gather_tile_kernel<<<(unsigned)(L.n_padded / FAST_TILE), 256, 0, st>>>(
                d_points, ld, n, dim, L.dim_padded, P.perm, P.center, P.sorted_f64, L.ld, P.shadow_f32, invperm,
                P.tile_box, P.sub_box, exact_info, exact_off, P.flags);
            KSG_LAUNCHED2();
            remap_axis_kernel<<<dim3(nb, dim), 256, 0, st>>>(P.axis_pos, L.ld, n, invperm);
            KSG_LAUNCHED2();
            return KSG_OK;
        }
        gather_sorted_kernel<<<(unsigned)gather_blocks, 256, 0, st>>>(
            d_points, ld, n, dim, L.dim_padded, P.perm, P.center, P.sorted_f64, L.ld, P.shadow_f32, L.n_padded,
            exact_info);
        KSG_LAUNCHED2();
        exact_flag_kernel<<<1, 256, 0, st>>>(exact_info, gather_blocks, dim, exact_off, P.flags);
        KSG_LAUNCHED2();
        invert_perm_kernel<<<nb, 256, 0, st>>>(P.perm, n, invperm);
        KSG_LAUNCHED2();
        remap_axis_kernel<<<dim3(nb, dim), 256, 0, st>>>(P.axis_pos, L.ld, n, invperm);
        KSG_LAUNCHED2();
        tile_box_kernel<<<(unsigned)(L.n_padded / FAST_TILE), 128, 0, st>>>(P.shadow_f32, L.dim_padded, P.tile_box,
                                                                             P.sub_box);
        KSG_LAUNCHED2();
        return KSG_OK;
    };
    rc = graph_cache().run(graph_key(GRAPH_PREPARE_TAIL, {d_points, d_buffer}, {n, ld}, {dim, primary, fused_gather() ? 1 : 0}), st, tail);
    if (rc) return rc;
    *h_out = P;
    return KSG_OK;
}

size_t ksg_fast_knn_workspace(int64_t n, int64_t nq, int dim, int kprime) {
    return ksg_fast_knn_workspace_margin(n, nq, dim, kprime, KNN_MARGIN);
}

size_t ksg_fast_knn_workspace_margin(int64_t n, int64_t nq, int dim, int kprime, int margin) {
    if (n <= 0 || nq <= 0 || dim < 1 || kprime < 1 || margin < 1) return 0;
    // either engine may be asked for with this workspace
    const size_t dense = fast_knn_plan(n, nq, dim, kprime, 0, margin).total();
    const size_t windowed = fast_knn_plan(n, nq, dim, kprime, 1, margin).total();
    return (dense > windowed ? dense : windowed) + 256;
}

int ksg_fast_knn(const ksg_prepared* P, int64_t q_begin, int64_t q_end, const int32_t* d_query_index, int kprime,
                 int windowed, double* d_eps_out, int32_t* d_flags, int32_t* d_n_flagged, void* d_workspace,
                 size_t workspace_bytes, void* stream) {
    return ksg_fast_knn_idx(P, q_begin, q_end, d_query_index, kprime, windowed, d_eps_out, nullptr, d_flags,
                            d_n_flagged, d_workspace, workspace_bytes, stream);
}

int ksg_fast_knn_idx(const ksg_prepared* P, int64_t q_begin, int64_t q_end, const int32_t* d_query_index, int kprime,
                     int windowed, double* d_eps_out, int64_t* d_idx_out, int32_t* d_flags, int32_t* d_n_flagged,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
    return ksg_fast_knn_margin(P, q_begin, q_end, d_query_index, kprime, KNN_MARGIN, windowed, d_eps_out, d_idx_out,
                               d_flags, d_n_flagged, d_workspace, workspace_bytes, stream);
}

int ksg_fast_knn_margin(const ksg_prepared* P, int64_t q_begin, int64_t q_end, const int32_t* d_query_index, int kprime,
                        int margin, int windowed, double* d_eps_out, int64_t* d_idx_out, int32_t* d_flags,
                        int32_t* d_n_flagged, void* d_workspace, size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(P && d_eps_out && d_flags && d_n_flagged && d_workspace, "ksg_fast_knn: null pointer");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && (d_query_index || q_end <= P->n), "ksg_fast_knn: bad query range");
    KSG_REQUIRE(kprime >= 1 && margin >= 1 && kprime + margin <= KNN_MAX_KCAP,
                "ksg_fast_knn: k' + margin out of range (1 <= margin, k' + margin <= 64)");
    if (d_query_index) windowed = 0;  // an explicit query list has no spatial coherence to prune with
    if (P->dim_padded > 8) {
        set_error("ksg_fast_knn: the fp32 filter is instantiated for up to 8 dimensions (got %d); use ksg_knn_radius",
                  P->dim);
        return KSG_EUNSUPPORTED;
    }
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    FastKnnPlan p = fast_knn_plan(P->n, nq, P->dim, kprime, windowed, margin);
    if (workspace_bytes < p.total()) {
        set_error("ksg_fast_knn: workspace of %zu bytes needed, %zu given", p.total(), workspace_bytes);
        return KSG_EWORKSPACE;
    }
    cudaStream_t st = as_stream2(stream);
    KnnF32Args a;
    a.shadow = P->shadow_f32; a.tile_box = P->tile_box; a.sub_box = P->sub_box; a.n_tiles = P->n_padded / FAST_TILE;
    a.q_begin = q_begin; a.nq = nq; a.q_index = d_query_index; a.kcap = p.kcap; a.splits = p.splits;
    a.windowed = windowed ? 1 : 0;
    a.maxabs = P->maxabs; a.dim = P->dim;
    a.qshadow = nullptr; a.qhome = nullptr;
    a.part_d32 = (float*)d_workspace;
    a.part_idx = (int32_t*)((char*)d_workspace + p.d32_bytes);
    a.tile_counter = (unsigned long long*)((char*)P->flags + 16);
    a.st = st;
    a.seed_thr = nullptr; a.plan = nullptr; a.plan_count = nullptr; a.plan_first = nullptr; a.plan_capacity = 0;
    a.unit_first = nullptr; a.unit_block = nullptr; a.unit_meta = nullptr; a.unit_done = nullptr; a.max_units = 0;
    fill_refine_args(a, P, kprime, nullptr, 0, P->maxabs, d_idx_out, d_eps_out, d_flags, d_n_flagged);
    int rc;
    if (a.windowed) {
        // planning pass (seeds, deferrals, per-block tile lists) -> unit table -> persistent filter
        char* w = (char*)d_workspace + p.d32_bytes + p.idx_bytes;
        a.seed_thr = (float*)w;                      w += p.seed_bytes;
        a.plan = w;                                  w += p.plan_bytes;
        a.plan_capacity = p.plan_capacity;
        a.plan_count = (int32_t*)w;                  w += p.count_bytes / 2;
        a.plan_first = (int32_t*)w;                  w += p.count_bytes / 2;
        a.unit_first = (int32_t*)w;                  w += p.first_bytes;
        a.unit_block = (int32_t*)w;                  w += p.ublock_bytes;
        a.unit_meta = (int32_t*)w;                   w += p.meta_bytes;
        a.unit_done = (int32_t*)w;
        a.max_units = p.max_units;
        rc = launch_knn_plan(P->dim_padded, a);
        if (rc) return rc;
        rc = launch_knn_units(P->dim_padded, a);
        if (rc) return rc;
        if (a.ra.fused) return wait_axes(P, st);  // fused: every block was refined inside the units kernel
    } else {
        rc = launch_knn_f32(P->dim_padded, a);
    }
    if (rc) return rc;
    knn_refine_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, st>>>(a.part_d32, a.part_idx, p.kcap, p.splits, kprime,
                                                                   P->sorted_f64, P->ld, P->dim, P->n, q_begin, nq,
                                                                   d_query_index, P->maxabs, a.seed_thr, a.unit_first,
                                                                   p.qb, P->perm, d_idx_out, nullptr, 0, P->center,
                                                                   d_eps_out, d_flags, d_n_flagged, a.ra.exact32);
    KSG_LAUNCHED2();
    // from here on the caller's stream is also behind the side-stream work of the prepared set (the
    // exact coordinate sorts ran underneath the kernels above)
    return wait_axes(P, st);
}

// ---------------------------------------------------------------- cross-set queries
struct CrossLayout {
    int64_t ld;
    size_t off_perm, off_sorted, off_shadow, off_home, off_maxabs, off_keys, off_keys_out, off_iota, off_stats, off_cub;
    size_t cub_bytes, total;
};
static CrossLayout cross_layout(int64_t nq, int dim) {
    CrossLayout L;
    L.ld = (int64_t)align_up((size_t)nq, 32);
    const int dp = (dim + 1) & ~1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    L.off_perm = take((size_t)L.ld * 4);
    L.off_sorted = take((size_t)dim * L.ld * 8);
    L.off_shadow = take((size_t)L.ld * dp * 4);
    L.off_home = take((size_t)L.ld * 4);
    L.off_maxabs = take((size_t)KSG_MAX_DIM * 8);
    L.off_keys = take((size_t)L.ld * 8);
    L.off_keys_out = take((size_t)L.ld * 8);
    L.off_iota = take((size_t)L.ld * 4);
    L.off_stats = take((size_t)dim * STATS_BLOCKS * 2 * 8 + 64);
    L.cub_bytes = cub_sort_bytes(nq);
    L.off_cub = take(L.cub_bytes);
    L.total = o;
    return L;
}

size_t ksg_cross_prepare_workspace(int64_t n_candidates, int64_t nq, int dim) {
    (void)n_candidates;
    if (nq <= 0 || dim < 1 || dim > KSG_MAX_DIM) return 0;
    return cross_layout(nq, dim).total;
}

int ksg_cross_prepare(const ksg_prepared* P, const double* d_queries, int64_t ldq, int64_t nq, void* d_buffer,
                      size_t buffer_bytes, ksg_cross* h_out, void* stream) {
    KSG_REQUIRE(P && d_queries && d_buffer && h_out, "ksg_cross_prepare: null pointer");
    KSG_REQUIRE(nq > 0 && nq < (1ll << 31) - 1024 && ldq >= nq, "ksg_cross_prepare: need 0 < nq < 2^31 and ldq >= nq");
    KSG_REQUIRE(P->primary == -1 && P->curve_keys, "ksg_cross_prepare: the candidate set must be in curve order (primary = -1)");
    const int dim = P->dim;
    const CrossLayout L = cross_layout(nq, dim);
    if (buffer_bytes < L.total) {
        set_error("ksg_cross_prepare: buffer of %zu bytes needed, %zu given", L.total, buffer_bytes);
        return KSG_EWORKSPACE;
    }
    char* base = (char*)d_buffer;
    ksg_cross X;
    memset(&X, 0, sizeof(X));
    X.nq = nq; X.ld = L.ld;
    X.perm = (int32_t*)(base + L.off_perm);
    X.sorted_f64 = (double*)(base + L.off_sorted);
    X.shadow_f32 = (float*)(base + L.off_shadow);
    X.home = (int32_t*)(base + L.off_home);
    X.maxabs = (double*)(base + L.off_maxabs);
    unsigned long long* keys = (unsigned long long*)(base + L.off_keys);
    unsigned long long* keys_out = (unsigned long long*)(base + L.off_keys_out);
    int32_t* iota = (int32_t*)(base + L.off_iota);
    double* stats = (double*)(base + L.off_stats);
    int32_t* scratch_flag = (int32_t*)(base + L.off_stats + (size_t)dim * STATS_BLOCKS * 2 * 8);
    cudaStream_t st = as_stream2(stream);
    const unsigned nb = (unsigned)ceil_div(nq, 256);
    int rank_bits, bits;
    curve_bits(P->n, dim, &rank_bits, &bits);
    if (P->curve_table) {
        // the candidates' curve was built from their coordinate table: place the queries with the same map
        const size_t tsmem = (size_t)dim * CURVE_TABLE * sizeof(float);
        if (tsmem > 48 * 1024) KSG_CUDA_TRY(cudaFuncSetAttribute(curve_key_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        curve_key_kernel<0><<<nb, 256, tsmem, st>>>(d_queries, ldq, nq, dim, P->curve_table, P->center, bits, keys, iota, 0, nullptr, 0, nullptr, 0);
    } else
    switch (dim) {
#define KSG_CROSS(D) case D: cross_key_kernel<D><<<nb, 256, 0, st>>>(d_queries, ldq, nq, P->axis_vals, P->ld, P->n, dim, rank_bits, bits, keys, iota); break;
        KSG_CROSS(1) KSG_CROSS(2) KSG_CROSS(3) KSG_CROSS(4) KSG_CROSS(5) KSG_CROSS(6) KSG_CROSS(7) KSG_CROSS(8)
#undef KSG_CROSS
        default: cross_key_kernel<0><<<nb, 256, 0, st>>>(d_queries, ldq, nq, P->axis_vals, P->ld, P->n, dim, rank_bits, bits, keys, iota);
    }
    KSG_LAUNCHED2();
    if (use_sample_sort(nq)) {
        const int rc2 = sample_sort_launch<unsigned long long>(base + L.off_cub, keys, keys_out, X.perm, nq, st);
        if (rc2) return rc2;
    } else {
        size_t need = L.cub_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(base + L.off_cub, need, keys, keys_out, iota, X.perm, (int)nq, 0,
                                                        bits * dim, st);
        if (e != cudaSuccess) { set_error("cub SortPairs (cross): %s", cudaGetErrorString(e)); return KSG_ECUDA; }
        note_launch();
    }
    cross_gather_kernel<<<nb, 256, 0, st>>>(d_queries, ldq, nq, dim, P->dim_padded, X.perm, keys_out, P->center,
                                            (const unsigned long long*)P->curve_keys, P->n, X.sorted_f64, L.ld,
                                            X.shadow_f32, X.home);
    KSG_LAUNCHED2();
    KSG_CUDA_TRY(cudaMemsetAsync(scratch_flag, 0, 64, st));
    stats_stage1_kernel<<<dim3(STATS_BLOCKS, dim), 256, 0, st>>>(d_queries, ldq, nq, stats, scratch_flag);
    KSG_LAUNCHED2();
    cross_maxabs_kernel<<<1, KSG_MAX_DIM, 0, st>>>(stats, dim, P->center, P->maxabs, X.maxabs);
    KSG_LAUNCHED2();
    *h_out = X;
    return KSG_OK;
}

int ksg_fast_knn_cross(const ksg_prepared* P, const ksg_cross* X, int64_t q_begin, int64_t q_end, int kprime,
                       double* d_eps_out, int32_t* d_flags, int32_t* d_n_flagged, void* d_workspace,
                       size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(P && X && d_eps_out && d_flags && d_n_flagged && d_workspace, "ksg_fast_knn_cross: null pointer");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= X->nq, "ksg_fast_knn_cross: bad query range");
    KSG_REQUIRE(kprime >= 1 && kprime + KNN_MARGIN <= KNN_MAX_KCAP, "ksg_fast_knn_cross: kprime out of range (k' + 3 <= 64)");
    if (P->dim_padded > 8) {
        set_error("ksg_fast_knn_cross: the fp32 filter is instantiated for up to 8 dimensions (got %d)", P->dim);
        return KSG_EUNSUPPORTED;
    }
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    FastKnnPlan p = fast_knn_plan(P->n, nq, P->dim, kprime, 1);
    if (workspace_bytes < p.total()) {
        set_error("ksg_fast_knn_cross: workspace of %zu bytes needed, %zu given", p.total(), workspace_bytes);
        return KSG_EWORKSPACE;
    }
    cudaStream_t st = as_stream2(stream);
    KnnF32Args a;
    a.shadow = P->shadow_f32; a.tile_box = P->tile_box; a.sub_box = P->sub_box; a.n_tiles = P->n_padded / FAST_TILE;
    a.q_begin = q_begin; a.nq = nq; a.q_index = nullptr; a.kcap = p.kcap; a.splits = 1; a.windowed = 1;
    a.maxabs = X->maxabs; a.dim = P->dim;
    a.qshadow = X->shadow_f32; a.qhome = X->home;
    a.part_d32 = (float*)d_workspace;
    a.part_idx = (int32_t*)((char*)d_workspace + p.d32_bytes);
    a.tile_counter = (unsigned long long*)((char*)P->flags + 16);
    a.st = st;
    char* w = (char*)d_workspace + p.d32_bytes + p.idx_bytes;
    a.seed_thr = (float*)w;                      w += p.seed_bytes;
    a.plan = w;                                  w += p.plan_bytes;
    a.plan_capacity = p.plan_capacity;
    a.plan_count = (int32_t*)w;                  w += p.count_bytes / 2;
    a.plan_first = (int32_t*)w;                  w += p.count_bytes / 2;
    a.unit_first = (int32_t*)w;                  w += p.first_bytes;
    a.unit_block = (int32_t*)w;                  w += p.ublock_bytes;
    a.unit_meta = (int32_t*)w;                   w += p.meta_bytes;
    a.unit_done = (int32_t*)w;
    a.max_units = p.max_units;
    fill_refine_args(a, P, kprime, X->sorted_f64, X->ld, X->maxabs, nullptr, d_eps_out, d_flags, d_n_flagged);
    int rc = launch_knn_plan(P->dim_padded, a);
    if (rc) return rc;
    rc = launch_knn_units(P->dim_padded, a);
    if (rc) return rc;
    if (a.ra.fused) return wait_axes(P, st);
    knn_refine_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, st>>>(a.part_d32, a.part_idx, p.kcap, 1, kprime,
                                                                   P->sorted_f64, P->ld, P->dim, P->n, q_begin, nq,
                                                                   nullptr, X->maxabs, a.seed_thr, a.unit_first, p.qb,
                                                                   P->perm, nullptr, X->sorted_f64, X->ld, P->center,
                                                                   d_eps_out, d_flags, d_n_flagged, nullptr);
    KSG_LAUNCHED2();
    return wait_axes(P, st);
}

struct FastCountPlan {
    int64_t qblocks, max_units, plan_capacity;
    size_t thr_bytes, plan_bytes, count_bytes, first_bytes, ublock_bytes, meta_bytes;
    size_t total() const { return thr_bytes + plan_bytes + count_bytes + first_bytes + ublock_bytes + meta_bytes; }
};
static FastCountPlan fast_count_plan(int64_t n, int64_t nq, int dim) {
    FastCountPlan p;
    memset(&p, 0, sizeof(p));
    if (nq < 1) nq = 1;
    const int dp = (dim + 1) & ~1;
    p.thr_bytes = align_up((size_t)nq * sizeof(float), 256);
    (void)dp;
    p.qblocks = ceil_div(nq, (int64_t)FAST_THREADS * COUNT_UNITS_Q);  // (the work-unit path's query blocks)
    p.max_units = (1 + (int64_t)knn_unit_factor()) * p.qblocks;
    const int64_t n_tiles = ceil_div(n, FAST_TILE);
    p.plan_capacity = p.qblocks * (n_tiles < knn_plan_cap() ? n_tiles : (int64_t)knn_plan_cap());
    if (p.plan_capacity > 0x7fff0000ll) p.plan_capacity = 0x7fff0000ll;
    p.plan_bytes = align_up((size_t)p.plan_capacity * 8, 256);
    p.count_bytes = 2 * align_up((size_t)p.qblocks * sizeof(int32_t), 256);
    p.first_bytes = align_up((size_t)2 * (p.qblocks + 1) * sizeof(int32_t), 256);  // starts + counts
    p.ublock_bytes = align_up((size_t)p.max_units * sizeof(int32_t), 256);
    p.meta_bytes = 256;
    return p;
}

size_t ksg_fast_count_workspace(int64_t n, int64_t nq, int dim, int n_subspaces) {
    (void)n_subspaces;
    if (n <= 0 || dim < 1) return 0;
    return fast_count_plan(n, nq, dim).total();
}

int ksg_fast_count(const ksg_prepared* P, int64_t q_begin, int64_t q_end, const uint32_t* h_masks,
                   int n_subspaces, const double* d_eps, int strict, int32_t bias, int windowed,
                   int32_t* d_counts, int32_t* d_flags, int32_t* d_n_flagged, void* d_workspace,
                   size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(P && h_masks && d_eps && d_counts && d_flags && d_n_flagged && d_workspace,
                "ksg_fast_count: null pointer");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= P->n, "ksg_fast_count: bad query range");
    KSG_REQUIRE(n_subspaces >= 1 && n_subspaces <= KSG_MAX_SUBSPACES, "ksg_fast_count: subspace count");
    KSG_REQUIRE(P->axes_mode == 0, "ksg_fast_count: the set has lazy axes (call ksg_prepare_axes first)");
    if (P->dim_padded > 8) {
        set_error("ksg_fast_count: the fp32 filter is instantiated for up to 8 dimensions (got %d)", P->dim);
        return KSG_EUNSUPPORTED;
    }
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    const FastCountPlan cp = fast_count_plan(P->n, nq, P->dim);
    if (workspace_bytes < cp.total()) {
        set_error("ksg_fast_count: workspace of %zu bytes needed, %zu given", cp.total(), workspace_bytes);
        return KSG_EWORKSPACE;
    }
    const int dim = P->dim;
    const uint32_t all = dim == 32 ? 0xffffffffu : ((1u << dim) - 1u);
    for (int s = 0; s < n_subspaces; ++s) {
        KSG_REQUIRE(h_masks[s] != 0 && (h_masks[s] & ~all) == 0, "ksg_fast_count: empty or out-of-range mask");
    }
    cudaStream_t st = as_stream2(stream);
    {
        const int rc0 = wait_axes(P, st);  // (the band resolution reads axis_vals / axis_pos)
        if (rc0) return rc0;
    }
    float* thresholds = (float*)d_workspace;
    const int64_t total = (int64_t)n_subspaces * nq;
    fill_i32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(d_counts, total, bias);
    KSG_LAUNCHED2();
    KSG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)nq * sizeof(int32_t), st));
    count_thresholds_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(d_eps, q_begin, nq, P->shadow_f32,
                                                                          P->dim_padded, strict ? 1 : 0, P->flags + 8,
                                                                          thresholds);
    KSG_LAUNCHED2();

    CountF32Args a;
    memset(&a, 0, sizeof(a));
    a.shadow = P->shadow_f32; a.tile_box = P->tile_box; a.n_tiles = P->n_padded / FAST_TILE;
    a.q_begin = q_begin; a.nq = nq; a.thresholds = thresholds;
    a.windowed = windowed ? 1 : 0;
    a.n_sub = n_subspaces; a.counts = d_counts; a.st = st;
    a.tile_counter = (unsigned long long*)((char*)P->flags + 24);
    a.sub_box = P->sub_box;
    const int64_t qblocks = ceil_div(nq, count_f32_queries_per_block(P->dim_padded));
    a.splits = choose_splits(qblocks, P->n, FAST_TILE);  // (the pruned walk interleaves the splits too)

    // recognise the fused structures
    int rc = KSG_EUNSUPPORTED;
    auto low = [](int bits) { return bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u); };
    if (n_subspaces == 1 && h_masks[0] == all && a.windowed) {
        // the subspace is the whole space of this prepared set: planned, unit-balanced pass
        CountUnitsArgs w;
        char* p = (char*)d_workspace + cp.thr_bytes;
        w.sub_box = P->sub_box;
        w.plan = p;                        p += cp.plan_bytes;
        w.plan_capacity = cp.plan_capacity;
        w.plan_count = (int32_t*)p;        p += cp.count_bytes / 2;
        w.plan_first = (int32_t*)p;        p += cp.count_bytes / 2;
        w.unit_first = (int32_t*)p;        p += cp.first_bytes;
        w.unit_block = (int32_t*)p;        p += cp.ublock_bytes;
        w.unit_meta = (int32_t*)p;
        w.max_units = cp.max_units;
        rc = launch_count_full(P->dim_padded, a, w);
    } else if (n_subspaces == 2) {
        const int dx = __builtin_popcount(h_masks[0]), dy = __builtin_popcount(h_masks[1]);
        if (dx + dy == dim && h_masks[0] == low(dx) && h_masks[1] == (low(dy) << dx)) rc = launch_count_pair(dx, dy, a);
    } else if (n_subspaces == 3) {
        const int dz = __builtin_popcount(h_masks[0]);
        const int dx = __builtin_popcount(h_masks[1]) - dz, dy = __builtin_popcount(h_masks[2]) - dz;
        if (dx >= 1 && dy >= 1 && dx + dy + dz == dim) {
            const uint32_t mz = low(dz) << (dx + dy), mx = low(dx), my = low(dy) << dx;
            if (h_masks[0] == mz && h_masks[1] == (mx | mz) && h_masks[2] == (my | mz)) rc = launch_count_cmi(dx, dy, dz, a);
        }
    }
    if (rc == KSG_EUNSUPPORTED && n_subspaces == dim && dim >= 3) {
        bool loo = true;
        for (int s = 0; s < dim; ++s) loo = loo && (h_masks[s] == (all & ~(1u << s)));
        if (loo) {
            // 8-query groups + sub-tile culling on the second-largest box gap (fast_count_loo_group.cu); small or
            // low-dimensional cases keep the block-level kernel
            rc = launch_count_loo_grouped(dim, a);
            if (rc == KSG_EUNSUPPORTED) rc = launch_count_loo(dim, a);
        }
    }
    if (rc == KSG_EUNSUPPORTED) {
        // generic: four run-time masks per launch
        for (int s0 = 0; s0 < n_subspaces; s0 += 4) {
            CountF32Args b = a;
            b.n_sub = n_subspaces - s0 < 4 ? n_subspaces - s0 : 4;
            for (int i = 0; i < 4; ++i) b.masks[i] = i < b.n_sub ? h_masks[s0 + i] : 0u;
            b.counts = d_counts + (int64_t)s0 * nq;
            rc = launch_count_masked(P->dim_padded, b);
            if (rc) return rc;
        }
    }
    if (rc) return rc;

    MaskTable64 mt;
    memset(&mt, 0, sizeof(mt));
    for (int s = 0; s < n_subspaces; ++s) mt.m[s] = h_masks[s];
    count_fix_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, st>>>(
        P->sorted_f64, P->ld, P->shadow_f32, dim, P->dim_padded, P->n, P->axis_vals, P->axis_pos, P->maxabs, q_begin,
        nq, d_eps, thresholds, mt, n_subspaces, strict ? 1 : 0, d_counts, d_flags, d_n_flagged, P->flags + 8);
    KSG_LAUNCHED2();
    return KSG_OK;
}

int ksg_axis_count(const ksg_prepared* P, int axis, int64_t q_begin, int64_t q_end, const double* d_eps, int strict,
                   int32_t bias, int32_t* d_counts, void* stream) {
    KSG_REQUIRE(P && d_eps && d_counts, "ksg_axis_count: null pointer");
    KSG_REQUIRE(axis >= 0 && axis < P->dim, "ksg_axis_count: axis out of range");
    KSG_REQUIRE(P->axes_mode == 0, "ksg_axis_count: the set has lazy axes (ksg_axis_count_buckets, or ksg_prepare_axes first)");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= P->n, "ksg_axis_count: bad query range");
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    {
        const int rc0 = wait_axes(P, as_stream2(stream));
        if (rc0) return rc0;
    }
    // KSG_B200_AXIS_ORDER=prepared selects the thread-per-query kernel (A/B measurements)
    static const bool by_value = [] {
        const char* e = getenv("KSG_B200_AXIS_ORDER");
        return !(e && e[0] == 'p');
    }();
    if (by_value)
        axis_count_sorted_kernel<<<(unsigned)ceil_div(P->n, 256), 256, 0, as_stream2(stream)>>>(
            P->axis_vals + (int64_t)axis * P->ld, P->axis_pos + (int64_t)axis * P->ld, P->n, q_begin, nq, d_eps,
            strict ? 1 : 0, bias, d_counts);
    else
        axis_count_kernel<<<(unsigned)ceil_div(nq, AXIS_THREADS), AXIS_THREADS, 0, as_stream2(stream)>>>(
            P->sorted_f64 + (int64_t)axis * P->ld, P->axis_vals + (int64_t)axis * P->ld, P->n, q_begin, nq, d_eps,
            strict ? 1 : 0, bias, d_counts);
    KSG_LAUNCHED2();
    return KSG_OK;
}

int ksg_axis_count_multi(const ksg_prepared* P, const int32_t* h_axes, int n_axes, int64_t q_begin, int64_t q_end,
                         const double* d_eps, int strict, int32_t bias, int32_t* d_counts, void* stream) {
    KSG_REQUIRE(P && h_axes && d_eps && d_counts, "ksg_axis_count_multi: null pointer");
    KSG_REQUIRE(n_axes >= 1 && n_axes <= KSG_MAX_DIM, "ksg_axis_count_multi: 1 <= n_axes <= KSG_MAX_DIM");
    KSG_REQUIRE(P->axes_mode == 0, "ksg_axis_count_multi: the set has lazy axes (ksg_axis_count_buckets, or ksg_prepare_axes first)");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= P->n, "ksg_axis_count_multi: bad query range");
    AxisList axes;
    for (int a = 0; a < n_axes; ++a) {
        KSG_REQUIRE(h_axes[a] >= 0 && h_axes[a] < P->dim, "ksg_axis_count_multi: axis out of range");
        axes.axis[a] = h_axes[a];
    }
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    const int rc0 = wait_axes(P, as_stream2(stream));
    if (rc0) return rc0;
    axis_count_sorted_multi_kernel<<<dim3((unsigned)ceil_div(P->n, 256), (unsigned)n_axes), 256, 0, as_stream2(stream)>>>(
        P->axis_vals, P->axis_pos, P->ld, axes, P->n, q_begin, nq, d_eps, strict ? 1 : 0, bias, d_counts);
    KSG_LAUNCHED2();
    return KSG_OK;
}

int ksg_axis_count_buckets(const ksg_prepared* P, const int32_t* h_axes, int n_axes, int64_t q_begin, int64_t q_end,
                           const double* d_eps, int strict, int32_t bias, int32_t* d_counts, int32_t* d_n_flagged,
                           void* stream) {
    KSG_REQUIRE(P && h_axes && d_eps && d_counts && d_n_flagged, "ksg_axis_count_buckets: null pointer");
    KSG_REQUIRE(n_axes >= 1 && n_axes <= KSG_MAX_DIM, "ksg_axis_count_buckets: 1 <= n_axes <= KSG_MAX_DIM");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= P->n, "ksg_axis_count_buckets: bad query range");
    KSG_REQUIRE(P->axes_mode > 0 && P->axes_mode <= BUCKET_MAX_BITS && P->curve_table,
                "ksg_axis_count_buckets: the set has no value buckets (prepare it with primary = -3)");
    AxisList axes;
    for (int a = 0; a < n_axes; ++a) {
        KSG_REQUIRE(h_axes[a] >= 0 && h_axes[a] < P->dim, "ksg_axis_count_buckets: axis out of range");
        axes.axis[a] = h_axes[a];
    }
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    axis_count_bucket_kernel<<<dim3((unsigned)ceil_div(nq, 256), (unsigned)n_axes), 256, 0, as_stream2(stream)>>>(
        P->sorted_f64, P->ld, P->axis_vals, P->ld, P->axis_pos, P->ld, P->curve_table, P->center, P->axes_mode, axes,
        q_begin, nq, d_eps, strict ? 1 : 0, bias, d_counts, P->flags + 10, d_n_flagged);
    KSG_LAUNCHED2();
    return KSG_OK;
}

int ksg_prepare_axes(ksg_prepared* P, void* stream) {
    KSG_REQUIRE(P, "ksg_prepare_axes: null pointer");
    if (P->axes_mode == 0) return KSG_OK;  // already exact
    // sorted from the set's own float64 copy: keys = coordinate d in the prepared order, payload = the
    // prepared position itself, so axis_pos needs no remap (ties come out in prepared order)
    const PreparedLayout L = prepared_layout(P->n, P->dim);
    char* base = (char*)P->perm - L.off_perm;
    const int32_t* iota = (const int32_t*)(base + L.off_iota);
    void* cub_tmp = base + L.off_cub;
    cudaStream_t st = as_stream2(stream);
    for (int d = 0; d < P->dim; ++d) {
        const int rc = sort_pairs(cub_tmp, L.cub_bytes, P->sorted_f64 + (int64_t)d * P->ld, P->axis_vals + (int64_t)d * P->ld,
                                  iota, P->axis_pos + (int64_t)d * P->ld, P->n, st);
        if (rc) return rc;
    }
    P->axes_mode = 0;
    P->axes_ready = nullptr;
    return KSG_OK;
}

int ksg_scatter_f64(const double* d_src, const int32_t* d_perm, int64_t n, double* d_dst, void* stream) {
    KSG_REQUIRE(d_src && d_perm && d_dst && n >= 0, "ksg_scatter_f64: bad argument");
    if (n == 0) return KSG_OK;
    scatter_f64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream2(stream)>>>(d_src, d_perm, n, d_dst);
    KSG_LAUNCHED2();
    return KSG_OK;
}

int ksg_scatter_i32(const int32_t* d_src, const int32_t* d_perm, int64_t n, int32_t* d_dst, void* stream) {
    KSG_REQUIRE(d_src && d_perm && d_dst && n >= 0, "ksg_scatter_i32: bad argument");
    if (n == 0) return KSG_OK;
    scatter_i32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream2(stream)>>>(d_src, d_perm, n, d_dst);
    KSG_LAUNCHED2();
    return KSG_OK;
}

int ksg_sum_i32(const int32_t* d_values, int64_t n, unsigned long long* d_out, void* stream) {
    KSG_REQUIRE(d_values && d_out && n >= 0, "ksg_sum_i32: bad argument");
    if (n == 0) return KSG_OK;
    int64_t blocks = ceil_div(n, 256 * 8);
    if (blocks > 4096) blocks = 4096;
    sum_i32_kernel<<<(unsigned)blocks, 256, 0, as_stream2(stream)>>>(d_values, n, d_out);
    KSG_LAUNCHED2();
    return KSG_OK;
}

size_t ksg_sort_workspace(int64_t n) {
    if (n <= 0) return 0;
    return align_up(cub_sort_bytes(n), 256) + align_up((size_t)n * 4, 256);
}

int ksg_sort_f64(const double* d_keys_in, int64_t n, double* d_keys_out, int32_t* d_index_out, int algorithm,
                 void* d_workspace, size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(d_keys_in && d_keys_out && d_index_out && d_workspace, "ksg_sort_f64: null pointer");
    KSG_REQUIRE(n > 0 && n < (1ll << 31) - 1024, "ksg_sort_f64: need 0 < n < 2^31");
    KSG_REQUIRE(algorithm == 0 || algorithm == 1, "ksg_sort_f64: algorithm must be 0 (default) or 1 (radix sort)");
    if (workspace_bytes < ksg_sort_workspace(n)) {
        set_error("ksg_sort_f64: workspace of %zu bytes needed, %zu given", ksg_sort_workspace(n), workspace_bytes);
        return KSG_EWORKSPACE;
    }
    cudaStream_t st = as_stream2(stream);
    const size_t sort_bytes = align_up(cub_sort_bytes(n), 256);
    if (algorithm == 0 && n <= SS_MAX_N)
        return sample_sort_launch<double>(d_workspace, d_keys_in, d_keys_out, d_index_out, n, st);
    int32_t* iota = (int32_t*)((char*)d_workspace + sort_bytes);
    iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(iota, n);
    KSG_LAUNCHED2();
    size_t need = sort_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_workspace, need, d_keys_in, d_keys_out, iota, d_index_out, (int)n,
                                                    0, 64, st);
    if (e != cudaSuccess) {
        set_error("cub::DeviceRadixSort::SortPairs: %s", cudaGetErrorString(e));
        return KSG_ECUDA;
    }
    note_launch();
    return KSG_OK;
}

int ksg_gather_f64(const double* d_src, const int32_t* d_perm, int64_t n, double* d_dst, void* stream) {
    KSG_REQUIRE(d_src && d_perm && d_dst && n >= 0, "ksg_gather_f64: bad argument");
    if (n == 0) return KSG_OK;
    gather_f64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream2(stream)>>>(d_src, d_perm, n, d_dst);
    KSG_LAUNCHED2();
    return KSG_OK;
}

}  // extern "C"

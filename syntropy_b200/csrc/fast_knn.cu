// k-NN filter instantiations (DP = 2, 4, 6, 8) and their launcher.
#include "fast_count_launch.cuh"
#include <stdlib.h>

#include "fast_knn.cuh"
#include "fast_knn_group.cuh"
#include "fast_knn_warp.cuh"

namespace ksg {

int knn_f32_queries_per_block(int dim_padded) { return FAST_THREADS * (dim_padded <= 4 ? 4 : 2); }

// Queries per thread of the WINDOWED kernels (planning + work units): ONE.  The culling is limited by
// the warp's largest threshold and by its query box, both taken over 32*Q queries, and every scanned
// sub-tile costs 32*Q pair evaluations per lane group -- measured on C2 (2-D): Q = 4 -> 2 -> 1 evaluates
// 1058 -> 610 -> 391 million pairs and takes 1.14 -> 0.86 -> 0.82 ms; C5 (8-D): Q = 2 -> 1 halves the
// pairs (2.47 -> 2.0 s).  The dense kernel keeps Q = 4 / 2: there every pair is evaluated anyway and
// more queries per thread amortise the shared-memory reads.
template <int DP>
struct UnitsQ {
    static constexpr int Q = UNITS_Q;
};
int knn_units_queries_per_block(int dim_padded) {
    (void)dim_padded;
    return FAST_THREADS * UNITS_Q;
}

template <int DP>
static int launch_knn_dp(const KnnF32Args& a) {
    constexpr int Q = FastGeom<DP>::Q, ST = FastGeom<DP>::STAGES;
    auto kern = knn_f32_kernel<DP, Q, ST>;
    const size_t smem = (size_t)ST * FAST_TILE * DP * sizeof(float) + (size_t)a.kcap * FAST_THREADS * Q * 8;
    if (smem > 32 * 1024) {  // static shared (mbarriers) counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    dim3 grid((unsigned)ceil_div(a.nq, FAST_THREADS * Q), (unsigned)a.splits, 1);
    kern<<<grid, FAST_THREADS, smem, a.st>>>(a.shadow, a.tile_box, a.n_tiles, a.q_begin, a.nq, a.q_index, a.kcap,
                                              a.maxabs, a.dim, a.part_d32, a.part_idx, a.tile_counter);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("knn_f32_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

// Rows on each side of a query that seed its threshold in the planning pass (see knn_plan_kernel).
// KSG_B200_SEED_W overrides (experiments).
static int plan_seed_width(int dp) {
    static const int forced = [] {
        const char* e = getenv("KSG_B200_SEED_W");
        return e ? atoi(e) : 0;
    }();
    if (forced > 0) return forced;
    // 6-D / 8-D: +-24 rows of the curve are a poor sample of a query's neighbourhood (seed / true radius 1.5); the
    // sweep (profiles/r2ab_seed_window.txt) has C3 k-NN 42.2 -> 30.1 ms, C5 699 -> 603 ms, C4 25.6 -> 23.4 ms going
    // from +-96 to +-768 rows, flat beyond (the planning pass then costs what the tighter seeds save)
    return dp <= 4 ? KNN_SEED_W : 32 * KNN_SEED_W;
}

template <int DP, int KC>
static int launch_plan_kc(const KnnF32Args& a) {
    constexpr int Q = UnitsQ<DP>::Q;
    auto kern = knn_plan_kernel<DP, Q, KC>;
    const size_t smem = KC > 0 ? 0 : (size_t)a.kseed * FAST_THREADS * Q * 4;
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    if (cudaMemsetAsync(a.unit_meta, 0, 32, a.st) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return KSG_ECUDA; }
    kern<<<(unsigned)ceil_div(a.nq, FAST_THREADS * Q), FAST_THREADS, smem, a.st>>>(
        a.shadow, a.tile_box, a.n_tiles, a.q_begin, a.nq, a.kcap, a.kseed, plan_seed_width(DP), a.maxabs, a.dim, a.qshadow, a.qhome, a.seed_thr, (int2*)a.plan,
        a.plan_capacity, a.plan_count, a.plan_first, a.unit_meta, a.unit_done);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("knn_plan_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    // KSG_B200_UNIT_TILES (experiments): shortest unit, in tiles
    static const int min_tiles_env = [] { const char* e = getenv("KSG_B200_UNIT_TILES"); return e ? atoi(e) : 0; }();
    knn_unit_table_kernel<<<1, 1024, 0, a.st>>>(a.plan_count, ceil_div(a.nq, FAST_THREADS * Q), a.max_units,
                                                 min_tiles_env > 0 ? min_tiles_env : KNN_UNIT_MIN_TILES, a.unit_first,
                                                 a.unit_block, a.unit_meta);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("knn_unit_table_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

// Few dimensions: every warp its own worker (fast_knn_warp.cuh).  KSG_B200_UNITS=cta keeps the CTA-level
// kernel with the shared tile ring (experiments, A/B runs).
static bool warp_units_enabled() {
    static const bool on = [] {
        const char* e = getenv("KSG_B200_UNITS");
        return !(e && e[0] == 'c');
    }();
    return on;
}

template <int DP, int KC>
static int launch_wunits_kc(const KnnF32Args& a) {
    auto kern = knn_wunits_kernel<DP, KC>;
    const size_t smem = (FAST_THREADS / 32) * wunits_smem_per_warp(KC, DP);
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    // persistent: as many CTAs as fit on the machine (four work items per unit, one per warp)
    static int cached_dev = -1, cached_per_sm = 0;
    static size_t cached_smem = 0;
    int dev = 0;
    (void)cudaGetDevice(&dev);
    int per_sm = 0;
    cudaError_t e = cudaSuccess;
    if (dev == cached_dev && smem == cached_smem && cached_per_sm > 0) {
        per_sm = cached_per_sm;
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FAST_THREADS, smem);
        if (e != cudaSuccess || per_sm < 1) { (void)cudaGetLastError(); per_sm = 1; }
        cached_dev = dev; cached_smem = smem; cached_per_sm = per_sm;
    }
    int64_t grid = (int64_t)per_sm * sm_count();
    if (grid > a.max_units) grid = a.max_units;
    kern<<<(unsigned)grid, FAST_THREADS, smem, a.st>>>(a.shadow, a.sub_box, a.n_tiles, a.q_begin, a.nq, a.kcap, a.qshadow, a.qhome,
                                                      a.seed_thr, (const int2*)a.plan, a.plan_count, a.plan_first, a.unit_first,
                                                      a.unit_block, a.unit_meta, a.part_d32, a.part_idx, a.tile_counter,
                                                      a.unit_done, a.ra);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("knn_wunits_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

template <int DP>
static int launch_wunits_dp(const KnnF32Args& a) {
    static const int roomy = [] { const char* e = getenv("KSG_B200_WU_ROOMY"); return e ? atoi(e) : 0; }();
    switch (wunits_list_capacity(a.kcap + (roomy > 0 && a.kcap <= 16 ? 4 : 0))) {
        case 8: return launch_wunits_kc<DP, 8>(a);
        case 12: return launch_wunits_kc<DP, 12>(a);
        case 16: return launch_wunits_kc<DP, 16>(a);
        default: return launch_wunits_kc<DP, WU_MAX_KC>(a);
    }
}

// many dimensions: 8-query groups inside the filter warps (fast_knn_group.cuh)
template <int DP, int ST>
static int launch_gunits_st(const KnnF32Args& a) {
    auto kern = knn_gunits_kernel<DP, ST>;
    const size_t smem = (size_t)ST * (FAST_TILE * DP + FAST_NSUB * DP * 2) * sizeof(float) + (size_t)a.kcap * FAST_THREADS * 8;
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    static int cached_dev = -1, cached_per_sm = 0;
    static size_t cached_smem = 0;
    int dev = 0;
    (void)cudaGetDevice(&dev);
    int per_sm = 0;
    cudaError_t e = cudaSuccess;
    if (dev == cached_dev && smem == cached_smem && cached_per_sm > 0) {
        per_sm = cached_per_sm;
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, UNITS_THREADS, smem);
        if (e != cudaSuccess || per_sm < 1) { (void)cudaGetLastError(); per_sm = 1; }
        cached_dev = dev; cached_smem = smem; cached_per_sm = per_sm;
    }
    int64_t grid = (int64_t)per_sm * sm_count();
    if (grid > a.max_units) grid = a.max_units;
    kern<<<(unsigned)grid, UNITS_THREADS, smem, a.st>>>(a.shadow, a.tile_box, a.sub_box, a.n_tiles, a.q_begin, a.nq, a.kcap, a.qshadow,
                                                       a.qhome, a.seed_thr, (const int2*)a.plan, a.plan_count, a.plan_first,
                                                       a.unit_first, a.unit_block, a.unit_meta, a.part_d32, a.part_idx,
                                                       a.tile_counter, a.unit_done, a.ra);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("knn_gunits_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

// ring depth of the grouped kernel: the filter warps of a CTA do very different amounts of work per tile, a
// deeper ring lets the fast ones run further ahead (KSG_B200_GU_STAGES=2|3|4 to compare)
template <int DP>
static int launch_gunits_dp(const KnnF32Args& a) {
    static const int stages = [] { const char* e = getenv("KSG_B200_GU_STAGES"); return e ? atoi(e) : 0; }();
    switch (stages) {
        case 3: return launch_gunits_st<DP, 3>(a);
        case 4: return launch_gunits_st<DP, 4>(a);
        case 2: return launch_gunits_st<DP, 2>(a);
        default: return launch_gunits_st<DP, FastGeom<DP>::STAGES>(a);
    }
}

template <int DP>
static int launch_units_dp(const KnnF32Args& a) {
    if constexpr (DP <= 4 && UNITS_Q == 1) {
        if (warp_units_enabled() && a.kcap <= WU_MAX_KC) return launch_wunits_dp<DP>(a);
    }
    constexpr int Q = UnitsQ<DP>::Q, ST = FastGeom<DP>::STAGES;
    // many dimensions: 8-query groups inside the filter warps (fast_knn_group.cuh); KSG_B200_GROUPS=0: whole warps
    static const bool grouped = [] { const char* e = getenv("KSG_B200_GROUPS"); return !(e && e[0] == '0'); }();
    if constexpr (DP >= 6 && UNITS_Q == 1) {
        if (grouped) return launch_gunits_dp<DP>(a);
    }
    auto kern = knn_units_kernel<DP, Q, ST, (DP <= 4)>;
    const size_t smem = (size_t)ST * (FAST_TILE * DP + FAST_NSUB * DP * 2) * sizeof(float) +
                        (size_t)a.kcap * FAST_THREADS * Q * 8;
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    // persistent: as many CTAs as fit on the machine (never more than there can be units)
    // (the occupancy query costs ~10 us of host time: remember the last answer per instantiation)
    static int cached_dev = -1, cached_per_sm = 0;
    static size_t cached_smem = 0;
    int dev = 0;
    (void)cudaGetDevice(&dev);
    int per_sm = 0;
    cudaError_t e = cudaSuccess;
    if (dev == cached_dev && smem == cached_smem && cached_per_sm > 0) {
        per_sm = cached_per_sm;
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, UNITS_THREADS, smem);
        if (e != cudaSuccess || per_sm < 1) { (void)cudaGetLastError(); per_sm = 1; }
        cached_dev = dev; cached_smem = smem; cached_per_sm = per_sm;
    }
    int64_t grid = (int64_t)per_sm * sm_count();
    if (grid > a.max_units) grid = a.max_units;
    kern<<<(unsigned)grid, UNITS_THREADS, smem, a.st>>>(a.shadow, a.sub_box, a.n_tiles, a.q_begin, a.nq, a.kcap, a.qshadow, a.qhome, a.seed_thr,
                                                       (const int2*)a.plan, a.plan_count, a.plan_first, a.unit_first, a.unit_block,
                                                       a.unit_meta, a.part_d32, a.part_idx, a.tile_counter, a.unit_done, a.ra);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("knn_units_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

template <int DP>
static int launch_plan_dp(const KnnF32Args& a) {
    // (the register-resident seed network holds kseed entries)
    if (a.kseed <= 6) return launch_plan_kc<DP, 6>(a);
    if (a.kseed <= 8) return launch_plan_kc<DP, 8>(a);
    if (a.kseed <= 10) return launch_plan_kc<DP, 10>(a);
    if (a.kseed <= 12) return launch_plan_kc<DP, 12>(a);
    return launch_plan_kc<DP, 0>(a);
}

int knn_plan_cap() { return KNN_PLAN_CAP; }
int knn_unit_factor() { return KNN_UNIT_FACTOR; }

int launch_knn_units(int dim_padded, const KnnF32Args& a) {
    switch (dim_padded) {
        case 2: return launch_units_dp<2>(a);
        case 4: return launch_units_dp<4>(a);
        case 6: return launch_units_dp<6>(a);
        case 8: return launch_units_dp<8>(a);
        default: return KSG_EUNSUPPORTED;
    }
}

int launch_knn_plan(int dim_padded, const KnnF32Args& a) {
    switch (dim_padded) {
        case 2: return launch_plan_dp<2>(a);
        case 4: return launch_plan_dp<4>(a);
        case 6: return launch_plan_dp<6>(a);
        case 8: return launch_plan_dp<8>(a);
        default: return KSG_EUNSUPPORTED;
    }
}

int launch_knn_f32(int dim_padded, const KnnF32Args& a) {
    switch (dim_padded) {
        case 2: return launch_knn_dp<2>(a);
        case 4: return launch_knn_dp<4>(a);
        case 6: return launch_knn_dp<6>(a);
        case 8: return launch_knn_dp<8>(a);
        default: return KSG_EUNSUPPORTED;
    }
}

}  // namespace ksg

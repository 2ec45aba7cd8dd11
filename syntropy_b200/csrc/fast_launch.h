// Host-side launchers of the fast (fp32-filtered) kernels; each family of template
// instantiations lives in its own translation unit so the library builds in parallel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ksg_b200.h"

namespace ksg {

// What the exact refine needs besides the fp32 lists (see knn_refine_kernel / refine_single_list).
struct RefineArgs {
    int fused;                 // 1: blocks that own a single work unit are refined inside knn_units_kernel
    int kp;                    // k'
    int dim;
    const double* sorted;      // candidate coordinates, float64 SoA [dim][ld], prepared order
    int64_t ld;
    const double* qsorted;     // cross-set search: query coordinates [dim][qld] (else null)
    int64_t qld;
    const double* center;
    const double* maxabs;      // [dim] max |x - centre| of candidates (and cross-set queries)
    const int32_t* exact32;    // device flag "the fp32 copy is exact" (prepared flags[8]) or null (cross-set search)
    const int32_t* perm;
    int64_t* idx_out;          // optional [nq][kp]
    double* eps_out;           // [nq]
    int32_t* flags;            // [nq]
    int32_t* n_flagged;
};

struct KnnF32Args {
    const float* shadow;
    const float* tile_box;
    const float* sub_box;
    int64_t n_tiles;
    int64_t q_begin, nq;
    const int32_t* q_index;  // optional: explicit query rows (then q_begin is ignored)
    // optional, cross-set search (windowed only): queries = rows of qshadow, seeded around qhome
    const float* qshadow;
    const int32_t* qhome;
    int kcap, splits, windowed;
    const double* maxabs;
    int dim;
    // windowed engine only (all null otherwise): planning pass outputs and the unit table
    float* seed_thr;       // [nq] seeded thresholds; < 0: query deferred to the dense launch
    void* plan;            // int2[plan_capacity] pool of (tile, box gap) lists
    int64_t plan_capacity;
    int32_t* plan_count;   // [n_query_blocks] list length
    int32_t* plan_first;   // [n_query_blocks] list start in the pool
    int32_t* unit_first;   // [n_query_blocks + 1]
    int32_t* unit_block;   // [max_units]
    int32_t* unit_meta;    // [4]: number of units, tiles per unit, work-queue cursor
    int32_t* unit_done;    // [n_query_blocks] finished units per block (last-unit-done merge)
    int64_t max_units;
    float* part_d32;
    int32_t* part_idx;
    unsigned long long* tile_counter;
    cudaStream_t st;
    // windowed engine: rank of the seed distance (kcap: legacy seeds; < kcap: tight seeds, see
    // knn_plan_kernel) and the exact refine's arguments (fused into the units kernel when ra.fused)
    int kseed;
    RefineArgs ra;
};
constexpr int UNITS_Q = 1;  // queries per thread of the windowed k-NN kernels (see fast_knn.cu)
constexpr int COUNT_UNITS_Q = 1;  // ... and of the single-subspace windowed count kernels
int knn_f32_queries_per_block(int dim_padded);
int knn_units_queries_per_block(int dim_padded);
int launch_knn_f32(int dim_padded, const KnnF32Args& a);  // KSG_EUNSUPPORTED if dim_padded not in {2,4,6,8}
int launch_knn_plan(int dim_padded, const KnnF32Args& a);   // fills seed_thr / plan / plan_count / unit table
int launch_knn_units(int dim_padded, const KnnF32Args& a);  // persistent windowed filter over the work units
int knn_plan_cap();
int knn_unit_factor();

struct CountF32Args {
    const float* shadow;
    const float* tile_box;
    int64_t n_tiles;
    int64_t q_begin, nq;
    const float* thresholds;
    int windowed;
    uint32_t masks[4];  // only for the masked (generic) kernel
    int n_sub;
    int32_t* counts;    // [n_sub][nq], added to
    int splits;
    unsigned long long* tile_counter;
    cudaStream_t st;
    const float* sub_box;  // boxes of the 32-row sub-tiles (grouped leave-one-out kernel), may be null
};
// planning / work-unit buffers of the single-subspace windowed pass (see fast_count_units.cuh)
struct CountUnitsArgs {
    const float* sub_box;
    void* plan;            // int2[plan_capacity]
    int64_t plan_capacity;
    int32_t* plan_count;   // [n_query_blocks]
    int32_t* plan_first;   // [n_query_blocks]
    int32_t* unit_first;   // [n_query_blocks + 1]
    int32_t* unit_block;   // [max_units]
    int32_t* unit_meta;    // [4]
    int64_t max_units;
};
int count_f32_queries_per_block(int dim_padded);
int launch_count_full(int dim_padded, const CountF32Args& a, const CountUnitsArgs& w);  // one subspace = all coordinates
int launch_count_pair(int dx, int dy, const CountF32Args& a);          // subspaces x, y
int launch_count_cmi(int dx, int dy, int dz, const CountF32Args& a);   // subspaces z, xz, yz
int launch_count_loo(int d, const CountF32Args& a);                    // all leave-one-out subspaces
int launch_count_loo_grouped(int d, const CountF32Args& a);            // ... 8-query groups, sub-tile culling (or KSG_EUNSUPPORTED)
int launch_count_masked(int dim_padded, const CountF32Args& a);        // <= 4 run-time masks

void note_launch();  // launch counter (ksg_launch_count)
void note_launches(long long delta);  // ... for a replayed launch sequence (CUDA graph)

}  // namespace ksg

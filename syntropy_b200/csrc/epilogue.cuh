// K3: digamma table, psi epilogue programs, deterministic float64 sum.  HBM-bound.
// All float64 arithmetic uses the explicit round-to-nearest intrinsics so that no
// FMA contraction can change the reference's operation order.
#pragma once

#include "common.cuh"

namespace ksg {

constexpr double EULER_GAMMA = 0.57721566490153286061;

// digamma at a non-negative integer: cephes `psi` restricted to the arguments the
// path uses.  n <= 0 -> -inf (psi(0) is reached when eps == 0).
__device__ inline double psi_of_int(int64_t n) {
    if (n <= 0) return -CUDART_INF;
    if (n <= 10) {
        double y = 0.0;
        for (int64_t i = 1; i < n; ++i) y = __dadd_rn(y, __ddiv_rn(1.0, (double)i));
        return __dsub_rn(y, EULER_GAMMA);
    }
    const double s = (double)n;
    const double z = __ddiv_rn(1.0, __dmul_rn(s, s));
    double p = 8.33333333333333333333e-2;
    p = __dadd_rn(__dmul_rn(p, z), -2.10927960927960927961e-2);
    p = __dadd_rn(__dmul_rn(p, z), 7.57575757575757575758e-3);
    p = __dadd_rn(__dmul_rn(p, z), -4.16666666666666666667e-3);
    p = __dadd_rn(__dmul_rn(p, z), 3.96825396825396825397e-3);
    p = __dadd_rn(__dmul_rn(p, z), -8.33333333333333333333e-3);
    p = __dadd_rn(__dmul_rn(p, z), 8.33333333333333333333e-2);
    const double y = __dmul_rn(z, p);
    return __dsub_rn(__dsub_rn(log(s), __ddiv_rn(0.5, s)), y);
}

static __global__ void psi_table_kernel(double* __restrict__ table, int64_t entries) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < entries) table[i] = psi_of_int(i);
}

constexpr int MAX_PSI_STEPS = 2 * KSG_MAX_SUBSPACES;

struct PsiProgram {
    int n_steps;
    int apply_scale;
    double init;
    double psi_n;
    double scale;
    int16_t subspace[MAX_PSI_STEPS];
    int8_t offset[MAX_PSI_STEPS];
    int8_t sign[MAX_PSI_STEPS];
    int8_t centered[MAX_PSI_STEPS];
    double divisor[MAX_PSI_STEPS];
};

// one sample's value of the psi program (the reference's float64 operation order)
__device__ __forceinline__ double psi_program_value(const int32_t* __restrict__ counts, int64_t nq,
                                                    const PsiProgram& prog, const double* __restrict__ table,
                                                    int64_t entries, int64_t qi) {
    double v = prog.init;
    for (int t = 0; t < prog.n_steps; ++t) {
        int64_t c = (int64_t)counts[(int64_t)prog.subspace[t] * nq + qi] + prog.offset[t];
        c = c < 0 ? 0 : (c >= entries ? entries - 1 : c);
        double term = table[c];
        if (prog.centered[t]) term = __ddiv_rn(__dsub_rn(term, prog.psi_n), prog.divisor[t]);
        v = prog.sign[t] > 0 ? __dadd_rn(v, term) : __dsub_rn(v, term);
    }
    if (prog.apply_scale) v = __dmul_rn(v, prog.scale);
    return v;
}

static __global__ void psi_epilogue_kernel(const int32_t* __restrict__ counts, int64_t nq,
                                    const __grid_constant__ PsiProgram prog,
                                    const double* __restrict__ table, int64_t entries,
                                    double* __restrict__ ptw) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    double v = prog.init;
    for (int t = 0; t < prog.n_steps; ++t) {
        int64_t c = (int64_t)counts[(int64_t)prog.subspace[t] * nq + qi] + prog.offset[t];
        c = c < 0 ? 0 : (c >= entries ? entries - 1 : c);
        double term = table[c];
        if (prog.centered[t]) term = __ddiv_rn(__dsub_rn(term, prog.psi_n), prog.divisor[t]);
        v = prog.sign[t] > 0 ? __dadd_rn(v, term) : __dsub_rn(v, term);
    }
    if (prog.apply_scale) v = __dmul_rn(v, prog.scale);
    ptw[qi] = v;
}

static __global__ void entropy_epilogue_kernel(const double* __restrict__ eps, int64_t nq, double dimf,
                                        double base, double* __restrict__ ptw) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    // 0 + ((-psi_k + psi_N) + d * log(2 * eps))     knn/shannon.py:84-85
    const double t = __dmul_rn(dimf, log(__dmul_rn(2.0, eps[qi])));
    ptw[qi] = __dadd_rn(0.0, __dadd_rn(base, t));
}

static __global__ void kl_epilogue_kernel(const double* __restrict__ nu, const double* __restrict__ rho,
                                   int64_t nq, double dimf, double log_ratio, double* __restrict__ ptw) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    // d * log(nu / rho) + log(m / (n - 1))          knn/shannon.py:407
    ptw[qi] = __dadd_rn(__dmul_rn(dimf, log(__ddiv_rn(nu[qi], rho[qi]))), log_ratio);
}

static __global__ void check_finite_kernel(const double* __restrict__ v, int64_t count, int32_t* __restrict__ flag) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x)
        bad |= !isfinite(v[i]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(flag, 1);
}

// Deterministic sum: SUM_BLOCKS fixed chunks, each reduced by a fixed-shape tree;
// the second stage reduces the SUM_BLOCKS partials the same way.  The summation
// tree depends only on n, never on the device or launch geometry.
constexpr int SUM_BLOCKS = 1024;
constexpr int SUM_THREADS = 256;

__device__ inline double block_tree_sum(double v) {
    __shared__ double sh[SUM_THREADS];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = SUM_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + off]);
        __syncthreads();
    }
    return sh[0];
}

static __global__ void __launch_bounds__(SUM_THREADS)
sum_stage1_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ partial) {
    const int64_t chunk = (n + SUM_BLOCKS - 1) / SUM_BLOCKS;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = min(n, lo + chunk);
    double acc = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += SUM_THREADS) acc = __dadd_rn(acc, v[i]);
    const double total = block_tree_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(SUM_THREADS)
sum_stage2_kernel(const double* __restrict__ partial, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < SUM_BLOCKS; i += SUM_THREADS) acc = __dadd_rn(acc, partial[i]);
    const double total = block_tree_sum(acc);
    if (threadIdx.x == 0) out[0] = total;
}

// The tail of an estimator step in ONE launch: pointwise values of ALL n samples in the prepared order
// (PSI: the psi program over the count rows; else: given values) -> the caller's sample order
// (out[perm[i]]) + the deterministic sum.  Same fixed tree as sum_stage1/2 (SUM_BLOCKS chunks of the
// PREPARED order), the second stage run by whichever block finishes last (ticket: 0 on entry, 0 again
// on exit).  The tree depends only on n and on the prepared order, never on the launch geometry or the
// number of ranks.
template <bool PSI>
static __global__ void __launch_bounds__(SUM_THREADS)
finish_kernel(const int32_t* __restrict__ counts, const double* __restrict__ values, int64_t n,
              const __grid_constant__ PsiProgram prog, const double* __restrict__ table, int64_t entries,
              const int32_t* __restrict__ perm, double* __restrict__ out, double* __restrict__ partial,
              int32_t* __restrict__ ticket, double* __restrict__ total) {
    const int64_t chunk = (n + SUM_BLOCKS - 1) / SUM_BLOCKS;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = min(n, lo + chunk);
    double acc = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += SUM_THREADS) {
        double v;
        if constexpr (PSI) v = psi_program_value(counts, n, prog, table, entries, i);
        else v = values[i];
        out[perm[i]] = v;
        acc = __dadd_rn(acc, v);
    }
    const double block_total = block_tree_sum(acc);
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_total;
        __threadfence();
        s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double acc2 = 0.0;
    for (int i = threadIdx.x; i < SUM_BLOCKS; i += SUM_THREADS) acc2 = __dadd_rn(acc2, __ldcg(&partial[i]));
    const double t = block_tree_sum(acc2);
    if (threadIdx.x == 0) {
        total[0] = t;
        *ticket = 0;
    }
}

}  // namespace ksg

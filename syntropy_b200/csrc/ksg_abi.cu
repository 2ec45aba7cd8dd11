// libksg_b200: the extern "C" surface declared in include/ksg_b200.h.
// Argument checking, launch geometry and dispatch only; the kernels live in the
// .cuh files next to this one.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "exact_f64.cuh"
#include "fast_launch.h"
#include "minkowski_f64.cuh"
#include "p2p.cuh"

namespace ksg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
        cached = sms;
    } else {
        (void)cudaGetLastError();
        return 148;
    }
    return cached;
}

static inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

// every kernel launch of the library goes through this: error check + launch counter
static unsigned long long g_launches = 0;
#define KSG_LAUNCHED()                           \
    do {                                         \
        __atomic_add_fetch(&ksg::g_launches, 1ull, __ATOMIC_RELAXED); \
        KSG_CUDA_TRY(cudaGetLastError());        \
    } while (0)

void note_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
void note_launches(long long delta) { __atomic_add_fetch(&g_launches, (unsigned long long)delta, __ATOMIC_RELAXED); }

template <typename K>
static int allow_smem(K kernel, size_t bytes) {
    if (bytes > 32 * 1024)
        KSG_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return KSG_OK;
}

// ------------------------------------------------------------------ K1 dispatch
struct KnnPlan {
    int64_t qblocks;
    int splits;
    int64_t cand_per_split;
    size_t dist_bytes;
    size_t idx_bytes;
};

static KnnPlan knn_plan(int64_t n, int64_t nq, int kp) {
    KnnPlan p;
    p.qblocks = ceil_div(nq, F64_THREADS);
    p.splits = choose_splits(p.qblocks, n, F64_TILE);
    int64_t per = ceil_div(n, p.splits);
    per = ceil_div(per, F64_TILE) * F64_TILE;
    p.cand_per_split = per;
    p.splits = (int)ceil_div(n, per);
    if (p.splits < 1) p.splits = 1;
    p.dist_bytes = (size_t)p.splits * kp * nq * sizeof(double);
    p.idx_bytes = (size_t)p.splits * kp * nq * sizeof(int64_t);
    return p;
}

template <int DT, bool WITH_IDX>
static int launch_knn_f64(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                          int64_t q_begin, int64_t nq, int kp, const KnnPlan& p, double* part_dist,
                          int64_t* part_idx, cudaStream_t st) {
    size_t smem = (size_t)dim * F64_TILE * sizeof(double) + (size_t)kp * F64_THREADS * sizeof(double) +
                  (WITH_IDX ? (size_t)kp * F64_THREADS * sizeof(int64_t) : 0);
    int rc = allow_smem(knn_f64_kernel<DT, WITH_IDX>, smem);
    if (rc) return rc;
    dim3 grid((unsigned)p.qblocks, (unsigned)p.splits, 1);
    knn_f64_kernel<DT, WITH_IDX><<<grid, F64_THREADS, smem, st>>>(pts, ld, n, qry, ldq, q_begin, nq, dim, kp,
                                                                   p.cand_per_split, part_dist, part_idx);
    KSG_LAUNCHED();
    return KSG_OK;
}

template <bool WITH_IDX>
static int dispatch_knn_f64(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                            int64_t q_begin, int64_t nq, int kp, const KnnPlan& p, double* part_dist,
                            int64_t* part_idx, cudaStream_t st) {
#define KSG_CASE(D)                                                                                        \
    case D:                                                                                                \
        return launch_knn_f64<D, WITH_IDX>(pts, ld, n, dim, qry, ldq, q_begin, nq, kp, p, part_dist,       \
                                           part_idx, st);
    switch (dim) {
        KSG_CASE(1) KSG_CASE(2) KSG_CASE(3) KSG_CASE(4) KSG_CASE(5) KSG_CASE(6) KSG_CASE(7) KSG_CASE(8)
        default:
            return launch_knn_f64<0, WITH_IDX>(pts, ld, n, dim, qry, ldq, q_begin, nq, kp, p, part_dist,
                                               part_idx, st);
    }
#undef KSG_CASE
}

template <int DT>
static int launch_count_f64(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                            int64_t q_begin, int64_t nq, const MaskTable& masks, int n_sub, const double* eps,
                            int64_t eps_stride, int strict, int32_t* counts, cudaStream_t st) {
    int64_t qblocks = ceil_div(nq, F64_THREADS);
    int zchunks = (int)ceil_div(n_sub, F64_SUBCHUNK);
    int splits = choose_splits(qblocks * zchunks, n, F64_TILE);
    int64_t per = ceil_div(ceil_div(n, splits), F64_TILE) * F64_TILE;
    splits = (int)ceil_div(n, per);
    if (splits < 1) splits = 1;
    size_t smem = (size_t)dim * F64_TILE * sizeof(double);
    int rc = allow_smem(count_f64_kernel<DT>, smem);
    if (rc) return rc;
    dim3 grid((unsigned)qblocks, (unsigned)splits, (unsigned)zchunks);
    count_f64_kernel<DT><<<grid, F64_THREADS, smem, st>>>(pts, ld, n, qry, ldq, q_begin, nq, dim, masks, n_sub,
                                                          eps, eps_stride, strict, per, counts);
    KSG_LAUNCHED();
    return KSG_OK;
}

template <int MT>
static int launch_paircount(const double* x, const double* y, int64_t nt, int m, double r, int64_t i_begin,
                            int64_t ni, unsigned long long* out, cudaStream_t st) {
    int64_t iblocks = ceil_div(ni, PC_THREADS);
    int splits = choose_splits(iblocks, nt, PC_TILE);
    int64_t per = ceil_div(ceil_div(nt, splits), PC_TILE) * PC_TILE;
    splits = (int)ceil_div(nt, per);
    if (splits < 1) splits = 1;
    dim3 grid((unsigned)iblocks, (unsigned)splits, 1);
    paircount_f64_kernel<MT><<<grid, PC_THREADS, 0, st>>>(x, y, nt, m, r, i_begin, ni, per, out);
    KSG_LAUNCHED();
    return KSG_OK;
}

}  // namespace ksg

using namespace ksg;

extern "C" {

int ksg_version(void) { return 1000; }

const char* ksg_last_error(void) { return g_err; }

unsigned long long ksg_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int ksg_device_info(int* sms, int* cc_major, int* cc_minor, int* sm_clock_khz) {
    int dev = 0;
    KSG_CUDA_TRY(cudaGetDevice(&dev));
    if (sms) KSG_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) KSG_CUDA_TRY(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) KSG_CUDA_TRY(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_clock_khz) KSG_CUDA_TRY(cudaDeviceGetAttribute(sm_clock_khz, cudaDevAttrClockRate, dev));
    return KSG_OK;
}

int ksg_check_finite(const double* d_values, int64_t count, int32_t* d_flag, void* stream) {
    KSG_REQUIRE(d_values && d_flag && count >= 0, "ksg_check_finite: null pointer or negative count");
    if (count == 0) return KSG_OK;
    int blocks = (int)(ceil_div(count, 256) < 4 * sm_count() ? ceil_div(count, 256) : 4 * sm_count());
    check_finite_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_values, count, d_flag);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_psi_table(double* d_table, int64_t n_entries, void* stream) {
    KSG_REQUIRE(d_table && n_entries > 0, "ksg_psi_table: null table or non-positive size");
    psi_table_kernel<<<(unsigned)ceil_div(n_entries, 256), 256, 0, as_stream(stream)>>>(d_table, n_entries);
    KSG_LAUNCHED();
    return KSG_OK;
}

size_t ksg_knn_radius_workspace(int64_t n, int64_t nq, int dim, int kprime) {
    (void)dim;
    if (n <= 0 || nq <= 0 || kprime <= 0) return 0;
    KnnPlan p = knn_plan(n, nq, kprime);
    return p.dist_bytes + p.idx_bytes + 256;
}

int ksg_knn_radius(const double* d_points, int64_t ld, int64_t n, int dim, const double* d_queries,
                   int64_t ldq, int64_t q_begin, int64_t q_end, int kprime, double* d_eps_out,
                   int64_t* d_idx_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(d_points && d_eps_out, "ksg_knn_radius: null pointer");
    KSG_REQUIRE(n > 0 && ld >= n, "ksg_knn_radius: need n > 0 and ld >= n");
    KSG_REQUIRE(dim >= 1 && dim <= KSG_MAX_DIM, "ksg_knn_radius: dim out of range");
    KSG_REQUIRE(kprime >= 1 && kprime <= KSG_MAX_KPRIME, "ksg_knn_radius: kprime out of range");
    if (!d_queries) {
        d_queries = d_points;
        ldq = ld;
        KSG_REQUIRE(q_end <= n, "ksg_knn_radius: query range exceeds the point set");
    }
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && ldq >= q_end, "ksg_knn_radius: bad query range");
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    KnnPlan p = knn_plan(n, nq, kprime);
    if (workspace_bytes < p.dist_bytes + p.idx_bytes || !d_workspace) {
        set_error("ksg_knn_radius: workspace of %zu bytes needed, %zu given", p.dist_bytes + p.idx_bytes,
                  workspace_bytes);
        return KSG_EWORKSPACE;
    }
    double* part_dist = (double*)d_workspace;
    int64_t* part_idx = (int64_t*)((char*)d_workspace + p.dist_bytes);
    cudaStream_t st = as_stream(stream);
    int rc;
    if (d_idx_out)
        rc = dispatch_knn_f64<true>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, kprime, p, part_dist,
                                    part_idx, st);
    else
        rc = dispatch_knn_f64<false>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, kprime, p, part_dist,
                                     part_idx, st);
    if (rc) return rc;
    unsigned mb = (unsigned)ceil_div(nq, 128);
    if (d_idx_out)
        knn_merge_kernel<true><<<mb, 128, 0, st>>>(part_dist, part_idx, nq, kprime, p.splits, d_eps_out, d_idx_out);
    else
        knn_merge_kernel<false><<<mb, 128, 0, st>>>(part_dist, part_idx, nq, kprime, p.splits, d_eps_out, nullptr);
    KSG_LAUNCHED();
    return KSG_OK;
}

size_t ksg_count_subspaces_workspace(int64_t n, int64_t nq, int dim, int n_subspaces) {
    (void)n; (void)nq; (void)dim; (void)n_subspaces;
    return 256;
}

int ksg_count_subspaces(const double* d_points, int64_t ld, int64_t n, int dim, const double* d_queries,
                        int64_t ldq, int64_t q_begin, int64_t q_end, const uint32_t* h_masks, int n_subspaces,
                        const double* d_eps, int64_t eps_stride, int strict, int32_t bias, int32_t* d_counts,
                        void* d_workspace, size_t workspace_bytes, void* stream) {
    (void)d_workspace; (void)workspace_bytes;
    KSG_REQUIRE(d_points && h_masks && d_eps && d_counts, "ksg_count_subspaces: null pointer");
    KSG_REQUIRE(n > 0 && ld >= n, "ksg_count_subspaces: need n > 0 and ld >= n");
    KSG_REQUIRE(dim >= 1 && dim <= KSG_MAX_DIM, "ksg_count_subspaces: dim out of range");
    KSG_REQUIRE(n_subspaces >= 1 && n_subspaces <= KSG_MAX_SUBSPACES, "ksg_count_subspaces: subspace count");
    if (!d_queries) {
        d_queries = d_points;
        ldq = ld;
        KSG_REQUIRE(q_end <= n, "ksg_count_subspaces: query range exceeds the point set");
    }
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && ldq >= q_end, "ksg_count_subspaces: bad query range");
    const int64_t nq = q_end - q_begin;
    KSG_REQUIRE(eps_stride == 0 || eps_stride >= nq, "ksg_count_subspaces: eps_stride must be 0 or >= nq");
    if (nq == 0) return KSG_OK;
    MaskTable masks;
    memset(&masks, 0, sizeof(masks));
    const uint32_t all = dim == 32 ? 0xffffffffu : ((1u << dim) - 1u);
    for (int s = 0; s < n_subspaces; ++s) {
        KSG_REQUIRE(h_masks[s] != 0 && (h_masks[s] & ~all) == 0, "ksg_count_subspaces: empty or out-of-range mask");
        masks.m[s] = h_masks[s];
    }
    cudaStream_t st = as_stream(stream);
    const int64_t total = (int64_t)n_subspaces * nq;
    fill_i32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(d_counts, total, bias);
    KSG_LAUNCHED();
#define KSG_CASE(D)                                                                                           \
    case D:                                                                                                   \
        return launch_count_f64<D>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, masks, n_subspaces,     \
                                   d_eps, eps_stride, strict, d_counts, st);
    switch (dim) {
        KSG_CASE(1) KSG_CASE(2) KSG_CASE(3) KSG_CASE(4) KSG_CASE(5) KSG_CASE(6) KSG_CASE(7) KSG_CASE(8)
        default:
            return launch_count_f64<0>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, masks, n_subspaces,
                                       d_eps, eps_stride, strict, d_counts, st);
    }
#undef KSG_CASE
}

int ksg_psi_epilogue(const int32_t* d_counts, int64_t nq, int n_subspaces, const ksg_psi_step* h_steps,
                     int n_steps, double init, double psi_n, int apply_scale, double scale,
                     const double* d_psi_table, int64_t table_entries, double* d_ptw, void* stream) {
    KSG_REQUIRE(d_counts && h_steps && d_psi_table && d_ptw, "ksg_psi_epilogue: null pointer");
    KSG_REQUIRE(n_steps >= 0 && n_steps <= MAX_PSI_STEPS, "ksg_psi_epilogue: too many steps");
    KSG_REQUIRE(table_entries >= 1 && nq >= 0, "ksg_psi_epilogue: bad sizes");
    if (nq == 0) return KSG_OK;
    PsiProgram prog;
    memset(&prog, 0, sizeof(prog));
    prog.n_steps = n_steps;
    prog.apply_scale = apply_scale;
    prog.init = init;
    prog.psi_n = psi_n;
    prog.scale = scale;
    for (int t = 0; t < n_steps; ++t) {
        KSG_REQUIRE(h_steps[t].subspace >= 0 && h_steps[t].subspace < n_subspaces,
                    "ksg_psi_epilogue: step refers to a missing subspace");
        KSG_REQUIRE(h_steps[t].count_offset >= -100 && h_steps[t].count_offset <= 100,
                    "ksg_psi_epilogue: count_offset out of range");
        prog.subspace[t] = (int16_t)h_steps[t].subspace;
        prog.offset[t] = (int8_t)h_steps[t].count_offset;
        prog.sign[t] = (int8_t)(h_steps[t].sign >= 0 ? 1 : -1);
        prog.centered[t] = (int8_t)(h_steps[t].centered != 0);
        prog.divisor[t] = h_steps[t].divisor;
    }
    psi_epilogue_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, as_stream(stream)>>>(d_counts, nq, prog, d_psi_table,
                                                                                   table_entries, d_ptw);
    KSG_LAUNCHED();
    return KSG_OK;
}

static int fill_psi_program(PsiProgram& prog, int n_subspaces, const ksg_psi_step* h_steps, int n_steps, double init,
                            double psi_n, int apply_scale, double scale) {
    KSG_REQUIRE(n_steps >= 0 && n_steps <= MAX_PSI_STEPS, "psi program: too many steps");
    memset(&prog, 0, sizeof(prog));
    prog.n_steps = n_steps;
    prog.apply_scale = apply_scale;
    prog.init = init;
    prog.psi_n = psi_n;
    prog.scale = scale;
    for (int t = 0; t < n_steps; ++t) {
        KSG_REQUIRE(h_steps[t].subspace >= 0 && h_steps[t].subspace < n_subspaces,
                    "psi program: step refers to a missing subspace");
        KSG_REQUIRE(h_steps[t].count_offset >= -100 && h_steps[t].count_offset <= 100,
                    "psi program: count_offset out of range");
        prog.subspace[t] = (int16_t)h_steps[t].subspace;
        prog.offset[t] = (int8_t)h_steps[t].count_offset;
        prog.sign[t] = (int8_t)(h_steps[t].sign >= 0 ? 1 : -1);
        prog.centered[t] = (int8_t)(h_steps[t].centered != 0);
        prog.divisor[t] = h_steps[t].divisor;
    }
    return KSG_OK;
}

size_t ksg_finish_workspace(int64_t n) {
    (void)n;
    return SUM_BLOCKS * sizeof(double);
}

int ksg_psi_finish(const int32_t* d_counts, int64_t n, int n_subspaces, const ksg_psi_step* h_steps, int n_steps,
                   double init, double psi_n, int apply_scale, double scale, const double* d_psi_table,
                   int64_t table_entries, const int32_t* d_perm, double* d_out, double* d_total, int32_t* d_ticket,
                   void* d_workspace, size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(d_counts && h_steps && d_psi_table && d_perm && d_out && d_total && d_ticket, "ksg_psi_finish: null pointer");
    KSG_REQUIRE(table_entries >= 1 && n >= 0, "ksg_psi_finish: bad sizes");
    if (!d_workspace || workspace_bytes < SUM_BLOCKS * sizeof(double)) {
        set_error("ksg_psi_finish: workspace of %zu bytes needed", SUM_BLOCKS * sizeof(double));
        return KSG_EWORKSPACE;
    }
    PsiProgram prog;
    const int rc = fill_psi_program(prog, n_subspaces, h_steps, n_steps, init, psi_n, apply_scale, scale);
    if (rc) return rc;
    finish_kernel<true><<<SUM_BLOCKS, SUM_THREADS, 0, as_stream(stream)>>>(d_counts, nullptr, n, prog, d_psi_table,
                                                                          table_entries, d_perm, d_out,
                                                                          (double*)d_workspace, d_ticket, d_total);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_scatter_sum_f64(const double* d_values_sorted, const int32_t* d_perm, int64_t n, double* d_out, double* d_total,
                        int32_t* d_ticket, void* d_workspace, size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(d_values_sorted && d_perm && d_out && d_total && d_ticket && n >= 0, "ksg_scatter_sum_f64: bad argument");
    if (!d_workspace || workspace_bytes < SUM_BLOCKS * sizeof(double)) {
        set_error("ksg_scatter_sum_f64: workspace of %zu bytes needed", SUM_BLOCKS * sizeof(double));
        return KSG_EWORKSPACE;
    }
    PsiProgram prog;
    memset(&prog, 0, sizeof(prog));
    finish_kernel<false><<<SUM_BLOCKS, SUM_THREADS, 0, as_stream(stream)>>>(nullptr, d_values_sorted, n, prog, nullptr, 0,
                                                                           d_perm, d_out, (double*)d_workspace,
                                                                           d_ticket, d_total);
    KSG_LAUNCHED();
    return KSG_OK;
}

// ---- peer-memory exchange of the sharded path (csrc/p2p.cuh)
static int fill_peer_list(PeerList& peers, void* const* h_peer_payload, unsigned long long* const* h_peer_flags,
                          int32_t* const* h_peer_aux, int rank, int world) {
    KSG_REQUIRE(h_peer_payload && h_peer_flags, "peer list: null pointer");
    KSG_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world, "peer list: bad rank / world size");
    memset(&peers, 0, sizeof(peers));
    peers.world = world;
    peers.rank = rank;
    for (int r = 0; r < world; ++r) {
        KSG_REQUIRE(h_peer_payload[r] && h_peer_flags[r], "peer list: null peer pointer");
        peers.payload[r] = h_peer_payload[r];
        peers.flags[r] = h_peer_flags[r];
        peers.aux[r] = h_peer_aux ? h_peer_aux[r] : nullptr;
    }
    return KSG_OK;
}

int ksg_psi_epilogue_push(const int32_t* d_counts, int64_t nq, int n_subspaces, const ksg_psi_step* h_steps, int n_steps,
                          double init, double psi_n, int apply_scale, double scale, const double* d_psi_table,
                          int64_t table_entries, void* const* h_peer_values, int64_t q_begin,
                          const int32_t* d_nflag, int32_t* const* h_peer_aux, unsigned long long* const* h_peer_flags,
                          int rank, int world, unsigned long long epoch, int32_t* d_ticket, void* stream) {
    KSG_REQUIRE(d_psi_table && d_ticket && h_steps, "ksg_psi_epilogue_push: null pointer");
    KSG_REQUIRE(nq >= 0 && q_begin >= 0 && table_entries >= 1 && (nq == 0 || d_counts), "ksg_psi_epilogue_push: bad sizes");
    PsiProgram prog;
    int rc = fill_psi_program(prog, n_subspaces, h_steps, n_steps, init, psi_n, apply_scale, scale);
    if (rc) return rc;
    PeerList peers;
    rc = fill_peer_list(peers, h_peer_values, h_peer_flags, h_peer_aux, rank, world);
    if (rc) return rc;
    // (a rank with an empty slice still publishes its flags: one block)
    int64_t nb = nq > 0 ? ceil_div(nq, 256 * 4) : 1;
    if (nb > 148 * 4) nb = 148 * 4;
    const unsigned blocks = (unsigned)nb;
    psi_push_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_counts, nq, prog, d_psi_table, table_entries, peers, q_begin,
                                                           d_nflag, epoch, d_ticket);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_p2p_push(const void* d_src, size_t bytes, void* const* h_peer_dst, size_t dst_offset,
                 unsigned long long* const* h_peer_flags, int rank, int world, unsigned long long epoch,
                 int32_t* d_ticket, void* stream) {
    KSG_REQUIRE(d_ticket && (bytes == 0 || d_src), "ksg_p2p_push: null pointer");
    KSG_REQUIRE(bytes % 8 == 0 && dst_offset % 8 == 0 && ((uintptr_t)d_src % 8) == 0,
                "ksg_p2p_push: sizes, offsets and pointers must be multiples of 8 bytes");
    PeerList peers;
    const int rc = fill_peer_list(peers, h_peer_dst, h_peer_flags, nullptr, rank, world);
    if (rc) return rc;
    const int64_t n8 = (int64_t)(bytes / 8);
    int64_t blocks = ceil_div(n8 > 0 ? n8 : 1, 256 * 4);
    if (blocks > 148 * 4) blocks = 148 * 4;
    p2p_push_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const uint2*)d_src, n8, peers, (int64_t)dst_offset,
                                                                     epoch, d_ticket);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_p2p_wait(const unsigned long long* d_flags, int world, unsigned long long epoch, const int32_t* d_aux,
                 int32_t* d_aux_sum, void* stream) {
    KSG_REQUIRE(d_flags && world >= 1 && world <= P2P_MAX_RANKS, "ksg_p2p_wait: bad argument");
    p2p_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(d_flags, world, epoch, d_aux, d_aux_sum);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_entropy_epilogue(const double* d_eps, int64_t nq, int dim, double psi_k, double psi_n, double* d_ptw,
                         void* stream) {
    KSG_REQUIRE(d_eps && d_ptw && nq >= 0 && dim >= 1, "ksg_entropy_epilogue: bad argument");
    if (nq == 0) return KSG_OK;
    const double base = -psi_k + psi_n;
    entropy_epilogue_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, as_stream(stream)>>>(d_eps, nq, (double)dim,
                                                                                       base, d_ptw);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_kl_epilogue(const double* d_nu, const double* d_rho, int64_t nq, int dim, double log_ratio,
                    double* d_ptw, void* stream) {
    KSG_REQUIRE(d_nu && d_rho && d_ptw && nq >= 0 && dim >= 1, "ksg_kl_epilogue: bad argument");
    if (nq == 0) return KSG_OK;
    kl_epilogue_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, as_stream(stream)>>>(d_nu, d_rho, nq, (double)dim,
                                                                                  log_ratio, d_ptw);
    KSG_LAUNCHED();
    return KSG_OK;
}

size_t ksg_sum_workspace(int64_t n) {
    (void)n;
    return SUM_BLOCKS * sizeof(double);
}

int ksg_sum_f64(const double* d_values, int64_t n, double* d_out, void* d_workspace, size_t workspace_bytes,
                void* stream) {
    KSG_REQUIRE(d_values && d_out && n >= 0, "ksg_sum_f64: bad argument");
    if (!d_workspace || workspace_bytes < SUM_BLOCKS * sizeof(double)) {
        set_error("ksg_sum_f64: workspace of %zu bytes needed", SUM_BLOCKS * sizeof(double));
        return KSG_EWORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    sum_stage1_kernel<<<SUM_BLOCKS, SUM_THREADS, 0, st>>>(d_values, n, (double*)d_workspace);
    KSG_LAUNCHED();
    sum_stage2_kernel<<<1, SUM_THREADS, 0, st>>>((const double*)d_workspace, d_out);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_subspace_radius_from_neighbours(const double* d_points, int64_t ld, int64_t n, int dim,
                                        int64_t q_begin, int64_t q_end, const int64_t* d_idx, int kprime,
                                        int skip_first, const uint32_t* h_masks, int n_subspaces,
                                        double* d_eps_out, void* stream) {
    KSG_REQUIRE(d_points && d_idx && h_masks && d_eps_out, "ksg_subspace_radius_from_neighbours: null pointer");
    KSG_REQUIRE(n > 0 && ld >= n && dim >= 1 && dim <= KSG_MAX_DIM, "ksg_subspace_radius_from_neighbours: sizes");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= n, "ksg_subspace_radius_from_neighbours: query range");
    KSG_REQUIRE(n_subspaces >= 1 && n_subspaces <= KSG_MAX_SUBSPACES, "ksg_subspace_radius_from_neighbours: subspaces");
    KSG_REQUIRE(kprime >= 1 && kprime <= KSG_MAX_KPRIME, "ksg_subspace_radius_from_neighbours: kprime");
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    MaskTable masks;
    memset(&masks, 0, sizeof(masks));
    for (int s = 0; s < n_subspaces; ++s) masks.m[s] = h_masks[s];
    subspace_radius_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, as_stream(stream)>>>(
        d_points, ld, dim, q_begin, nq, n, d_idx, kprime, skip_first, masks, n_subspaces, d_eps_out);
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_paircount_fixed_radius(const double* d_x, const double* d_y, int64_t n_templates, int m, double r,
                               int64_t i_begin, int64_t i_end, unsigned long long* d_counts, void* stream) {
    KSG_REQUIRE(d_x && d_y && d_counts, "ksg_paircount_fixed_radius: null pointer");
    KSG_REQUIRE(n_templates > 0 && m >= 1 && m < KSG_MAX_DIM, "ksg_paircount_fixed_radius: sizes");
    KSG_REQUIRE(i_begin >= 0 && i_end >= i_begin && i_end <= n_templates, "ksg_paircount_fixed_radius: range");
    const int64_t ni = i_end - i_begin;
    if (ni == 0) return KSG_OK;
    cudaStream_t st = as_stream(stream);
#define KSG_CASE(M) \
    case M:         \
        return launch_paircount<M>(d_x, d_y, n_templates, m, r, i_begin, ni, d_counts, st);
    switch (m) {
        KSG_CASE(1) KSG_CASE(2) KSG_CASE(3) KSG_CASE(4) KSG_CASE(5) KSG_CASE(6)
        default:
            return launch_paircount<0>(d_x, d_y, n_templates, m, r, i_begin, ni, d_counts, st);
    }
#undef KSG_CASE
}

// ------------------------------------------------------------------ Minkowski p
}  // extern "C"

namespace ksg {

static int norm_of(double p) { return p == 1.0 ? NORM_1 : (p == 2.0 ? NORM_2 : NORM_P); }

static int sub_dims_of(const uint32_t* h_masks, int n_subspaces, int dim, SubDims* out, const char* who) {
    memset(out, 0, sizeof(*out));
    const uint32_t all = dim == 32 ? 0xffffffffu : ((1u << dim) - 1u);
    for (int s = 0; s < n_subspaces; ++s) {
        if (h_masks[s] == 0 || (h_masks[s] & ~all) != 0) {
            set_error("invalid argument: %s: empty or out-of-range mask", who);
            return KSG_EINVAL;
        }
        int nd = 0;
        for (int d = 0; d < dim; ++d)
            if ((h_masks[s] >> d) & 1u) out->d[s][nd++] = (int8_t)d;
        out->n[s] = (int8_t)nd;
    }
    return KSG_OK;
}

template <int NORM, bool WITH_IDX>
static int launch_knn_pow(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                          int64_t q_begin, int64_t nq, int kp, double p, const KnnPlan& plan, double* part_dist,
                          int64_t* part_idx, cudaStream_t st) {
    size_t smem = (size_t)dim * F64_TILE * sizeof(double) + (size_t)kp * F64_THREADS * sizeof(double) +
                  (WITH_IDX ? (size_t)kp * F64_THREADS * sizeof(int64_t) : 0);
    int rc = allow_smem(knn_pow_kernel<NORM, WITH_IDX>, smem);
    if (rc) return rc;
    dim3 grid((unsigned)plan.qblocks, (unsigned)plan.splits, 1);
    knn_pow_kernel<NORM, WITH_IDX><<<grid, F64_THREADS, smem, st>>>(pts, ld, n, qry, ldq, q_begin, nq, dim, kp, p,
                                                                    plan.cand_per_split, part_dist, part_idx);
    KSG_LAUNCHED();
    return KSG_OK;
}

template <int NORM>
static int launch_count_pow(const double* pts, int64_t ld, int64_t n, int dim, const double* qry, int64_t ldq,
                            int64_t q_begin, int64_t nq, const SubDims& subs, int n_sub, const double* eps,
                            int64_t eps_stride, int strict, double p, int32_t* counts, cudaStream_t st) {
    const int64_t qblocks = ceil_div(nq, F64_THREADS);
    int splits = choose_splits(qblocks * n_sub, n, F64_TILE);
    const int64_t per = ceil_div(ceil_div(n, splits), F64_TILE) * F64_TILE;
    splits = (int)ceil_div(n, per);
    if (splits < 1) splits = 1;
    const size_t smem = (size_t)dim * F64_TILE * sizeof(double);
    int rc = allow_smem(count_pow_kernel<NORM>, smem);
    if (rc) return rc;
    dim3 grid((unsigned)qblocks, (unsigned)splits, (unsigned)n_sub);
    count_pow_kernel<NORM><<<grid, F64_THREADS, smem, st>>>(pts, ld, n, qry, ldq, q_begin, nq, dim, subs, n_sub, eps,
                                                            eps_stride, strict, p, per, counts);
    KSG_LAUNCHED();
    return KSG_OK;
}

}  // namespace ksg

extern "C" {

#define KSG_NORM_SWITCH(p, CALL)                            switch (norm_of(p)) {                                       case NORM_1: { constexpr int NORM = NORM_1; CALL; } break;         case NORM_2: { constexpr int NORM = NORM_2; CALL; } break;         default: { constexpr int NORM = NORM_P; CALL; } break;         }

int ksg_knn_radius_p(const double* d_points, int64_t ld, int64_t n, int dim, const double* d_queries, int64_t ldq,
                     int64_t q_begin, int64_t q_end, int kprime, double p, double* d_eps_out, int64_t* d_idx_out,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
    KSG_REQUIRE(d_points && d_eps_out, "ksg_knn_radius_p: null pointer");
    KSG_REQUIRE(n > 0 && ld >= n, "ksg_knn_radius_p: need n > 0 and ld >= n");
    KSG_REQUIRE(dim >= 1 && dim <= KSG_MAX_DIM, "ksg_knn_radius_p: dim out of range");
    KSG_REQUIRE(kprime >= 1 && kprime <= KSG_MAX_KPRIME, "ksg_knn_radius_p: kprime out of range");
    KSG_REQUIRE(p >= 1.0 && p <= 1.7976931348623157e308, "ksg_knn_radius_p: need 1 <= p < inf (p = inf: ksg_knn_radius)");
    if (!d_queries) {
        d_queries = d_points;
        ldq = ld;
        KSG_REQUIRE(q_end <= n, "ksg_knn_radius_p: query range exceeds the point set");
    }
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && ldq >= q_end, "ksg_knn_radius_p: bad query range");
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    KnnPlan plan = knn_plan(n, nq, kprime);
    if (workspace_bytes < plan.dist_bytes + plan.idx_bytes || !d_workspace) {
        set_error("ksg_knn_radius_p: workspace of %zu bytes needed, %zu given", plan.dist_bytes + plan.idx_bytes,
                  workspace_bytes);
        return KSG_EWORKSPACE;
    }
    double* part_dist = (double*)d_workspace;
    int64_t* part_idx = (int64_t*)((char*)d_workspace + plan.dist_bytes);
    cudaStream_t st = as_stream(stream);
    int rc = KSG_OK;
    if (d_idx_out) {
        KSG_NORM_SWITCH(p, (rc = launch_knn_pow<NORM, true>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, kprime, p,
                                                            plan, part_dist, part_idx, st)))
    } else {
        KSG_NORM_SWITCH(p, (rc = launch_knn_pow<NORM, false>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, kprime, p,
                                                             plan, part_dist, part_idx, st)))
    }
    if (rc) return rc;
    const unsigned mb = (unsigned)ceil_div(nq, 128);
    if (d_idx_out)
        knn_merge_kernel<true><<<mb, 128, 0, st>>>(part_dist, part_idx, nq, kprime, plan.splits, d_eps_out, d_idx_out);
    else
        knn_merge_kernel<false><<<mb, 128, 0, st>>>(part_dist, part_idx, nq, kprime, plan.splits, d_eps_out, nullptr);
    KSG_LAUNCHED();
    KSG_NORM_SWITCH(p, (pow_root_kernel<NORM><<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(d_eps_out, nq, p)))
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_count_subspaces_p(const double* d_points, int64_t ld, int64_t n, int dim, const double* d_queries,
                          int64_t ldq, int64_t q_begin, int64_t q_end, const uint32_t* h_masks, int n_subspaces,
                          const double* d_eps, int64_t eps_stride, int strict, int32_t bias, double p,
                          int32_t* d_counts, void* stream) {
    KSG_REQUIRE(d_points && h_masks && d_eps && d_counts, "ksg_count_subspaces_p: null pointer");
    KSG_REQUIRE(n > 0 && ld >= n, "ksg_count_subspaces_p: need n > 0 and ld >= n");
    KSG_REQUIRE(dim >= 1 && dim <= KSG_MAX_DIM, "ksg_count_subspaces_p: dim out of range");
    KSG_REQUIRE(n_subspaces >= 1 && n_subspaces <= KSG_MAX_SUBSPACES, "ksg_count_subspaces_p: subspace count");
    KSG_REQUIRE(p >= 1.0 && p <= 1.7976931348623157e308, "ksg_count_subspaces_p: need 1 <= p < inf (p = inf: ksg_count_subspaces)");
    if (!d_queries) {
        d_queries = d_points;
        ldq = ld;
        KSG_REQUIRE(q_end <= n, "ksg_count_subspaces_p: query range exceeds the point set");
    }
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && ldq >= q_end, "ksg_count_subspaces_p: bad query range");
    const int64_t nq = q_end - q_begin;
    KSG_REQUIRE(eps_stride == 0 || eps_stride >= nq, "ksg_count_subspaces_p: eps_stride must be 0 or >= nq");
    if (nq == 0) return KSG_OK;
    SubDims subs;
    int rc = sub_dims_of(h_masks, n_subspaces, dim, &subs, "ksg_count_subspaces_p");
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const int64_t total = (int64_t)n_subspaces * nq;
    fill_i32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(d_counts, total, bias);
    KSG_LAUNCHED();
    KSG_NORM_SWITCH(p, (rc = launch_count_pow<NORM>(d_points, ld, n, dim, d_queries, ldq, q_begin, nq, subs, n_subspaces,
                                                    d_eps, eps_stride, strict, p, d_counts, st)))
    return rc;
}

int ksg_subspace_radius_from_neighbours_p(const double* d_points, int64_t ld, int64_t n, int dim, int64_t q_begin,
                                          int64_t q_end, const int64_t* d_idx, int kprime, int skip_first,
                                          const uint32_t* h_masks, int n_subspaces, double p, double* d_eps_out,
                                          void* stream) {
    KSG_REQUIRE(d_points && d_idx && h_masks && d_eps_out, "ksg_subspace_radius_from_neighbours_p: null pointer");
    KSG_REQUIRE(n > 0 && ld >= n && dim >= 1 && dim <= KSG_MAX_DIM, "ksg_subspace_radius_from_neighbours_p: sizes");
    KSG_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= n, "ksg_subspace_radius_from_neighbours_p: query range");
    KSG_REQUIRE(n_subspaces >= 1 && n_subspaces <= KSG_MAX_SUBSPACES, "ksg_subspace_radius_from_neighbours_p: subspaces");
    KSG_REQUIRE(kprime >= 1 && kprime <= KSG_MAX_KPRIME, "ksg_subspace_radius_from_neighbours_p: kprime");
    KSG_REQUIRE(p >= 1.0 && p <= 1.7976931348623157e308, "ksg_subspace_radius_from_neighbours_p: need 1 <= p < inf");
    const int64_t nq = q_end - q_begin;
    if (nq == 0) return KSG_OK;
    SubDims subs;
    int rc = sub_dims_of(h_masks, n_subspaces, dim, &subs, "ksg_subspace_radius_from_neighbours_p");
    if (rc) return rc;
    KSG_NORM_SWITCH(p, (subspace_radius_pow_kernel<NORM><<<(unsigned)ceil_div(nq, 128), 128, 0, as_stream(stream)>>>(
                           d_points, ld, q_begin, nq, n, d_idx, kprime, skip_first, subs, n_subspaces, p, d_eps_out)))
    KSG_LAUNCHED();
    return KSG_OK;
}

int ksg_paircount_fixed_radius_p(const double* d_x, const double* d_y, int64_t n_templates, int m, double r, double p,
                                 int64_t i_begin, int64_t i_end, unsigned long long* d_counts, void* stream) {
    KSG_REQUIRE(d_x && d_y && d_counts, "ksg_paircount_fixed_radius_p: null pointer");
    KSG_REQUIRE(n_templates > 0 && m >= 1 && m < KSG_MAX_DIM, "ksg_paircount_fixed_radius_p: sizes");
    KSG_REQUIRE(i_begin >= 0 && i_end >= i_begin && i_end <= n_templates, "ksg_paircount_fixed_radius_p: range");
    KSG_REQUIRE(p >= 1.0 && p <= 1.7976931348623157e308, "ksg_paircount_fixed_radius_p: need 1 <= p < inf");
    const int64_t ni = i_end - i_begin;
    if (ni == 0) return KSG_OK;
    const int64_t iblocks = ceil_div(ni, PC_THREADS);
    int splits = choose_splits(iblocks, n_templates, PC_TILE);
    const int64_t per = ceil_div(ceil_div(n_templates, splits), PC_TILE) * PC_TILE;
    splits = (int)ceil_div(n_templates, per);
    if (splits < 1) splits = 1;
    dim3 grid((unsigned)iblocks, (unsigned)splits, 1);
    KSG_NORM_SWITCH(p, (paircount_pow_kernel<NORM><<<grid, PC_THREADS, 0, as_stream(stream)>>>(
                           d_x, d_y, n_templates, m, r, p, i_begin, ni, per, d_counts)))
    KSG_LAUNCHED();
    return KSG_OK;
}

}  // extern "C"

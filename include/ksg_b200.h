/*
 * ksg_b200.h -- C ABI of the B200-native KSG hot path (libksg_b200.so).
 *
 * The reference (thosvarley/syntropy, package `syntropy.knn`) has no FFI of its
 * own: the path is pure Python over scipy's compiled cKDTree.  The entry points
 * below are therefore what a binding for this path has to offer the reference's
 * Python call sites; each one names the reference interface it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * stub a maintainer would add on the reference side.
 *
 * Conventions
 *  - plain C types only; every buffer is a caller-owned DEVICE pointer unless the
 *    name starts with h_ (small host arrays read before launch);
 *  - point sets are float64, structure-of-arrays: coordinate d of point j lives at
 *    base[d * ld + j]  (`data` of shape (n_variables, n_samples), row stride ld);
 *  - every function returns 0 on success or a negative KSG_E* code, never throws;
 *    ksg_last_error() gives the message for the calling thread;
 *  - no hidden allocation: scratch comes from the caller (size from *_workspace);
 *  - `stream` is a cudaStream_t passed as void*; work is only enqueued, the caller
 *    synchronises;
 *  - there is no CPU fallback anywhere: without a CUDA device every compute entry
 *    point returns KSG_ECUDA;
 *  - threading: ONE host thread per device at a time.  ksg_prepare* share per-device helper
 *    streams, fork / join events and a graph cache; two threads (or two user streams)
 *    preparing on the same device concurrently must be serialised by the caller.  Different
 *    devices (one process or thread per GPU) are independent.
 */
#ifndef KSG_B200_H
#define KSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSG_OK 0
#define KSG_EINVAL (-1)   /* bad argument */
#define KSG_ECUDA (-2)    /* CUDA runtime error / no device */
#define KSG_EWORKSPACE (-3) /* workspace too small */
#define KSG_EUNSUPPORTED (-4)

#define KSG_MAX_DIM 32        /* joint-space dimensionality limit */
#define KSG_MAX_KPRIME 128    /* k+1 limit */
#define KSG_MAX_SUBSPACES 64  /* subspaces per count call */

/* ABI version (major*1000 + minor). */
int ksg_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* ksg_last_error(void);

/* Number of kernels this library has launched in this process so far (bench.py's
 * gpu_launches is a difference of two readings). */
unsigned long long ksg_launch_count(void);

/* SM count, compute capability and SM clock (kHz) of the current device. */
int ksg_device_info(int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz);

/*
 * Sets *d_flag (device int32, caller-zeroed) to 1 if any of the `count` values is
 * NaN or +-inf.  Replaces the ValueError scipy's cKDTree constructor raises on
 * non-finite input (reached from syntropy/knn/utils.py:29).
 */
int ksg_check_finite(const double* d_values, int64_t count, int32_t* d_flag, void* stream);

/*
 * d_table[i] = digamma(i) for i in [0, n_entries): table[0] = -inf, harmonic sum
 * minus Euler's constant for i <= 10, the cephes asymptotic series above.
 * Replaces scipy.special.digamma at the reference's call sites, which only ever
 * pass integers (counts+1, k, N): syntropy/knn/shannon.py:74-75,182-183,194,323,
 * 342-344; syntropy/knn/multivariate_mi.py:97-98,113,238-239,257,305,329-330,388,
 * 409-410.
 */
int ksg_psi_table(double* d_table, int64_t n_entries, void* stream);

/* ------------------------------------------------------------------ K1: radius
 *
 * eps[i - q_begin] = kprime-th smallest Chebyshev distance
 *                    max_d |fl64(q_id - p_jd)|  over ALL points j in [0, n)
 * for queries i in [q_begin, q_end).  d_queries == NULL means "the point set
 * itself" (self is then one of the candidates, at distance 0, exactly as in
 * tree.query(data.T, k=k+1)); kprime = k+1 within-set, k cross-set.  If kprime > n
 * the radius is +inf.  Replaces build_tree_and_get_distances
 * (syntropy/knn/utils.py:6-35: cKDTree(data.T) :29 + tree.query(..., k=k+1, p=inf)
 * :33) as used for `distances[:, -1]`, and the two cKDTree queries of
 * kullback_leibler_divergence (syntropy/knn/shannon.py:401-402).
 *
 * d_idx_out (optional, may be NULL): int64 [nq][kprime], neighbour indices in
 * ascending (distance, index) order -- `indices` of the same reference call, needed
 * by the KSG-2 variants (syntropy/knn/shannon.py:249; multivariate_mi.py:169).
 */
size_t ksg_knn_radius_workspace(int64_t n, int64_t nq, int dim, int kprime);
int ksg_knn_radius(const double* d_points, int64_t ld, int64_t n, int dim,
                   const double* d_queries, int64_t ldq,
                   int64_t q_begin, int64_t q_end, int kprime,
                   double* d_eps_out, int64_t* d_idx_out,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ K2: counts
 *
 * For every subspace s (bit d of h_masks[s] set = joint dimension d belongs to s)
 * and every query i in [q_begin, q_end):
 *   counts[s * nq + (i - q_begin)] =
 *        #{ j in [0,n) : max_{d in s} |fl64(q_id - p_jd)| <  eps }  + bias   (strict)
 *        #{ ...                                           <= eps }  + bias   (closed)
 * with eps = d_eps[s * eps_stride + (i - q_begin)]  (eps_stride = 0: one radius
 * vector shared by all subspaces, the KSG-1 family; eps_stride = nq: one per
 * subspace, the KSG-2 family).  bias = -1 removes the query itself, as
 * get_counts_from_tree does (syntropy/knn/utils.py:38-86: nextafter :75-76, the
 * per-sample query_ball_point(..., return_length=True) - 1 loop :79-84).  One call
 * serves every marginal an estimator needs from the single joint radius
 * (shannon.py:189-194,335-344; multivariate_mi.py:103-113,247-257,314-330,396-410).
 */
size_t ksg_count_subspaces_workspace(int64_t n, int64_t nq, int dim, int n_subspaces);
int ksg_count_subspaces(const double* d_points, int64_t ld, int64_t n, int dim,
                        const double* d_queries, int64_t ldq,
                        int64_t q_begin, int64_t q_end,
                        const uint32_t* h_masks, int n_subspaces,
                        const double* d_eps, int64_t eps_stride,
                        int strict, int32_t bias,
                        int32_t* d_counts,
                        void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ K3: psi epilogue
 *
 * Pointwise estimator values in the reference's float64 operation order:
 *   v = init
 *   for each step t (in order):
 *       p    = psi_table[ clamp(counts[step.subspace][i] + step.count_offset, 0, ..) ]
 *       term = step.centered ? (p - psi_n) / step.divisor : p
 *       v    = step.sign > 0 ? v + term : v - term
 *   if (apply_scale) v = v * scale
 *   ptw[i] = v
 * which covers mutual_information_1/_2 (shannon.py:188-194,251-263),
 * conditional_mutual_information (:329-344), total_correlation_1/_2
 * (multivariate_mi.py:100-113,164-191), dual_total_correlation (:241-258),
 * s_information (:311-333) and o_information (:394-413).
 */
typedef struct {
    int32_t subspace;     /* row of d_counts */
    int32_t count_offset; /* +1 for the KSG-1 family, 0 for KSG-2 */
    int32_t sign;         /* +1: add the term, -1: subtract it */
    int32_t centered;     /* 0: psi(c+off); 1: (psi(c+off) - psi_n) / divisor */
    double divisor;
} ksg_psi_step;

int ksg_psi_epilogue(const int32_t* d_counts, int64_t nq, int n_subspaces,
                     const ksg_psi_step* h_steps, int n_steps,
                     double init, double psi_n, int apply_scale, double scale,
                     const double* d_psi_table, int64_t table_entries,
                     double* d_ptw, void* stream);

/*
 * Peer-memory exchange of the sharded path (one process per GPU; SURVEY.md 8e: query rows shard, every rank
 * needs the whole pointwise vector -- the reference has no counterpart, its ptw array simply lives in one
 * address space, shannon.py:188-196).  The producing kernel stores its results straight into EVERY rank's
 * buffer over NVLink (h_peer_*: host arrays of `world` device pointers, the peers' memory mapped into this
 * process by the caller, own rank included) and releases one flag per peer when its last block is done; the
 * consumer acquires the `world` flags of its own array.  Flags carry the epoch of the exchange (they only
 * grow, nothing is reset); the caller double-buffers the payload by the epoch's parity.  d_ticket: one int32,
 * 0 on entry, 0 again on exit.  No library collective, no host synchronisation.
 *   ksg_psi_epilogue_push  ksg_psi_epilogue whose values land at [q_begin, q_begin + nq) of every rank's
 *                          vector; *d_nflag (this rank's unsettled-query counter, may be null) goes to
 *                          slot `rank` of every rank's aux array;
 *   ksg_p2p_push           `bytes` of device memory to byte offset dst_offset of every rank's buffer (multiples
 *                          of 8); several pushes of one exchange use consecutive epochs, the consumer waits for the last;
 *   ksg_p2p_wait           returns (in stream order) once all `world` flags have reached `epoch`;
 *                          *d_aux_sum = sum of the aux slots (every rank gets the same number).
 */
int ksg_psi_epilogue_push(const int32_t* d_counts, int64_t nq, int n_subspaces, const ksg_psi_step* h_steps, int n_steps,
                          double init, double psi_n, int apply_scale, double scale, const double* d_psi_table,
                          int64_t table_entries, void* const* h_peer_values, int64_t q_begin,
                          const int32_t* d_nflag, int32_t* const* h_peer_aux, unsigned long long* const* h_peer_flags,
                          int rank, int world, unsigned long long epoch, int32_t* d_ticket, void* stream);
int ksg_p2p_push(const void* d_src, size_t bytes, void* const* h_peer_dst, size_t dst_offset,
                 unsigned long long* const* h_peer_flags, int rank, int world, unsigned long long epoch,
                 int32_t* d_ticket, void* stream);
int ksg_p2p_wait(const unsigned long long* d_flags, int world, unsigned long long epoch, const int32_t* d_aux,
                 int32_t* d_aux_sum, void* stream);

/*
 * ptw[i] = 0 + ((-psi_k + psi_n) + dim * log(2 * eps[i]))   -- Kozachenko-Leonenko,
 * syntropy/knn/shannon.py:84-85.
 */
int ksg_entropy_epilogue(const double* d_eps, int64_t nq, int dim, double psi_k, double psi_n,
                         double* d_ptw, void* stream);

/*
 * ptw[i] = dim * log(nu[i] / rho[i]) + log_ratio   -- kNN KL divergence,
 * syntropy/knn/shannon.py:407 (log_ratio = log(m / (n - 1)) computed by the host).
 */
int ksg_kl_epilogue(const double* d_nu, const double* d_rho, int64_t nq, int dim, double log_ratio,
                    double* d_ptw, void* stream);

/*
 * Deterministic float64 sum of n values into d_out[0] (fixed-shape tree, independent
 * of launch geometry).  Replaces ptw.mean()'s sum (shannon.py:87,196,347;
 * multivariate_mi.py:115,260,335,415); the caller divides by n.
 * Workspace: ksg_sum_workspace(n) bytes.
 */
size_t ksg_sum_workspace(int64_t n);
int ksg_sum_f64(const double* d_values, int64_t n, double* d_out,
                void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ KSG-2 helper
 *
 * eps_out[s * nq + i] = max_{j < k} max_{d in s} |p_id - p_{nbr(i,j), d}|  for the k
 * joint-space neighbours nbr(i, 1..k) of each query (column 0, the point itself, is
 * skipped by the caller passing idx + 1 semantics via `skip_first`).
 * Replaces the per-marginal radius loops of mutual_information_2
 * (syntropy/knn/shannon.py:253-259) and total_correlation_2
 * (syntropy/knn/multivariate_mi.py:176-183).
 */
int ksg_subspace_radius_from_neighbours(const double* d_points, int64_t ld, int64_t n, int dim,
                                        int64_t q_begin, int64_t q_end,
                                        const int64_t* d_idx, int kprime, int skip_first,
                                        const uint32_t* h_masks, int n_subspaces,
                                        double* d_eps_out, void* stream);

/* ------------------------------------------------------------------ K4: fixed-radius pair count
 *
 * Templates u_i = (x[i], ..., x[i+m-1]) and v_j = (y[j], ..., y[j+m-1]) ("small"),
 * and their (m+1)-long extensions ("big"), i, j in [0, n_templates).
 *   d_counts[0] += #{(i, j), i in [i_begin, i_end) : max_w |x[i+w] - y[j+w]| <= r, w < m}
 *   d_counts[1] += the same with w <= m
 * (closed ball, ordered pairs, self pairs included; d_counts is caller-zeroed
 * uint64[2]).  Both series need n_templates + m readable values.  Replaces the two
 * cKDTree.count_neighbors calls of sample_entropy (syntropy/knn/temporal.py:65-72)
 * and cross_sample_entropy (:136-143).
 */
int ksg_paircount_fixed_radius(const double* d_x, const double* d_y, int64_t n_templates, int m,
                               double r, int64_t i_begin, int64_t i_end,
                               unsigned long long* d_counts, void* stream);


/* ================================================================== fast path
 *
 * The entry points above are the exact float64 dense kernels (correctness anchor and
 * fallback).  The ones below are the production path: same results, bit for bit, from
 * an fp32 tile filter whose boundary cases are re-decided in float64 on the device.
 * They work on a "prepared" point set built once per estimator call.
 */

/* Device-side views of a prepared point set (all pointers into the caller's buffer). */
typedef struct {
    int64_t n;           /* points */
    int32_t dim;         /* joint dimensionality D */
    int32_t dim_padded;  /* D rounded up to even: row length of shadow_f32 */
    int32_t primary;     /* coordinate the set is sorted by; -1: Hilbert-curve order of the per-coordinate ranks; -2: kd order */
    int32_t axes_mode;   /* 0: axis_vals / axis_pos are the exact per-coordinate sorts.  b > 0 ("lazy axes", built with
                            primary == -3): no coordinate was sorted; axis_vals[d] holds coordinate d's values GROUPED
                            into 2^b value buckets (any order inside a bucket) and axis_pos[d] the buckets' prefix sums
                            (2^b + 1 entries) -- what ksg_axis_count_buckets reads; ksg_prepare_axes turns the set into
                            an ordinary one (axes_mode 0) when the sorted arrays are needed after all */
    int64_t ld;          /* row stride (elements) of the [dim][ld] arrays */
    int64_t n_padded;    /* rows of shadow_f32: n rounded up to the tile length, +inf padded */
    int32_t* perm;       /* [n]        sorted position -> original sample index */
    double* sorted_f64;  /* [dim][ld]  the points in primary-sorted order */
    float* shadow_f32;   /* [n_padded][dim_padded] centred fp32 copy, array-of-structures */
    double* axis_vals;   /* [dim][ld]  every coordinate sorted ascending on its own */
    int32_t* axis_pos;   /* [dim][ld]  sorted position (primary order) of each axis_vals entry */
    double* center;      /* [dim]      per-coordinate centre of the fp32 copy */
    double* maxabs;      /* [dim]      max |x - centre| (bounds the fp32 rounding error) */
    float* tile_box;     /* [n_padded/512][dim_padded][2] min / max of shadow_f32 over each 512-row tile */
    float* sub_box;      /* [n_padded/512][16][dim_padded][2] the same over each 32-row sub-tile */
    int32_t* flags;      /* 64 bytes: int32 [0] = 1 if the input held a NaN or an infinity;
                            uint64 at byte 16 / 24 = (query, candidate) pairs evaluated so far by
                            the k-NN / count filter kernels (for the roofline of the windowed
                            engine, which skips most of the n x n pairs) */
    uint64_t* curve_keys; /* [n] ascending space-filling-curve keys of the rows (primary == -1 only; else NULL):
                            lets ksg_cross_prepare place foreign query points on the same curve */
    float* curve_table;   /* [dim][1026] (primary == -1, early order; else NULL): per coordinate the centred fp32
                            [min, sorted sample of 1024 values, max] -- the monotone map behind the curve coordinates */
    void* axes_ready;     /* cudaEvent_t or NULL.  Early order (primary == -1): the curve is built from the
                            table above, and the EXACT coordinate sorts behind axis_vals / axis_pos run on
                            side streams of the library underneath whatever the caller enqueues next
                            (the k-NN stage).  Every entry point that reads them (ksg_axis_count,
                            ksg_fast_count) and ksg_fast_knn* order the caller's stream after this event
                            themselves; a caller that frees or reuses the buffer without having called any
                            of them calls ksg_prepare_join first. */
} ksg_prepared;

/*
 * Order the n points of d_points ([dim][ld] float64) and build the views above inside d_buffer
 * (ksg_prepare_workspace(n, dim) bytes).  primary >= 0: the order of that coordinate;
 * primary == -1: the Hilbert curve over the per-coordinate ranks (spatially compact 512-point
 * tiles and 32-point sub-tiles in every coordinate); primary == -2: the leaf order of a balanced
 * kd-tree over the ranks (conditional medians: the better order from ~7 coordinates up);
 * primary == -3: the curve order with LAZY axes (ksg_prepared::axes_mode): no coordinate is sorted, the values
 * are grouped into value buckets instead -- enough for ksg_fast_knn* and ksg_axis_count_buckets, i.e. for the
 * mutual information of scalar series; everything else needs ksg_prepare_axes first.
 * *h_out is filled on the host; the device work is only enqueued (the per-coordinate sorts run
 * concurrently on helper streams of the library, forked from and joined to `stream`).
 * Replaces the cKDTree constructions of the path (syntropy/knn/utils.py:29 and every
 * per-marginal tree: shannon.py:190,336; multivariate_mi.py:106,251,318-323,399-404) -- one
 * sorted layout serves the joint space and all marginals.
 *
 * ksg_prepare_rows: the same, for rows that are still on their way to the device.
 * h_row_ready (host array of dim cudaEvent_t handles, entries may be NULL) names, per coordinate,
 * the event after which d_points[d][*] is valid -- typically recorded on a copy stream after the
 * host-to-device copy of that row.  The sort of coordinate d waits for its own event only, so it
 * overlaps the transfer of the rows behind it; everything `stream` does after the call is ordered
 * after all the events.
 */
size_t ksg_prepare_workspace(int64_t n, int dim);
/* Order `stream` after every piece of library-side work on the prepared set (see axes_ready). */
int ksg_prepare_join(const ksg_prepared* h_prep, void* stream);
int ksg_prepare(const double* d_points, int64_t ld, int64_t n, int dim, int primary,
                void* d_buffer, size_t buffer_bytes, ksg_prepared* h_out, void* stream);
int ksg_prepare_rows(const double* d_points, int64_t ld, int64_t n, int dim, int primary,
                     void* const* h_row_ready, void* d_buffer, size_t buffer_bytes, ksg_prepared* h_out,
                     void* stream);

/*
 * K1 on the prepared set: d_eps_out[i - q_begin] = exact kprime-th smallest Chebyshev
 * distance (self included) of the point at SORTED position i, i in [q_begin, q_end).
 * d_query_index (optional, device int32[q_end - q_begin]): explicit sorted positions of
 * the queries instead of the contiguous range (q_begin is then ignored; always dense).
 * windowed != 0 prunes candidate tiles by their bounding boxes (same result, fewer pairs).
 * Per query, d_flags[i] is 0 (done), 1 (the fp32 filter could not decide -- tie-saturated
 * neighbourhood: recompute with ksg_knn_radius on that row of sorted_f64) or 2 (an
 * isolated outlier the windowed pass set aside: recompute with a second ksg_fast_knn call
 * passing the flagged positions as d_query_index); *d_n_flagged (caller-zeroed) counts
 * the non-zero flags.
 * Same reference call sites as ksg_knn_radius.
 */
size_t ksg_fast_knn_workspace(int64_t n, int64_t nq, int dim, int kprime);
int ksg_fast_knn(const ksg_prepared* h_prep, int64_t q_begin, int64_t q_end,
                 const int32_t* d_query_index, int kprime, int windowed,
                 double* d_eps_out, int32_t* d_flags, int32_t* d_n_flagged,
                 void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * The same with the neighbours' identities (KSG algorithm 2, syntropy/knn/shannon.py:248-259,
 * multivariate_mi.py:170-183: `indices[:, 1:]` of tree.query): d_idx_out[(i - q_begin) * kprime + r]
 * = position in the prepared order of the r-th nearest point (r = 0 is the query itself unless
 * it has exact duplicates), ascending by (exact distance, original sample index).  Queries with
 * fewer than kprime points are flagged 1.  d_idx_out may be NULL (then identical to ksg_fast_knn).
 */
int ksg_fast_knn_idx(const ksg_prepared* h_prep, int64_t q_begin, int64_t q_end, const int32_t* d_query_index,
                     int kprime, int windowed, double* d_eps_out, int64_t* d_idx_out, int32_t* d_flags,
                     int32_t* d_n_flagged, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * The same with an explicit list margin: the filter keeps the k' + margin smallest fp32 distances
 * per query (ksg_fast_knn / ksg_fast_knn_idx use margin = 3).  A query is decided exactly when
 * fewer than `margin` further candidates lie inside the fp32 error band of its k'-th distance;
 * quantised data with many exact ties needs a longer list (k' + margin <= 64; the lists live in
 * shared memory, so large values cost occupancy).  Workspace: ksg_fast_knn_workspace_margin.
 */
size_t ksg_fast_knn_workspace_margin(int64_t n, int64_t nq, int dim, int kprime, int margin);
int ksg_fast_knn_margin(const ksg_prepared* h_prep, int64_t q_begin, int64_t q_end, const int32_t* d_query_index,
                        int kprime, int margin, int windowed, double* d_eps_out, int64_t* d_idx_out,
                        int32_t* d_flags, int32_t* d_n_flagged, void* d_workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------ cross-set k-NN (KL divergence)
 *
 * Queries that are NOT points of the prepared (candidate) set: kullback_leibler_divergence asks
 * for the k-th neighbour of every posterior sample among the prior samples
 * (syntropy/knn/shannon.py:401-402, `cKDTree(prior).query(posterior, k)`).  ksg_cross_prepare puts
 * the nq query points (d_queries, [dim][ldq] float64) on the candidate set's curve: rank of every
 * coordinate among the candidates' (binary search), same Hilbert key, radix sort; it stores them
 * in that order as float64 SoA and as fp32 rows in the candidate set's centred frame, together
 * with the candidate row next to each of them on the curve.  Needs primary == -1.
 */
typedef struct {
    int64_t nq;          /* queries */
    int64_t ld;          /* row stride of sorted_f64 */
    int32_t* perm;       /* [nq] position in the cross order -> original query index */
    double* sorted_f64;  /* [dim][ld] the queries in cross order */
    float* shadow_f32;   /* [nq][dim_padded] fp32(x - candidate centre) */
    int32_t* home;       /* [nq] candidate row (prepared order) next to the query on the curve */
    double* maxabs;      /* [dim] max over BOTH sets of |x - centre| (fp32 error bound) */
} ksg_cross;
size_t ksg_cross_prepare_workspace(int64_t n_candidates, int64_t nq, int dim);
int ksg_cross_prepare(const ksg_prepared* h_candidates, const double* d_queries, int64_t ldq, int64_t nq,
                      void* d_buffer, size_t buffer_bytes, ksg_cross* h_out, void* stream);

/*
 * d_eps_out[i - q_begin] = exact kprime-th smallest Chebyshev distance from the query at cross
 * position i to the candidate set (no self: the query is not a candidate), i in [q_begin, q_end).
 * Always windowed.  Flags as in ksg_fast_knn; flagged queries (1 or 2) are to be recomputed with
 * ksg_knn_radius (queries = those rows of sorted_f64).  Workspace: ksg_fast_knn_workspace(n, nq, ...).
 */
int ksg_fast_knn_cross(const ksg_prepared* h_candidates, const ksg_cross* h_queries, int64_t q_begin,
                       int64_t q_end, int kprime, double* d_eps_out, int32_t* d_flags, int32_t* d_n_flagged,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * K2 on the prepared set, multi-dimensional subspaces: semantics of ksg_count_subspaces
 * with one shared radius vector (d_eps, sorted order, exact float64), queries = sorted
 * positions [q_begin, q_end).  Recognised subspace structures (x|y, z|xz|yz,
 * leave-one-out) run fused kernels that share partial maxima; anything else runs four
 * masks at a time.  Flagging as in ksg_fast_knn (fallback: ksg_count_subspaces).
 * Same reference call sites as ksg_count_subspaces.
 */
size_t ksg_fast_count_workspace(int64_t n, int64_t nq, int dim, int n_subspaces);
int ksg_fast_count(const ksg_prepared* h_prep, int64_t q_begin, int64_t q_end,
                   const uint32_t* h_masks, int n_subspaces, const double* d_eps,
                   int strict, int32_t bias, int windowed,
                   int32_t* d_counts, int32_t* d_flags, int32_t* d_n_flagged,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * K2 for a ONE-dimensional subspace (coordinate `axis`): the count is the length of the
 * run around x_q in that coordinate's sorted array satisfying |fl64(x_q - v)| < eps
 * (<= when strict == 0), found by two binary searches with the exact predicate.
 * d_counts[i - q_begin] = run length + bias.  Always exact, never flags.
 */
int ksg_axis_count(const ksg_prepared* h_prep, int axis, int64_t q_begin, int64_t q_end,
                   const double* d_eps, int strict, int32_t bias, int32_t* d_counts, void* stream);

/* Several ONE-dimensional subspaces in one launch (the searches are latency-bound: two coordinates side
 * by side cost little more than one): d_counts[s * nq + (i - q_begin)] for coordinate h_axes[s], nq =
 * q_end - q_begin.  Both marginals of mutual_information (shannon.py:190-194), every single of
 * total_correlation / o_information (multivariate_mi.py:106-113,399-404). */
int ksg_axis_count_multi(const ksg_prepared* h_prep, const int32_t* h_axes, int n_axes, int64_t q_begin,
                         int64_t q_end, const double* d_eps, int strict, int32_t bias, int32_t* d_counts,
                         void* stream);

/* The same counts for a prepared set with LAZY axes (ksg_prepare* with primary == -3, axes_mode > 0): no
 * coordinate is sorted; whole value buckets inside the ball are counted by prefix sums, the buckets its two ends
 * fall into value by value with the exact float64 predicate (csrc/axis_buckets.cuh has the exactness argument).
 * *d_n_flagged is incremented when the counts could NOT all be settled this way (tied data fills single buckets
 * with thousands of equal values): the caller then calls ksg_prepare_axes and ksg_axis_count_multi.  Takes the
 * two float64 radix sorts per call off the mutual-information path (shannon.py:190-194). */
int ksg_axis_count_buckets(const ksg_prepared* h_prep, const int32_t* h_axes, int n_axes, int64_t q_begin,
                           int64_t q_end, const double* d_eps, int strict, int32_t bias, int32_t* d_counts,
                           int32_t* d_n_flagged, void* stream);
/* Build the exact per-coordinate sorts of a lazy set after all (in stream order, from the set's own float64
 * copy); h_prep->axes_mode becomes 0.  No-op for an ordinary set. */
int ksg_prepare_axes(ksg_prepared* h_prep, void* stream);

/* The tail of an estimator step in ONE launch, for values of ALL n samples in the prepared order:
 *   ksg_psi_finish      v_i = the psi program of ksg_psi_epilogue over d_counts[s * n + i],
 *   ksg_scatter_sum_f64 v_i = d_values_sorted[i] (e.g. all-gathered slices of several ranks),
 * then d_out[perm[i]] = v_i (the caller's sample order: the `ptw` the reference returns) and
 * d_total[0] = the deterministic float64 sum of all v_i (ptw.mean() = d_total / n; a fixed tree over
 * the prepared order, independent of launch geometry and of the number of ranks).  d_ticket: one
 * int32 that is 0 on entry and 0 again on exit (the block that finishes last folds the partial sums).
 * Workspace: ksg_finish_workspace(n) bytes. */
size_t ksg_finish_workspace(int64_t n);
int ksg_psi_finish(const int32_t* d_counts, int64_t n, int n_subspaces, const ksg_psi_step* h_steps, int n_steps,
                   double init, double psi_n, int apply_scale, double scale, const double* d_psi_table,
                   int64_t table_entries, const int32_t* d_perm, double* d_out, double* d_total,
                   int32_t* d_ticket, void* d_workspace, size_t workspace_bytes, void* stream);
int ksg_scatter_sum_f64(const double* d_values_sorted, const int32_t* d_perm, int64_t n, double* d_out,
                        double* d_total, int32_t* d_ticket, void* d_workspace, size_t workspace_bytes,
                        void* stream);

/* d_dst[perm[i]] = d_src[i], i in [0, n): pointwise values back to the caller's sample order. */
int ksg_scatter_f64(const double* d_src, const int32_t* d_perm, int64_t n, double* d_dst, void* stream);

/* The same for int32 counts, and its inverse d_dst[i] = d_src[perm[i]] for float64 radii: a
 * multi-dimensional marginal is counted in a prepared set of its OWN coordinates (its own
 * space-filling order prunes far better than the joint order does for that marginal), so the
 * joint radii travel into that order and the counts travel back to the caller's sample order.
 * (The reference builds one cKDTree per marginal for the same reason: shannon.py:190,336;
 * multivariate_mi.py:251,318-323,399-404.) */
int ksg_scatter_i32(const int32_t* d_src, const int32_t* d_perm, int64_t n, int32_t* d_dst, void* stream);
int ksg_gather_f64(const double* d_src, const int32_t* d_perm, int64_t n, double* d_dst, void* stream);

/* d_out[0] += sum of n int32 counts (integer, order-independent).  With ksg_prepare +
 * ksg_fast_count(one subspace, strict = 0, constant radius) on the template matrix this is the
 * windowed form of cKDTree.count_neighbors(self, r) (syntropy/knn/temporal.py:71-72). */
int ksg_sum_i32(const int32_t* d_values, int64_t n, unsigned long long* d_out, void* stream);

/* ---- Minkowski p-norms, 1 <= p < inf (the `p` / `norm` arguments of syntropy.knn: knn/shannon.py:15,96,274,354;
 * knn/temporal.py:16,85).  Same contracts as ksg_knn_radius / ksg_count_subspaces /
 * ksg_subspace_radius_from_neighbours / ksg_paircount_fixed_radius with the distance of scipy's cKDTree for
 * that p (p-th-power space: sum |d|, sum d^2 with its four-way partial sums, sum pow(|d|, p); a ball of radius r
 * holds s <= r, r*r, pow(r, p); strict = the radius nudged down with nextafter first, knn/utils.py:75-76).
 * Exact float64, dense; bit-exact for p = 1 and p = 2, device pow() for other p.  p = inf is rejected here
 * (use the functions above).  ksg_knn_radius_p takes the workspace of ksg_knn_radius_workspace. */
int ksg_knn_radius_p(const double* d_points, int64_t ld, int64_t n, int dim, const double* d_queries, int64_t ldq,
                     int64_t q_begin, int64_t q_end, int kprime, double p, double* d_eps_out, int64_t* d_idx_out,
                     void* d_workspace, size_t workspace_bytes, void* stream);
int ksg_count_subspaces_p(const double* d_points, int64_t ld, int64_t n, int dim, const double* d_queries,
                          int64_t ldq, int64_t q_begin, int64_t q_end, const uint32_t* h_masks, int n_subspaces,
                          const double* d_eps, int64_t eps_stride, int strict, int32_t bias, double p,
                          int32_t* d_counts, void* stream);
int ksg_subspace_radius_from_neighbours_p(const double* d_points, int64_t ld, int64_t n, int dim, int64_t q_begin,
                                          int64_t q_end, const int64_t* d_idx, int kprime, int skip_first,
                                          const uint32_t* h_masks, int n_subspaces, double p, double* d_eps_out,
                                          void* stream);
int ksg_paircount_fixed_radius_p(const double* d_x, const double* d_y, int64_t n_templates, int m, double r, double p,
                                 int64_t i_begin, int64_t i_end, unsigned long long* d_counts, void* stream);

/* The key sort ksg_prepare is built on (argsort of one coordinate; reference: the ordering work
 * cKDTree's constructor does at syntropy/knn/utils.py:29): d_keys_out = the n float64 keys ascending,
 * d_index_out = their positions in d_keys_in, ties in position order (a stable sort).
 * algorithm 1: cub::DeviceRadixSort -- what ksg_prepare uses (eight latency-bound passes per float64
 * coordinate).  algorithm 0: the five-launch sample sort of csrc/sample_sort.cuh for n <= 2^21 (radix sort
 * above); measured slower on B200 at n = 1e6 (profiles/r2b_*), kept as an independent cross-check of the
 * order.  Workspace: ksg_sort_workspace(n). */
size_t ksg_sort_workspace(int64_t n);
int ksg_sort_f64(const double* d_keys_in, int64_t n, double* d_keys_out, int32_t* d_index_out, int algorithm,
                 void* d_workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KSG_B200_H */

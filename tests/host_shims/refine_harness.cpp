// Host-side harness for knn_refine_kernel (tests/test_refine_host.py): the kernel's CUDA source is
// plain C++ apart from a few keywords, so it is compiled here for the CPU -- one call per "thread" --
// and checked against brute force in float64.  The filter's partial lists are emulated: the
// candidate set is cut into disjoint chunks (work units / splits), every chunk keeps the kcap
// smallest fp32 distances to the query, ascending, exactly what the filter kernels emit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <vector>
#define __global__
#define __device__
#define __host__
#define __restrict__
#define __launch_bounds__(x)
#define __noinline__
#define __ldcg(p) (*(p))
#define __forceinline__ inline
#define CUDART_INF (std::numeric_limits<double>::infinity())
#define CUDART_INF_F (std::numeric_limits<float>::infinity())
#define CUDART_NAN (std::numeric_limits<double>::quiet_NaN())
struct Dim3 { unsigned x; };
static Dim3 blockIdx{0}, blockDim{128}, threadIdx{0};
static inline int atomicAdd(int32_t* p, int v) { int o = *p; *p += v; return o; }
using std::fabs;
using std::fmax;
constexpr int KSG_MAX_DIM = 16;
constexpr int KNN_MAX_KCAP = 64;
// round-up float arithmetic of tight_seed(), emulated through double (exact for + and * of two floats
// up to the final rounding, which is then pushed up)
static inline float round_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}
static inline float __fmul_ru(float a, float b) { return round_up((double)a * (double)b); }
static inline float __fadd_ru(float a, float b) { return round_up((double)a + (double)b); }
static inline float __double2float_ru(double x) { return round_up(x); }
static inline float __double2float_rd(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
using std::fmaxf;
using std::fminf;
// RefineArgs (csrc/fast_launch.h) + band_query_f64 + tight_seed + constants + knn_refine_kernel +
// refine_single_list, cut out of csrc/fast_knn.cuh (+ refine_unsorted_list of csrc/fast_knn_warp.cuh)
#include "refine_kernel.inc"

// refine_single_list on the list the windowed filter leaves in shared memory for a block with ONE
// work unit: every candidate with d32 < seed, the kcap smallest of them, ascending.  The seed is
// tight_seed(s, mq) with s an upper bound of the k'-th smallest d32 (k'-th .. (k'+2)-th smallest).
static int single_list_trials(int trials, long* checked_out, long* flagged_cont_out, long* flagged_tied_out) {
    std::mt19937_64 rng(77);
    long checked = 0, wrong = 0, flagged_cont = 0, flagged_tied = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int dim = 1 + (int)(rng() % 8);
        const int kp = 1 + (int)(rng() % 10);
        const int kcap = kp + ((rng() % 4 == 0) ? 24 : 3);
        const int64_t n = 300 + (int64_t)(rng() % 600);
        const int64_t nq = std::min<int64_t>(n, 64 + (int64_t)(rng() % 100));
        const bool want_idx = rng() % 2 == 0;
        const bool tied = rng() % 4 == 0;
        const bool cross = rng() % 5 == 0;  // queries that are not members of the candidate set
        const double scale = (rng() % 3 == 0) ? 1e-9 : 1.0;
        std::vector<double> pts((size_t)dim * n), qry((size_t)dim * nq), center(dim);
        std::normal_distribution<double> nd(0.0, 1.0);
        for (int d = 0; d < dim; ++d) {
            const double shift = (rng() % 3 == 0) ? 1e3 * scale : 0.0;
            center[d] = shift;
            for (int64_t i = 0; i < n; ++i) {
                double v = nd(rng);
                if (tied) v = std::round(v * 4) / 4;
                pts[(size_t)d * n + i] = v * scale + shift;
            }
            for (int64_t i = 0; i < nq; ++i) {
                double v = nd(rng);
                if (tied) v = std::round(v * 4) / 4;
                qry[(size_t)d * nq + i] = cross ? v * scale + shift : pts[(size_t)d * n + i];
            }
        }
        std::vector<int32_t> perm(n);
        for (int64_t i = 0; i < n; ++i) perm[i] = (int32_t)i;
        std::shuffle(perm.begin(), perm.end(), rng);
        std::vector<double> eps(nq, -7.0);
        std::vector<int32_t> flags(nq, -7);
        std::vector<int64_t> nbr((size_t)nq * kp, -7);
        int32_t n_flagged = 0;
        RefineArgs ra;
        ra.fused = 1; ra.kp = kp; ra.dim = dim; ra.sorted = pts.data(); ra.ld = n;
        ra.qsorted = cross ? qry.data() : nullptr; ra.qld = nq; ra.center = center.data(); ra.perm = perm.data();
        std::vector<double> maxabs(dim, 1e6);
        ra.maxabs = maxabs.data();
        ra.exact32 = nullptr;
        ra.idx_out = want_idx ? nbr.data() : nullptr; ra.eps_out = eps.data(); ra.flags = flags.data();
        ra.n_flagged = &n_flagged;
        int counted = 0;
        for (int64_t qi = 0; qi < nq; ++qi) {
            std::vector<std::pair<float, int32_t>> all32(n);
            std::vector<std::pair<double, int32_t>> all64(n);
            float mq = 0.0f;
            for (int d = 0; d < dim; ++d) mq = std::max(mq, std::fabs((float)(qry[(size_t)d * nq + qi] - center[d])));
            for (int64_t j = 0; j < n; ++j) {
                float m32 = 0.0f;
                double m64 = 0.0;
                for (int d = 0; d < dim; ++d) {
                    const float x = (float)(qry[(size_t)d * nq + qi] - center[d]), y = (float)(pts[(size_t)d * n + j] - center[d]);
                    m32 = std::max(m32, std::fabs(x - y));
                    m64 = fmax(m64, fabs(qry[(size_t)d * nq + qi] - pts[(size_t)d * n + j]));
                }
                all32[j] = {m32, (int32_t)j};
                all64[j] = {m64, perm[j]};
            }
            std::vector<std::pair<float, int32_t>> s32 = all32;
            std::sort(s32.begin(), s32.end());
            const int loose = kp - 1 + (int)(rng() % 3);
            const float s = s32[std::min<int64_t>(loose, n - 1)].first;
            const float seed = std::nextafterf(tight_seed(s, mq), CUDART_INF_F);
            std::vector<float> col_d(kcap, CUDART_INF_F);
            std::vector<int32_t> col_i(kcap, -1);
            int fill = 0;
            for (int64_t j = 0; j < n && fill < kcap; ++j)
                if (s32[j].first < seed) { col_d[fill] = s32[j].first; col_i[fill] = s32[j].second; ++fill; }
            refine_single_list(col_d.data(), col_i.data(), 1, kcap, seed, qi, qi, ra);
            if (flags[qi] != 0) { ++counted; (tied ? flagged_tied : flagged_cont)++; continue; }
            ++checked;
            std::vector<int64_t> order(n);
            for (int64_t j = 0; j < n; ++j) order[j] = j;
            std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return all64[a] < all64[b]; });
            bool ok = eps[qi] == all64[order[kp - 1]].first;
            if (ok && want_idx)
                for (int r = 0; r < kp; ++r) ok = ok && nbr[(size_t)qi * kp + r] == order[r];
            if (!ok && ++wrong < 10)
                printf("WRONG single-list trial %d qi %ld: eps %.17g want %.17g (dim %d kp %d kcap %d tied %d cross %d)\n",
                       trial, (long)qi, eps[qi], all64[order[kp - 1]].first, dim, kp, kcap, (int)tied, (int)cross);
        }
        if (counted != n_flagged) { ++wrong; printf("WRONG single-list trial %d: n_flagged %d, flags say %d\n", trial, n_flagged, counted); }
    }
    *checked_out = checked; *flagged_cont_out = flagged_cont; *flagged_tied_out = flagged_tied;
    return (int)wrong;
}

// refine_unsorted_list (csrc/fast_knn_warp.cuh): the same decision from the UNSORTED list the warp-level
// filter leaves in shared memory (entry r of a lane's column at [r * 32]).  Whenever it answers, the
// answer must be refine_single_list's on the sorted copy of the same list and the brute-force radius.
#ifdef HAVE_UNSORTED
template <int KC>
static bool call_unsorted(int kp, const float* col_d, const int32_t* col_i, int nfill, float seed, int64_t qi,
                          const RefineArgs& ra) {
    if (kp <= 4) return refine_unsorted_list<4, KC, 4>(col_d, col_i, nfill, seed, qi, qi, false, ra);
    if (kp <= 6) return refine_unsorted_list<4, KC, 6>(col_d, col_i, nfill, seed, qi, qi, false, ra);
    return refine_unsorted_list<4, KC, 8>(col_d, col_i, nfill, seed, qi, qi, false, ra);
}
static int unsorted_trials(int trials, long* checked_out, long* fallback_cont_out, long* fallback_tied_out) {
    std::mt19937_64 rng(4242);
    long checked = 0, wrong = 0, fb_cont = 0, fb_tied = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int dim = 1 + (int)(rng() % 4);
        const int kp = 1 + (int)(rng() % 8);
        const int kc = wunits_list_capacity(kp + 1 + (int)(rng() % 12));
        const int64_t n = 300 + (int64_t)(rng() % 600);
        const int64_t nq = std::min<int64_t>(n, 64 + (int64_t)(rng() % 100));
        const bool tied = rng() % 4 == 0;
        const bool cross = rng() % 5 == 0;
        const double scale = (rng() % 3 == 0) ? 1e-9 : 1.0;
        std::vector<double> pts((size_t)dim * n), qry((size_t)dim * nq), center(dim);
        std::normal_distribution<double> nd(0.0, 1.0);
        for (int d = 0; d < dim; ++d) {
            const double shift = (rng() % 3 == 0) ? 1e3 * scale : 0.0;
            center[d] = shift;
            for (int64_t i = 0; i < n; ++i) {
                double v = nd(rng);
                if (tied) v = std::round(v * 4) / 4;
                pts[(size_t)d * n + i] = v * scale + shift;
            }
            for (int64_t i = 0; i < nq; ++i) {
                double v = nd(rng);
                if (tied) v = std::round(v * 4) / 4;
                qry[(size_t)d * nq + i] = cross ? v * scale + shift : pts[(size_t)d * n + i];
            }
        }
        std::vector<int32_t> perm(n);
        for (int64_t i = 0; i < n; ++i) perm[i] = (int32_t)i;
        std::vector<double> eps(nq, -7.0), eps2(nq, -7.0);
        std::vector<int32_t> flags(nq, -7), flags2(nq, -7);
        int32_t n_flagged = 0, n_flagged2 = 0;
        std::vector<double> maxabs(dim, 1e6);
        RefineArgs ra;
        ra.fused = 1; ra.kp = kp; ra.dim = dim; ra.sorted = pts.data(); ra.ld = n;
        ra.qsorted = cross ? qry.data() : nullptr; ra.qld = nq; ra.center = center.data(); ra.perm = perm.data();
        ra.maxabs = maxabs.data(); ra.exact32 = nullptr; ra.idx_out = nullptr;
        RefineArgs rb = ra;
        ra.eps_out = eps.data(); ra.flags = flags.data(); ra.n_flagged = &n_flagged;
        rb.eps_out = eps2.data(); rb.flags = flags2.data(); rb.n_flagged = &n_flagged2;
        for (int64_t qi = 0; qi < nq; ++qi) {
            std::vector<std::pair<float, int32_t>> s32(n);
            std::vector<double> all64(n);
            float mq = 0.0f;
            for (int d = 0; d < dim; ++d) mq = std::max(mq, std::fabs((float)(qry[(size_t)d * nq + qi] - center[d])));
            for (int64_t j = 0; j < n; ++j) {
                float m32 = 0.0f;
                double m64 = 0.0;
                for (int d = 0; d < dim; ++d) {
                    const float x = (float)(qry[(size_t)d * nq + qi] - center[d]), y = (float)(pts[(size_t)d * n + j] - center[d]);
                    m32 = std::max(m32, std::fabs(x - y));
                    m64 = fmax(m64, fabs(qry[(size_t)d * nq + qi] - pts[(size_t)d * n + j]));
                }
                s32[j] = {m32, (int32_t)j};
                all64[j] = m64;
            }
            std::sort(s32.begin(), s32.end());
            const int loose = kp - 1 + (int)(rng() % 6);  // (looser seeds than the planner's: fuller lists)
            const float s = s32[std::min<int64_t>(loose, n - 1)].first;
            const float seed = std::nextafterf(tight_seed(s, mq), CUDART_INF_F);
            std::vector<std::pair<float, int32_t>> lst;
            for (int64_t j = 0; j < n && (int)lst.size() < kc; ++j)
                if (s32[j].first < seed) lst.push_back(s32[j]);
            // sorted copy, padded (what refine_single_list reads) ...
            std::vector<float> col_d(kc, CUDART_INF_F);
            std::vector<int32_t> col_i(kc, -1);
            for (size_t r = 0; r < lst.size(); ++r) { col_d[r] = lst[r].first; col_i[r] = lst[r].second; }
            refine_single_list(col_d.data(), col_i.data(), 1, kc, seed, qi, qi, rb);
            // ... and the unsorted column (stride 32; entries beyond nfill hold garbage)
            std::shuffle(lst.begin(), lst.end(), rng);
            std::vector<float> ucol_d((size_t)kc * 32, -123.0f);
            std::vector<int32_t> ucol_i((size_t)kc * 32, 7);
            for (size_t r = 0; r < lst.size(); ++r) { ucol_d[r * 32] = lst[r].first; ucol_i[r * 32] = lst[r].second; }
            bool done = false;
            switch (kc) {
                case 8: done = call_unsorted<8>(kp, ucol_d.data(), ucol_i.data(), (int)lst.size(), seed, qi, ra); break;
                case 12: done = call_unsorted<12>(kp, ucol_d.data(), ucol_i.data(), (int)lst.size(), seed, qi, ra); break;
                case 16: done = call_unsorted<16>(kp, ucol_d.data(), ucol_i.data(), (int)lst.size(), seed, qi, ra); break;
                default: done = call_unsorted<24>(kp, ucol_d.data(), ucol_i.data(), (int)lst.size(), seed, qi, ra); break;
            }
            if (!done) { (tied ? fb_tied : fb_cont)++; continue; }
            ++checked;
            std::vector<double> srt = all64;
            std::sort(srt.begin(), srt.end());
            const double want = kp <= n ? srt[kp - 1] : CUDART_INF;
            const bool ok = flags[qi] == 0 && flags2[qi] == 0 && eps[qi] == eps2[qi] && eps[qi] == want;
            if (!ok && ++wrong < 10)
                printf("WRONG unsorted trial %d qi %ld: eps %.17g sorted-path %.17g (flag %d) want %.17g (dim %d kp %d kc %d tied %d)\n",
                       trial, (long)qi, eps[qi], eps2[qi], flags2[qi], want, dim, kp, kc, (int)tied);
        }
        if (n_flagged != 0) { ++wrong; printf("WRONG unsorted trial %d: the fast path flagged %d queries itself\n", trial, n_flagged); }
    }
    *checked_out = checked; *fallback_cont_out = fb_cont; *fallback_tied_out = fb_tied;
    return (int)wrong;
}
#endif

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 400;
    std::mt19937_64 rng(2024);
    long checked = 0, wrong = 0, flagged_continuous = 0, flagged_tied = 0, decided_tied = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int dim = 1 + (int)(rng() % 8);
        const int kp = 1 + (int)(rng() % 10);
        const int kcap = kp + ((rng() % 4 == 0) ? 24 : 3);
        const int qb = 128;
        const int64_t n = 600 + (int64_t)(rng() % 900);
        const int64_t nq = std::min<int64_t>(n, 128 + (int64_t)(rng() % 200));
        const bool windowed = rng() % 3 != 0;
        const bool want_idx = rng() % 2 == 0;
        const bool tied = rng() % 4 == 0;  // quantised coordinates: exact ties, saturated lists
        std::vector<double> pts((size_t)dim * n), center(dim);
        std::normal_distribution<double> nd(0.0, 1.0);
        for (int d = 0; d < dim; ++d) {
            const double shift = (rng() % 3 == 0) ? 1e3 : 0.0;  // off-centre data: the band must use centred magnitudes
            center[d] = shift;
            for (int64_t i = 0; i < n; ++i) {
                double v = nd(rng);
                if (tied) v = std::round(v * 4) / 4;
                pts[(size_t)d * n + i] = v + shift;
            }
        }
        std::vector<int32_t> perm(n);
        for (int64_t i = 0; i < n; ++i) perm[i] = (int32_t)i;
        std::shuffle(perm.begin(), perm.end(), rng);
        const int nblocks = (int)((nq + qb - 1) / qb);
        // chunks of the candidate set per query block (windowed) or for everybody (dense)
        std::vector<int32_t> unit_first(2 * (nblocks + 1), 0);  // starts, then counts (knn_unit_table_kernel)
        const int splits = 1 + (int)(rng() % 6);
        int total_units = 0;
        for (int b = 0; b < nblocks; ++b) {
            unit_first[b] = total_units;
            total_units += 1 + (int)(rng() % 4 == 0 ? rng() % 30 : rng() % 3);
        }
        unit_first[nblocks] = total_units;
        for (int b = 0; b < nblocks; ++b) unit_first[nblocks + 1 + b] = unit_first[b + 1] - unit_first[b];
        const int n_lists = windowed ? total_units : splits;
        const int64_t stride = windowed ? qb : nq;
        std::vector<float> d32((size_t)n_lists * kcap * stride, CUDART_INF_F);
        std::vector<int32_t> idx((size_t)n_lists * kcap * stride, -1);
        auto dist64 = [&](int64_t a, int64_t b) {
            double m = 0.0;
            for (int d = 0; d < dim; ++d) m = fmax(m, fabs(pts[(size_t)d * n + a] - pts[(size_t)d * n + b]));
            return m;
        };
        auto dist32 = [&](int64_t a, int64_t b) {  // the filter's arithmetic: centred fp32 rows
            float m = 0.0f;
            for (int d = 0; d < dim; ++d) {
                const float x = (float)(pts[(size_t)d * n + a] - center[d]), y = (float)(pts[(size_t)d * n + b] - center[d]);
                m = std::max(m, std::fabs(x - y));
            }
            return m;
        };
        for (int64_t qi = 0; qi < nq; ++qi) {
            int s_begin = 0, s_end = splits;
            int64_t off = qi;
            if (windowed) { const int64_t b = qi / qb; s_begin = unit_first[b]; s_end = unit_first[b + 1]; off = qi - b * qb; }
            const int ns = s_end - s_begin;
            std::vector<std::vector<std::pair<float, int32_t>>> chunk(ns);
            for (int64_t j = 0; j < n; ++j) chunk[(size_t)((j * 2654435761ull + qi) % ns)].push_back({dist32(qi, j), (int32_t)j});
            for (int s = 0; s < ns; ++s) {
                std::sort(chunk[s].begin(), chunk[s].end());
                for (int r = 0; r < kcap && r < (int)chunk[s].size(); ++r) {
                    d32[((size_t)(s_begin + s) * kcap + r) * stride + off] = chunk[s][r].first;
                    idx[((size_t)(s_begin + s) * kcap + r) * stride + off] = chunk[s][r].second;
                }
            }
        }
        std::vector<double> eps(nq, -7.0);
        std::vector<int32_t> flags(nq, -7);
        std::vector<int64_t> nbr((size_t)nq * kp, -7);
        int32_t n_flagged = 0;
        for (int64_t qi = 0; qi < nq; ++qi) {
            blockIdx.x = (unsigned)(qi / 128);
            threadIdx.x = (unsigned)(qi % 128);
            knn_refine_kernel(d32.data(), idx.data(), kcap, splits, kp, pts.data(), n, dim, n, 0, nq, nullptr, nullptr,
                              nullptr, windowed ? unit_first.data() : nullptr, qb, perm.data(),
                              want_idx ? nbr.data() : nullptr, nullptr, 0, center.data(), eps.data(), flags.data(), &n_flagged,
                              nullptr);
        }
        int counted = 0;
        for (int64_t qi = 0; qi < nq; ++qi) {
            if (flags[qi] != 0) {
                ++counted;
                (tied ? flagged_tied : flagged_continuous)++;
                continue;
            }
            if (tied) ++decided_tied;
            ++checked;
            std::vector<std::pair<double, int32_t>> all(n);  // (exact distance, original index) of every point
            std::vector<int64_t> pos(n);
            for (int64_t j = 0; j < n; ++j) all[j] = {dist64(qi, j), perm[j]};
            std::vector<int64_t> order(n);
            for (int64_t j = 0; j < n; ++j) order[j] = j;
            std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return all[a] < all[b]; });
            bool ok = eps[qi] == all[order[kp - 1]].first;
            if (ok && want_idx)
                for (int r = 0; r < kp; ++r) ok = ok && nbr[(size_t)qi * kp + r] == order[r];
            if (!ok && ++wrong < 10)
                printf("WRONG trial %d qi %ld: eps %.17g want %.17g (dim %d kp %d kcap %d windowed %d tied %d)\n", trial,
                       (long)qi, eps[qi], all[order[kp - 1]].first, dim, kp, kcap, (int)windowed, (int)tied);
        }
        if (counted != n_flagged) { ++wrong; printf("WRONG trial %d: n_flagged %d, flags say %d\n", trial, n_flagged, counted); }
    }
    long s_checked = 0, s_flagged_cont = 0, s_flagged_tied = 0;
    wrong += single_list_trials(trials, &s_checked, &s_flagged_cont, &s_flagged_tied);
    long u_checked = 0, u_fb_cont = 0, u_fb_tied = 0;
#ifdef HAVE_UNSORTED
    wrong += unsorted_trials(trials, &u_checked, &u_fb_cont, &u_fb_tied);
#endif
    printf("checked %ld wrong %ld flagged_continuous %ld flagged_tied %ld decided_tied %ld single_checked %ld "
           "single_flagged_continuous %ld single_flagged_tied %ld unsorted_checked %ld unsorted_fallback_continuous %ld "
           "unsorted_fallback_tied %ld\n", checked, wrong,
           flagged_continuous, flagged_tied, decided_tied, s_checked, s_flagged_cont, s_flagged_tied, u_checked, u_fb_cont,
           u_fb_tied);
    return wrong ? 1 : 0;
}

// KSG-1 mutual information of a 2-D sample through the C ABI of include/ksg_b200.h only:
// CUDA runtime for device memory, no PyTorch, no Python.  This is the call sequence a
// maintainer's binding (cgo / JNI / ctypes ...) issues for syntropy.knn.mutual_information
// (reference: syntropy/knn/shannon.py:137-196), see INTEGRATION.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o ksg_mi_cabi examples/ksg_mi_cabi.cu \
//        -L syntropy_b200/lib -lksg_b200 -Xlinker -rpath -Xlinker $PWD/syntropy_b200/lib
//   ./ksg_mi_cabi [n] [k]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <vector>

#include "ksg_b200.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != 0) {                                                                    \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ksg_last_error());         \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)
#define CUDA(call)                                                                         \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                    \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

template <class T>
static T* dev_alloc(size_t count) {
    void* p = nullptr;
    if (cudaMalloc(&p, count * sizeof(T) + 256) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(1); }
    return (T*)p;
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000;
    const int k = argc > 2 ? atoi(argv[2]) : 5;
    const int dim = 2, kprime = k + 1;

    // x ~ N(0,1), y = x^2 + 0.5 noise (the README example of the reference), rows = variables
    std::vector<double> host((size_t)dim * n);
    std::mt19937_64 gen(0);
    std::normal_distribution<double> normal(0.0, 1.0);
    for (int64_t i = 0; i < n; ++i) {
        const double x = normal(gen);
        host[i] = x;
        host[n + i] = x * x + 0.5 * normal(gen);
    }
    cudaStream_t st;
    CUDA(cudaStreamCreate(&st));
    double* d_pts = dev_alloc<double>((size_t)dim * n);
    CUDA(cudaMemcpyAsync(d_pts, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));

    // every cKDTree of the reference at once: the prepared set (curve order)
    const size_t prep_bytes = ksg_prepare_workspace(n, dim);
    void* d_prep = dev_alloc<char>(prep_bytes);
    ksg_prepared P;
    CHECK(ksg_prepare(d_pts, n, n, dim, -1, d_prep, prep_bytes, &P, st));

    // joint radius: tree.query(data.T, k + 1, p = inf)[:, -1]
    double* d_eps = dev_alloc<double>(n);
    int32_t* d_flags = dev_alloc<int32_t>(n);
    int32_t* d_nflag = dev_alloc<int32_t>(1);
    CUDA(cudaMemsetAsync(d_nflag, 0, sizeof(int32_t), st));
    const size_t ws_bytes = ksg_fast_knn_workspace(n, n, dim, kprime);
    void* d_ws = dev_alloc<char>(ws_bytes);
    CHECK(ksg_fast_knn(&P, 0, n, nullptr, kprime, /*windowed=*/1, d_eps, d_flags, d_nflag, d_ws, ws_bytes, st));
    int32_t nflag = 0;
    CUDA(cudaMemcpyAsync(&nflag, d_nflag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA(cudaStreamSynchronize(st));
    if (nflag) {
        // (tie-saturated or set-aside queries: rerun them with d_query_index / ksg_knn_radius, see
        //  syntropy_b200/_engine.py:fast_knn; continuous data never gets here)
        fprintf(stderr, "%d queries need the fallback path; this example does not implement it\n", nflag);
        return 2;
    }

    // marginal counts at the joint radius: get_counts_from_tree for x and for y (strict, minus self)
    int32_t* d_counts = dev_alloc<int32_t>(2 * (size_t)n);
    for (int axis = 0; axis < dim; ++axis)
        CHECK(ksg_axis_count(&P, axis, 0, n, d_eps, /*strict=*/1, /*bias=*/-1, d_counts + (size_t)axis * n, st));

    // ptw = (psi(k) + psi(N)) - psi(n_x + 1) - psi(n_y + 1), in the reference's order; mean
    double* d_psi = dev_alloc<double>(n + 2);
    CHECK(ksg_psi_table(d_psi, n + 2, st));
    double psi_kn[2];
    CUDA(cudaMemcpyAsync(&psi_kn[0], d_psi + k, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA(cudaMemcpyAsync(&psi_kn[1], d_psi + n, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA(cudaStreamSynchronize(st));
    ksg_psi_step steps[2] = {{0, 1, -1, 0, 1.0}, {1, 1, -1, 0, 1.0}};
    double* d_ptw_sorted = dev_alloc<double>(n);
    double* d_ptw = dev_alloc<double>(n);
    CHECK(ksg_psi_epilogue(d_counts, n, 2, steps, 2, psi_kn[0] + psi_kn[1], 0.0, 0, 0.0, d_psi, n + 2, d_ptw_sorted, st));
    CHECK(ksg_scatter_f64(d_ptw_sorted, P.perm, n, d_ptw, st));  // back to the caller's sample order
    double* d_sum = dev_alloc<double>(1);
    const size_t sum_bytes = ksg_sum_workspace(n);
    void* d_sum_ws = dev_alloc<char>(sum_bytes);
    CHECK(ksg_sum_f64(d_ptw, n, d_sum, d_sum_ws, sum_bytes, st));
    double total = 0.0;
    CUDA(cudaMemcpyAsync(&total, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA(cudaStreamSynchronize(st));
    printf("KSG-1 MI(x; x^2 + noise), n = %lld, k = %d: %.15f nats  (%llu kernel launches)\n", (long long)n, k,
           total / (double)n, (unsigned long long)ksg_launch_count());
    return 0;
}

// Thin PyTorch C++ extension over the C ABI of include/ksg_b200.h (BASELINE.json: "Host code stays in
// Python and calls the kernels through a thin PyTorch C++/CUDA extension over a C ABI").
//
// torch supplies what the C ABI deliberately does not own: device memory (at::empty through the
// caching allocator) and the current stream.  Everything else is a call into libksg_b200.so.
// Two kinds of entry points:
//   * ksg1_step / finish_step -- one host call per estimator step for the headline family (joint
//     radius -> 1-D marginal counts -> psi program, run speculatively, see _engine.ksg1_device): the
//     ~12 Python + ctypes round trips and ~12 tensor allocations of that step collapse into one
//     pybind call (host time per step ~120 us -> ~25 us);
//   * a CUDA-event span recorder so that bench.py still gets per-stage device times (prepare, fast_knn,
//     axis_count, psi_epilogue, sum_f64) out of the fused call.
// The ctypes view (syntropy_b200/_cabi.py) stays the general binding for every other entry point.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <torch/extension.h>

#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ksg_b200.h"

namespace {

void check(int rc, const char* what) {
    TORCH_CHECK(rc == KSG_OK, what, " failed with code ", rc, ": ", ksg_last_error());
}

void* stream_of(const at::Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

at::Tensor bytes(size_t n, const at::Tensor& like) {
    return at::empty({(int64_t)(n > 0 ? n : 1)}, like.options().dtype(at::kByte));
}

// ---- per-stage device timing (CUDA events on the launching stream), switched on by bench.py
struct Span {
    std::string name;
    cudaEvent_t a, b;
};
bool g_timing = false;
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t take_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    TORCH_CHECK(cudaEventCreate(&e) == cudaSuccess, "cudaEventCreate failed");
    return e;
}

// (every stage is also an NVTX range -- prepare / fast_knn / axis_count / psi_* / p2p_wait / scatter_sum -- so a
//  timeline tool shows the step's structure; a push / pop without a tool attached costs nanoseconds)
struct Scope {
    bool on;
    size_t at = 0;
    cudaStream_t st;
    Scope(const char* name, void* stream) : on(g_timing), st((cudaStream_t)stream) {
        nvtxRangePushA(name);
        if (!on) return;
        g_spans.push_back(Span{name, take_event(), take_event()});
        at = g_spans.size() - 1;
        cudaEventRecord(g_spans[at].a, st);
    }
    ~Scope() {
        if (on) cudaEventRecord(g_spans[at].b, st);
        nvtxRangePop();
    }
};

void set_timing(bool on) { g_timing = on; }

// name -> list of milliseconds; synchronises the device, empties the recorder
std::map<std::string, std::vector<double>> collect_spans() {
    cudaDeviceSynchronize();
    std::map<std::string, std::vector<double>> out;
    for (auto& s : g_spans) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) out[s.name].push_back((double)ms);
        g_pool.push_back(s.a);
        g_pool.push_back(s.b);
    }
    g_spans.clear();
    return out;
}

// The speculative KSG-1 step for 1-D marginals (mutual information of two scalar series, total
// correlation): prepared set -> windowed / dense fp32-filtered k-NN radii of sorted positions
// [q0, q1) (unsettled queries counted into `nflag`, not read) -> one exact binary-search count per
// marginal -> psi program.  Returns {ptw (q1 - q0, prepared order), prepared buffer, prepared
// struct (CPU bytes), eps, flags, counts}: what _engine needs for Pointwise / its redo closure.
std::vector<at::Tensor> ksg1_step(at::Tensor pts, int64_t k, std::vector<int64_t> axes, double init, double psi_n,
                                  bool apply_scale, double scale, std::vector<int64_t> step_sub,
                                  std::vector<int64_t> step_off, std::vector<int64_t> step_sign,
                                  std::vector<int64_t> step_centered, std::vector<double> step_div,
                                  at::Tensor psi_table, int64_t q0, int64_t q1, bool windowed, int64_t primary,
                                  int64_t margin, at::Tensor nflag, std::vector<int64_t> row_ready,
                                  c10::optional<at::Tensor> total, c10::optional<at::Tensor> ticket,
                                  std::vector<int64_t> peer_values, std::vector<int64_t> peer_aux,
                                  std::vector<int64_t> peer_flags, int64_t rank, int64_t epoch, int64_t p2p_ticket) {
    TORCH_CHECK(pts.is_cuda() && pts.scalar_type() == at::kDouble && pts.dim() == 2 && pts.stride(1) == 1,
                "pts must be a (D, n) float64 CUDA tensor with unit column stride");
    c10::cuda::CUDAGuard guard(pts.device());
    const int64_t dim = pts.size(0), n = pts.size(1), nq = q1 - q0;
    const int kprime = (int)k + 1;
    void* st = stream_of(pts);
    auto opts = pts.options();

    // ---- prepared set
    const size_t prep_bytes = ksg_prepare_workspace(n, (int)dim);
    at::Tensor prep_buf = bytes(prep_bytes, pts);
    ksg_prepared P;
    std::vector<void*> ready(dim, nullptr);
    const bool have_ready = (int64_t)row_ready.size() == dim;
    for (int64_t d = 0; have_ready && d < dim; ++d) ready[d] = (void*)(intptr_t)row_ready[d];
    {
        Scope s("prepare", st);
        check(ksg_prepare_rows(pts.data_ptr<double>(), pts.stride(0), n, (int)dim, (int)primary,
                               have_ready ? ready.data() : nullptr, prep_buf.data_ptr(), prep_bytes, &P, st),
              "ksg_prepare_rows");
    }
    at::Tensor prep_struct = at::empty({(int64_t)sizeof(ksg_prepared)}, at::TensorOptions().dtype(at::kByte));
    std::memcpy(prep_struct.data_ptr(), &P, sizeof(P));

    // ---- joint radii (speculative: flags / nflag are left on the device)
    at::Tensor eps = at::empty({nq}, opts);
    at::Tensor flags = at::empty({nq}, opts.dtype(at::kInt));
    at::Tensor counts = at::empty({(int64_t)axes.size(), nq}, opts.dtype(at::kInt));
    at::Tensor ptw = at::empty({nq}, opts);
    if (nq > 0) {
        const size_t ws_bytes = ksg_fast_knn_workspace_margin(n, nq, (int)dim, kprime, (int)margin);
        at::Tensor ws = bytes(ws_bytes, pts);
        {
            Scope s("fast_knn", st);
            check(ksg_fast_knn_margin(&P, q0, q1, nullptr, kprime, (int)margin, windowed ? 1 : 0, eps.data_ptr<double>(),
                                      nullptr, flags.data_ptr<int32_t>(), nflag.data_ptr<int32_t>(), ws.data_ptr(),
                                      ws_bytes, st),
                  "ksg_fast_knn_margin");
        }
        // ---- the exact 1-D counts of all marginals, one launch (lazy axes: from the value buckets; what they
        //      cannot settle is counted into nflag like the k-NN stage's unsettled queries)
        {
            std::vector<int32_t> ax(axes.begin(), axes.end());
            Scope s("axis_count", st);
            if (P.axes_mode > 0)
                check(ksg_axis_count_buckets(&P, ax.data(), (int)ax.size(), q0, q1, eps.data_ptr<double>(), 1, -1,
                                             counts.data_ptr<int32_t>(), nflag.data_ptr<int32_t>(), st),
                      "ksg_axis_count_buckets");
            else
                check(ksg_axis_count_multi(&P, ax.data(), (int)ax.size(), q0, q1, eps.data_ptr<double>(), 1, -1,
                                           counts.data_ptr<int32_t>(), st),
                      "ksg_axis_count_multi");
        }
        // ---- psi program
        std::vector<ksg_psi_step> steps(step_sub.size() ? step_sub.size() : 1);
        for (size_t t = 0; t < step_sub.size(); ++t) {
            steps[t].subspace = (int32_t)step_sub[t];
            steps[t].count_offset = (int32_t)step_off[t];
            steps[t].sign = (int32_t)step_sign[t];
            steps[t].centered = (int32_t)step_centered[t];
            steps[t].divisor = step_div[t];
        }
        if (total.has_value() && ticket.has_value() && q0 == 0 && q1 == n) {
            // one rank owns every query: psi program + sample order + deterministic sum in ONE launch
            // (ptw then holds all n values in the CALLER's order and *total their sum)
            const size_t fin_bytes = ksg_finish_workspace(n);
            at::Tensor fin_ws = bytes(fin_bytes, pts);
            Scope s("psi_finish", st);
            check(ksg_psi_finish(counts.data_ptr<int32_t>(), n, (int)axes.size(), steps.data(), (int)step_sub.size(), init,
                                 psi_n, apply_scale ? 1 : 0, scale, psi_table.data_ptr<double>(), psi_table.numel(), P.perm,
                                 ptw.data_ptr<double>(), total->data_ptr<double>(), ticket->data_ptr<int32_t>(),
                                 fin_ws.data_ptr(), fin_bytes, st),
                  "ksg_psi_finish");
            return {ptw, prep_buf, prep_struct, eps, flags, counts};
        }
        if (!peer_values.empty()) {
            // several ranks: the psi kernel stores this rank's slice straight into EVERY rank's vector over
            // NVLink and publishes the unsettled-query counter with it (no library collective in the step)
            const int world = (int)peer_values.size();
            std::vector<void*> pv(world), pa(world);
            std::vector<unsigned long long*> pf(world);
            for (int r = 0; r < world; ++r) {
                pv[r] = (void*)(intptr_t)peer_values[r];
                pa[r] = (void*)(intptr_t)peer_aux[r];
                pf[r] = (unsigned long long*)(intptr_t)peer_flags[r];
            }
            Scope s("psi_push", st);
            check(ksg_psi_epilogue_push(counts.data_ptr<int32_t>(), nq, (int)axes.size(), steps.data(), (int)step_sub.size(),
                                        init, psi_n, apply_scale ? 1 : 0, scale, psi_table.data_ptr<double>(),
                                        psi_table.numel(), pv.data(), q0, nflag.data_ptr<int32_t>(), (int32_t* const*)pa.data(),
                                        pf.data(), (int)rank, world, (unsigned long long)epoch,
                                        (int32_t*)(intptr_t)p2p_ticket, st),
                  "ksg_psi_epilogue_push");
            return {ptw, prep_buf, prep_struct, eps, flags, counts};
        }
        Scope s("psi_epilogue", st);
        check(ksg_psi_epilogue(counts.data_ptr<int32_t>(), nq, (int)axes.size(), steps.data(), (int)step_sub.size(), init,
                               psi_n, apply_scale ? 1 : 0, scale, psi_table.data_ptr<double>(), psi_table.numel(),
                               ptw.data_ptr<double>(), st),
              "ksg_psi_epilogue");
    } else if (!peer_values.empty()) {
        // an empty slice still has to publish its flags (the peers wait for every rank)
        const int world = (int)peer_values.size();
        std::vector<void*> pv(world), pa(world);
        std::vector<unsigned long long*> pf(world);
        for (int r = 0; r < world; ++r) {
            pv[r] = (void*)(intptr_t)peer_values[r];
            pa[r] = (void*)(intptr_t)peer_aux[r];
            pf[r] = (unsigned long long*)(intptr_t)peer_flags[r];
        }
        ksg_psi_step none{};
        check(ksg_psi_epilogue_push(nullptr, 0, (int)axes.size(), &none, 0, init, psi_n, 0, scale,
                                    psi_table.data_ptr<double>(), psi_table.numel(), pv.data(), q0,
                                    nflag.data_ptr<int32_t>(), (int32_t* const*)pa.data(), pf.data(), (int)rank, world,
                                    (unsigned long long)epoch, (int32_t*)(intptr_t)p2p_ticket, st),
              "ksg_psi_epilogue_push");
    }
    return {ptw, prep_buf, prep_struct, eps, flags, counts};
}

// full pointwise vector (prepared order) -> caller's sample order + fixed-order sum into `total`
// (wait_flags != 0: values_sorted is this rank's peer-written vector -- first acquire the `world` flags of
//  exchange `epoch` and fold the ranks' unsettled-query counters into *nflag, see ksg_p2p_wait)
at::Tensor finish_step(at::Tensor values_sorted, at::Tensor prep_struct, at::Tensor total,
                       c10::optional<at::Tensor> ticket, int64_t wait_flags, int64_t world, int64_t epoch,
                       int64_t wait_aux, c10::optional<at::Tensor> nflag) {
    c10::cuda::CUDAGuard guard(values_sorted.device());
    ksg_prepared P;
    std::memcpy(&P, prep_struct.data_ptr(), sizeof(P));
    const int64_t n = values_sorted.numel();
    void* st = stream_of(values_sorted);
    if (wait_flags != 0) {
        Scope s("p2p_wait", st);
        check(ksg_p2p_wait((const unsigned long long*)(intptr_t)wait_flags, (int)world, (unsigned long long)epoch,
                           (const int32_t*)(intptr_t)wait_aux, nflag.has_value() ? nflag->data_ptr<int32_t>() : nullptr, st),
              "ksg_p2p_wait");
    }
    at::Tensor out = at::empty_like(values_sorted);
    if (ticket.has_value()) {  // scatter + both stages of the sum in one launch
        const size_t fin_bytes = ksg_finish_workspace(n);
        at::Tensor fin_ws = bytes(fin_bytes, values_sorted);
        Scope s("scatter_sum", st);
        check(ksg_scatter_sum_f64(values_sorted.data_ptr<double>(), P.perm, n, out.data_ptr<double>(),
                                  total.data_ptr<double>(), ticket->data_ptr<int32_t>(), fin_ws.data_ptr(), fin_bytes, st),
              "ksg_scatter_sum_f64");
        return out;
    }
    check(ksg_scatter_f64(values_sorted.data_ptr<double>(), P.perm, n, out.data_ptr<double>(), st), "ksg_scatter_f64");
    const size_t ws_bytes = ksg_sum_workspace(n);
    at::Tensor ws = bytes(ws_bytes, values_sorted);
    Scope s("sum_f64", st);
    check(ksg_sum_f64(out.data_ptr<double>(), n, total.data_ptr<double>(), ws.data_ptr(), ws_bytes, st), "ksg_sum_f64");
    return out;
}

int64_t abi_version() { return ksg_version(); }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "thin torch binding of libksg_b200.so (include/ksg_b200.h)";
    m.def("abi_version", &abi_version);
    m.def("ksg1_step", &ksg1_step);
    m.def("finish_step", &finish_step);
    m.def("set_timing", &set_timing);
    m.def("collect_spans", &collect_spans);
}

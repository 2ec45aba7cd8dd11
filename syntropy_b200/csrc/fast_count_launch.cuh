#pragma once
#include "fast_count.cuh"
#include "fast_launch.h"

namespace ksg {

template <int DP>
struct FastGeom {
    static constexpr int Q = DP <= 4 ? 4 : 2;
    static constexpr int STAGES = DP <= 4 ? 3 : 2;
};

template <class Spec, bool MASKED>
static int launch_count_spec(const CountF32Args& a) {
    constexpr int Q = FastGeom<Spec::DP>::Q, ST = FastGeom<Spec::DP>::STAGES;
    auto kern = count_f32_kernel<Spec, Q, ST, MASKED>;
    const size_t smem = (size_t)ST * FAST_TILE * Spec::DP * sizeof(float);
    if (smem > 32 * 1024) {  // static shared (mbarriers) counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    }
    MaskQuad mq;
    for (int i = 0; i < 4; ++i) mq.m[i] = a.masks[i];
    dim3 grid((unsigned)ceil_div(a.nq, FAST_THREADS * Q), (unsigned)a.splits, 1);
    kern<<<grid, FAST_THREADS, smem, a.st>>>(a.shadow, a.tile_box, a.n_tiles, a.q_begin, a.nq, a.thresholds,
                                              a.windowed, mq, a.n_sub, a.counts, a.tile_counter);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("count_f32_kernel launch: %s", cudaGetErrorString(e)); return KSG_ECUDA; }
    return KSG_OK;
}

}  // namespace ksg

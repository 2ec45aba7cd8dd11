// Shared helpers for libksg_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ksg_b200.h"

namespace ksg {

void set_error(const char* fmt, ...);

#define KSG_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ksg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                           __FILE__, __LINE__);                                         \
            return KSG_ECUDA;                                                           \
        }                                                                               \
    } while (0)

#define KSG_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ksg::set_error("invalid argument: %s", msg);         \
            return KSG_EINVAL;                                   \
        }                                                        \
    } while (0)

// Number of SMs of the current device (148 on B200); 148 if no device is visible so
// that workspace sizing stays callable on a CPU-only host.
int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// How many candidate ranges a query block is split into.  The dense kernels keep ~4 CTAs resident
// per SM; the grid should be several waves of that (here >= 6), or the last, partly filled wave
// costs a large share of the time when few queries are resident on this GPU (8-GPU sharding:
// 244 query blocks x 3 splits was 1.24 waves) -- without making a split shorter than two tiles.
static inline int choose_splits(int64_t n_query_blocks, int64_t n_candidates, int tile) {
    int64_t want = ceil_div(24 * (int64_t)sm_count(), n_query_blocks > 0 ? n_query_blocks : 1);
    int64_t max_by_len = n_candidates / (2 * (int64_t)tile);
    if (max_by_len < 1) max_by_len = 1;
    if (want > max_by_len) want = max_by_len;
    if (want > 128) want = 128;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace ksg

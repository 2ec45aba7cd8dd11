// Step-0 microbenchmark: issue rates of the instructions the KSG pair
// loops are made of, measured on the B200 itself so that the FP-pipe roofline
// denominator is a measurement, not an assumption.
//
// One CTA of 1024 threads per SM; every thread runs ITER iterations of 16 independent
// chains of the op under test; cycles come from clock64() inside the kernel, so the
// result is lane-ops per clock per SM independent of the DVFS state.
//
// Build/run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu && ./pipe_microbench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

constexpr int THREADS = 1024;
constexpr int ITER = 4096;
constexpr int CH = 16;

#define KERNEL_BEGIN(name)                                                         \
    __global__ void __launch_bounds__(THREADS) name(float* out, long long* cyc, float seed) { \
        float a[CH];                                                               \
        float b = seed * 1.0001f + threadIdx.x * 1e-7f;                            \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) a[i] = seed + i + threadIdx.x; \
        int cnt[CH];                                                               \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) cnt[i] = 0;                 \
        __syncthreads();                                                           \
        long long t0 = clock64();                                                  \
        for (int it = 0; it < ITER; ++it) {                                        \
            _Pragma("unroll") for (int i = 0; i < CH; ++i) {

#define KERNEL_END                                                                 \
            }                                                                      \
        }                                                                          \
        long long t1 = clock64();                                                  \
        float s = 0.f; int c = 0;                                                  \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) { s += a[i]; c += cnt[i]; } \
        out[blockIdx.x * THREADS + threadIdx.x] = s + c;                           \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                           \
    }

KERNEL_BEGIN(k_fadd)
asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END

KERNEL_BEGIN(k_ffma)
asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END

KERNEL_BEGIN(k_fmnmx)
asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END

// 3-input max (FMNMX3)
KERNEL_BEGIN(k_fmnmx3)
asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(a[(i + 1) % CH]));
KERNEL_END

// 2-input max with |.| on both operands
KERNEL_BEGIN(k_fmnmx_abs)
asm volatile("{ .reg .f32 x, y; abs.f32 x, %0; abs.f32 y, %1; max.f32 %0, x, y; }" : "+f"(a[i]) : "f"(b));
KERNEL_END

// FSET alone (1.0 / 0.0)
KERNEL_BEGIN(k_fset)
asm volatile("set.lt.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END

// funnel shift
KERNEL_BEGIN(k_shf)
asm volatile("shf.l.clamp.b32 %0, %1, %0, 1;" : "+r"(cnt[i]) : "r"(__float_as_int(b)));
a[i] += 0.f;
KERNEL_END

// FADD (FMA pipe) + FMNMX (ALU pipe) interleaved: do they overlap?
KERNEL_BEGIN(k_fadd_fmnmx)
asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
asm volatile("max.f32 %0, %0, %1;" : "+f"(a[(i + 8) % CH]) : "f"(b));
KERNEL_END

// FMNMX + FSET (both ALU?)
KERNEL_BEGIN(k_fmnmx_fset)
asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
asm volatile("set.lt.f32.f32 %0, %0, %1;" : "+f"(a[(i + 8) % CH]) : "f"(b));
KERNEL_END

// integer add (which pipe?) next to FMNMX
KERNEL_BEGIN(k_iadd_fmnmx)
asm volatile("add.s32 %0, %0, %1;" : "+r"(cnt[i]) : "r"(__float_as_int(b)));
asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END

// |a - b| then max: expect FADD + FMNMX(|.|)
KERNEL_BEGIN(k_sub_absmax)
asm volatile("{ .reg .f32 t; sub.f32 t, %0, %1; abs.f32 t, t; max.f32 %0, %0, t; }" : "+f"(a[i]) : "f"(b));
KERNEL_END

// compare + predicated integer increment (FSETP + IADD)
KERNEL_BEGIN(k_fsetp_iadd)
asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt[i]) : "f"(a[i]), "f"(b));
a[i] += 0.f;
KERNEL_END

// compare to 1.0/0.0 (FSET) + float accumulate (FADD)
KERNEL_BEGIN(k_fset_fadd)
asm volatile("{ .reg .f32 t; set.lt.f32.f32 t, %1, %2; add.f32 %0, %0, t; }" : "+f"(a[i]) : "f"(a[(i + 1) % CH]), "f"(b));
KERNEL_END

// FSETP alone feeding a predicate chain (and-combined), one IADD per 4 compares
KERNEL_BEGIN(k_fsetp_chain)
if ((i & 3) == 3) {
    asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %5; setp.lt.and.f32 p, %2, %5, p; setp.lt.and.f32 p, %3, %5, p; "
                 "setp.lt.and.f32 p, %4, %5, p; @p add.s32 %0, %0, 1; }"
                 : "+r"(cnt[i]) : "f"(a[i]), "f"(a[i - 1]), "f"(a[i - 2]), "f"(a[i - 3]), "f"(b));
}
KERNEL_END

// saturating multiply as an indicator on the FMA pipe: sat((t - |d|) * BIG), then FADD
KERNEL_BEGIN(k_fmulsat_fadd)
asm volatile("{ .reg .f32 t; sub.f32 t, %2, %1; mul.sat.f32 t, t, 0f7E800000; add.f32 %0, %0, t; }"
             : "+f"(a[i]) : "f"(a[(i + 1) % CH]), "f"(b));
KERNEL_END

// packed fp32x2 add (Blackwell FADD2): two lanes of work per instruction
__global__ void __launch_bounds__(THREADS) k_fadd2(float* out, long long* cyc, float seed) {
    unsigned long long a[CH];
    unsigned long long b;
    float bl = seed * 1.0001f + threadIdx.x * 1e-7f;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bl));
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float v = seed + i + threadIdx.x;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(v));
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
        s += lo + hi;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

#define DKERNEL_BEGIN(name)                                                        \
    __global__ void __launch_bounds__(THREADS) name(float* out, long long* cyc, float seed) { \
        double a[CH];                                                              \
        double b = seed * 1.0001 + threadIdx.x * 1e-7;                             \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) a[i] = seed + i + threadIdx.x; \
        int cnt[CH];                                                               \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) cnt[i] = 0;                 \
        __syncthreads();                                                           \
        long long t0 = clock64();                                                  \
        for (int it = 0; it < ITER / 4; ++it) {                                    \
            _Pragma("unroll") for (int i = 0; i < CH; ++i) {

#define DKERNEL_END                                                                \
            }                                                                      \
        }                                                                          \
        long long t1 = clock64();                                                  \
        double s = 0.; int c = 0;                                                  \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) { s += a[i]; c += cnt[i]; } \
        out[blockIdx.x * THREADS + threadIdx.x] = (float)s + c;                    \
        if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) * 4;                     \
    }

DKERNEL_BEGIN(k_dadd)
asm volatile("add.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
DKERNEL_END

DKERNEL_BEGIN(k_dmax)
asm volatile("max.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
DKERNEL_END

DKERNEL_BEGIN(k_dsetp_iadd)
asm volatile("{ .reg .pred p; setp.lt.f64 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt[i]) : "d"(a[i]), "d"(b));
DKERNEL_END

DKERNEL_BEGIN(k_dsub_abs_setp)
asm volatile("{ .reg .pred p; .reg .f64 t; sub.f64 t, %1, %2; abs.f64 t, t; setp.lt.f64 p, t, %2; @p add.s32 %0, %0, 1; }"
             : "+r"(cnt[i]) : "d"(a[i]), "d"(b));
DKERNEL_END

// shared-memory broadcast loads (all lanes same address), 128-bit
__global__ void __launch_bounds__(THREADS) k_lds128_bcast(float* out, long long* cyc, float seed) {
    __shared__ float4 sm[1024];
    sm[threadIdx.x] = make_float4(seed, seed + 1, seed + 2, seed + 3);
    __syncthreads();
    float s = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            float4 v;
            unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(it + i) & 1023]);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            s += v.x;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(float*, long long*, float);

static void run(const char* name, kern_t k, double lane_ops_per_chain_step, int sms, float* out, long long* cyc) {
    k<<<sms, THREADS>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    k<<<sms, THREADS>>>(out, cyc, 1.5f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-18s failed: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double med = (double)h[sms / 2];
    double steps = (double)THREADS * ITER * CH;  // chain steps per SM
    printf("%-18s %9.0f cycles  %7.2f chain-steps/clk/SM  %7.2f lane-ops/clk/SM\n", name, med, steps / med,
           steps * lane_ops_per_chain_step / med);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("SMs %d, max SM clock %d kHz\n", sms, khz);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * THREADS);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    run("fadd", k_fadd, 1, sms, out, cyc);
    run("fmnmx3", k_fmnmx3, 1, sms, out, cyc);
    run("fmnmx |a|,|b|", k_fmnmx_abs, 1, sms, out, cyc);
    run("fset", k_fset, 1, sms, out, cyc);
    run("shf", k_shf, 1, sms, out, cyc);
    run("fadd+fmnmx", k_fadd_fmnmx, 2, sms, out, cyc);
    run("fmnmx+fset", k_fmnmx_fset, 2, sms, out, cyc);
    run("iadd+fmnmx", k_iadd_fmnmx, 2, sms, out, cyc);
    run("ffma", k_ffma, 1, sms, out, cyc);
    run("fmnmx", k_fmnmx, 1, sms, out, cyc);
    run("sub+|max|", k_sub_absmax, 2, sms, out, cyc);
    run("fsetp+@iadd", k_fsetp_iadd, 2, sms, out, cyc);
    run("fset+fadd", k_fset_fadd, 2, sms, out, cyc);
    run("fsetp-chain4", k_fsetp_chain, 5.0 / 4.0, sms, out, cyc);
    run("sub+mulsat+fadd", k_fmulsat_fadd, 3, sms, out, cyc);
    run("fadd2 (x2)", k_fadd2, 2, sms, out, cyc);
    run("dadd", k_dadd, 1, sms, out, cyc);
    run("dmax", k_dmax, 1, sms, out, cyc);
    run("dsetp+@iadd", k_dsetp_iadd, 2, sms, out, cyc);
    run("dsub+|dsetp|+iadd", k_dsub_abs_setp, 3, sms, out, cyc);
    run("lds128 bcast", k_lds128_bcast, 1, sms, out, cyc);
    return 0;
}

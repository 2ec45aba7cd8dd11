// Round-2 microbenchmark, two parts.
//
// (1) Issue rates of the instructions the leave-one-out count loop could be made of (saturating
//     FMA / multiply / add, three-register FFMA, packed FFMA2, IMAD, IMAD.HI, shift + add, LOP3):
//     same harness as pipe_microbench.cu (one 1024-thread CTA per SM, 16 independent chains,
//     clock64 inside the kernel -> lane-ops per clock per SM).
// (2) The inner loop of count_f32_kernel<SpecLoo<8>> in isolation: 128 threads, Q = 2 queries per
//     thread in registers, a 512-row 8-D tile in shared memory walked REPS times, all eight
//     leave-one-out counts per query -- in several arithmetic formulations that give identical
//     counts (checked against each other), timed with CUDA events -> pairs per second per GPU.
//
// Build/run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipe_microbench2 pipe_microbench2.cu && ./pipe_microbench2
#include <cuda_runtime.h>
#include <math.h>
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
        float c = seed * 0.37f + threadIdx.x * 3e-7f;                              \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) a[i] = seed + i + threadIdx.x; \
        int cnt[CH];                                                               \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) cnt[i] = i + threadIdx.x;   \
        __syncthreads();                                                           \
        long long t0 = clock64();                                                  \
        for (int it = 0; it < ITER; ++it) {                                        \
            _Pragma("unroll") for (int i = 0; i < CH; ++i) {

#define KERNEL_END                                                                 \
            }                                                                      \
        }                                                                          \
        long long t1 = clock64();                                                  \
        float s = c; int n = 0;                                                    \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) { s += a[i]; n += cnt[i]; } \
        out[blockIdx.x * THREADS + threadIdx.x] = s + n;                           \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                           \
    }

KERNEL_BEGIN(k_fadd)
asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END
KERNEL_BEGIN(k_ffma3)
asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
KERNEL_END
KERNEL_BEGIN(k_ffma3_sat)
asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
KERNEL_END
KERNEL_BEGIN(k_ffma3_sat_abs)
asm volatile("{ .reg .f32 t; abs.f32 t, %0; fma.rn.sat.f32 %0, t, %1, %2; }" : "+f"(a[i]) : "f"(b), "f"(c));
KERNEL_END
KERNEL_BEGIN(k_ffma_imm)
asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END
KERNEL_BEGIN(k_ffma_imm_sat)
asm volatile("fma.rn.sat.f32 %0, %0, 0f3F800001, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END
KERNEL_BEGIN(k_fmul_imm_sat)
asm volatile("mul.sat.f32 %0, %0, 0f7E800000;" : "+f"(a[i]));
KERNEL_END
KERNEL_BEGIN(k_fadd_sat)
asm volatile("add.sat.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END
KERNEL_BEGIN(k_fadd_abs)
asm volatile("{ .reg .f32 t; abs.f32 t, %0; sub.f32 %0, t, %1; }" : "+f"(a[i]) : "f"(b));
KERNEL_END
KERNEL_BEGIN(k_imad)
asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(cnt[i]) : "r"(__float_as_int(b)));
KERNEL_END
KERNEL_BEGIN(k_imad_hi)
asm volatile("mad.hi.u32 %0, %1, 2, %0;" : "+r"(cnt[i]) : "r"(__float_as_int(a[i])));
a[i] += 0.f;
KERNEL_END
KERNEL_BEGIN(k_shr_add)
asm volatile("{ .reg .b32 t; shr.u32 t, %1, 31; add.s32 %0, %0, t; }" : "+r"(cnt[i]) : "r"(__float_as_int(a[i])));
a[i] += 0.f;
KERNEL_END
KERNEL_BEGIN(k_lop3)
asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(cnt[i]) : "r"(__float_as_int(b)), "r"(__float_as_int(c)));
KERNEL_END
KERNEL_BEGIN(k_iadd3)
asm volatile("{ .reg .b32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(cnt[i]) : "r"(__float_as_int(b)), "r"(__float_as_int(c)));
KERNEL_END
KERNEL_BEGIN(k_fmnmx)
asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
KERNEL_END
// FADD(|x| - T) then the sign accumulated by IMAD.HI
KERNEL_BEGIN(k_fadd_imadhi)
asm volatile("{ .reg .f32 t; .reg .b32 u; abs.f32 t, %1; sub.f32 t, t, %2; mov.b32 u, t; mad.hi.u32 %0, u, 2, %0; }"
             : "+r"(cnt[i]) : "f"(a[i]), "f"(b));
a[i] += 0.f;
KERNEL_END
// FADD(|x| - T) then the sign accumulated by shift + add (LEA.HI?)
KERNEL_BEGIN(k_fadd_shradd)
asm volatile("{ .reg .f32 t; .reg .b32 u; abs.f32 t, %1; sub.f32 t, t, %2; mov.b32 u, t; shr.u32 u, u, 31; add.s32 %0, %0, u; }"
             : "+r"(cnt[i]) : "f"(a[i]), "f"(b));
a[i] += 0.f;
KERNEL_END

__global__ void __launch_bounds__(THREADS) k_ffma2(float* out, long long* cyc, float seed) {
    unsigned long long a[CH];
    unsigned long long b, c;
    float bl = seed * 1.0001f + threadIdx.x * 1e-7f;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bl));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(bl * 0.3f));
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float v = seed + i + threadIdx.x;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(v));
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
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

typedef void (*kern_t)(float*, long long*, float);

static void run(const char* name, kern_t k, double lane_ops_per_chain_step, int sms, float* out, long long* cyc) {
    k<<<sms, THREADS>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    k<<<sms, THREADS>>>(out, cyc, 1.5f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-22s failed: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double med = (double)h[sms / 2];
    double steps = (double)THREADS * ITER * CH;
    printf("%-22s %9.0f cycles  %7.2f chain-steps/clk/SM  %7.2f lane-ops/clk/SM\n", name, med, steps / med,
           steps * lane_ops_per_chain_step / med);
}

// ------------------------------------------------------------------ part 2: the leave-one-out loop
constexpr int LT = 128, LQ = 2, LDP = 8, LTILE = 512;

__device__ __forceinline__ float max3f(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t add.f32x2 rc, ra, rb;\n\t mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t sub.f32x2 rc, ra, rb;\n\t mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t mov.b64 rc, {%6, %7};\n\t"
        " fma.rn.f32x2 rd, ra, rb, rc;\n\t mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
struct IndScale { float neg_b, tb; };
__device__ __forceinline__ IndScale make_ind_scale(float T) {
    IndScale s;
    const int e = (int)((__float_as_uint(T) >> 23) & 0xffu);
    int be = 354 - e;
    be = be > 253 ? 253 : (be < 1 ? 1 : be);
    const float b = __uint_as_float((unsigned)be << 23);
    s.neg_b = -b; s.tb = T * b;
    return s;
}
__device__ __forceinline__ float ind_lt(float m, IndScale s) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(s.neg_b), "f"(s.tb));
    return r;
}
__device__ __forceinline__ void loo8_maxima(const float* a, float* m) {
    const float p01 = fmaxf(a[0], a[1]), p23 = fmaxf(a[2], a[3]);
    const float p45 = fmaxf(a[4], a[5]), p67 = fmaxf(a[6], a[7]);
    const float lo = fmaxf(p01, p23), hi = fmaxf(p45, p67);
    m[0] = max3f(a[1], p23, hi); m[1] = max3f(a[0], p23, hi);
    m[2] = max3f(a[3], p01, hi); m[3] = max3f(a[2], p01, hi);
    m[4] = max3f(a[5], p67, lo); m[5] = max3f(a[4], p67, lo);
    m[6] = max3f(a[7], p45, lo); m[7] = max3f(a[6], p45, lo);
}

// V = 0 maxima + saturating-FMA indicators + packed float counters (round-1 kernel)
//     1 per-coordinate indicators, packed indicator sum, [sigma == 8] / [sigma == 7] (all FMA pipe)
//     2 pair maxima on the ALU pipe, per-coordinate + per-pair-complement indicators
//     3 maxima + FADD(m - T) sign, accumulated with shift + add (integer counters)
//     4 maxima + FADD(m - T) sign, accumulated with IMAD.HI
//     5 per-coordinate FADD(|d| - T) signs funnel-shifted into a mask; contribution byte; spread by IMAD into
//       two words of four 8-bit counters (flushed every 128 candidates)
//     6 as 5 with the mask built by IMAD.HI chains (sign bits summed with weights)
template <int V>
__global__ void __launch_bounds__(LT) loo_loop(const float* __restrict__ tile_g, const float* __restrict__ qs,
                                               const float* __restrict__ Ts, int reps, int* __restrict__ counts) {
    __shared__ __align__(16) float tile[LTILE * LDP];
    for (int i = threadIdx.x; i < LTILE * LDP; i += LT) tile[i] = tile_g[i];
    __syncthreads();
    float2 q[LQ][LDP / 2];
    float T[LQ];
    IndScale sc[LQ];
    int cnt[LQ][8];
    float so[LQ];
    for (int j = 0; j < LQ; ++j) {
        const int qi = (blockIdx.x * LQ + j) * LT + threadIdx.x;
        for (int h = 0; h < LDP / 2; ++h) q[j][h] = make_float2(qs[qi * LDP + 2 * h], qs[qi * LDP + 2 * h + 1]);
        T[j] = Ts[qi];
        sc[j] = make_ind_scale(T[j]);
        so[j] = -(nextafterf(T[j], 0.0f) * -sc[j].neg_b);
        for (int s = 0; s < 8; ++s) cnt[j][s] = 0;
    }
    for (int rep = 0; rep < reps; ++rep) {
        if constexpr (V == 0 || V == 1 || V == 2 || V >= 7) {
            float2 acc[LQ][4], ae[LQ];
#pragma unroll
            for (int j = 0; j < LQ; ++j) {
                ae[j] = make_float2(0.f, 0.f);
#pragma unroll
                for (int h = 0; h < 4; ++h) acc[j][h] = make_float2(0.f, 0.f);
            }
#pragma unroll 2
            for (int c = 0; c < LTILE; ++c) {
                float2 cc[4];
                const float4 lo4 = *reinterpret_cast<const float4*>(tile + c * LDP);
                const float4 hi4 = *reinterpret_cast<const float4*>(tile + c * LDP + 4);
                cc[0] = make_float2(lo4.x, lo4.y); cc[1] = make_float2(lo4.z, lo4.w);
                cc[2] = make_float2(hi4.x, hi4.y); cc[3] = make_float2(hi4.z, hi4.w);
#pragma unroll
                for (int j = 0; j < LQ; ++j) {
                    if constexpr (V == 0) {
                        float a[8], m[8];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            a[2 * h] = fabsf(d.x); a[2 * h + 1] = fabsf(d.y);
                        }
                        loo8_maxima(a, m);
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            acc[j][h] = add2(acc[j][h], make_float2(ind_lt(m[2 * h], sc[j]), ind_lt(m[2 * h + 1], sc[j])));
                    } else if constexpr (V == 1) {
                        float2 ind[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            ind[h] = make_float2(ind_lt(fabsf(d.x), sc[j]), ind_lt(fabsf(d.y), sc[j]));
                        }
                        float2 s2 = add2(add2(ind[0], ind[1]), add2(ind[2], ind[3]));
                        const float sigma = s2.x + s2.y;
                        float all, ge;
                        asm("add.sat.f32 %0, %1, 0fC0E00000;" : "=f"(all) : "f"(sigma));  // sigma - 7
                        asm("add.sat.f32 %0, %1, 0fC0C00000;" : "=f"(ge) : "f"(sigma));   // sigma - 6
                        const float e = ge - all;
                        ae[j] = add2(ae[j], make_float2(all, e));
#pragma unroll
                        for (int h = 0; h < 4; ++h) acc[j][h] = fma2(ind[h], make_float2(e, e), acc[j][h]);
                    } else if constexpr (V == 7) {
                        // "outside" indicators o_d = [a_d >= T] = sat((a_d - T') B), T' = the float below T;
                        // sigma = number of outside coordinates; scalar FMAs into eight float counters
                        float o[8];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(o[2 * h]) : "f"(fabsf(d.x)), "f"(-sc[j].neg_b), "f"(so[j]));
                            asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(o[2 * h + 1]) : "f"(fabsf(d.y)), "f"(-sc[j].neg_b), "f"(so[j]));
                        }
                        const float2 s2 = add2(add2(make_float2(o[0], o[1]), make_float2(o[2], o[3])),
                                               add2(make_float2(o[4], o[5]), make_float2(o[6], o[7])));
                        const float sigma = s2.x + s2.y;
                        float all, le1;
                        asm("add.sat.f32 %0, %1, 0f3F800000;" : "=f"(all) : "f"(-sigma));  // 1 - sigma
                        asm("add.sat.f32 %0, %1, 0f40000000;" : "=f"(le1) : "f"(-sigma));  // 2 - sigma
                        const float e = le1 - all;
                        ae[j].x += all;
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            acc[j][h].x = __fmaf_rn(e, o[2 * h], acc[j][h].x);
                            acc[j][h].y = __fmaf_rn(e, o[2 * h + 1], acc[j][h].y);
                        }
                    } else if constexpr (V == 8) {
                        // inside indicators; products: j_h = i_2h i_2h+1, I_h = product of the other pairs' j
                        float i8[8], jh[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            i8[2 * h] = ind_lt(fabsf(d.x), sc[j]);
                            i8[2 * h + 1] = ind_lt(fabsf(d.y), sc[j]);
                            jh[h] = i8[2 * h] * i8[2 * h + 1];
                        }
                        const float j01 = jh[0] * jh[1], j23 = jh[2] * jh[3];
                        const float I[4] = {jh[1] * j23, jh[0] * j23, jh[3] * j01, jh[2] * j01};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            acc[j][h].x = __fmaf_rn(i8[2 * h + 1], I[h], acc[j][h].x);  // subspace 2h: its partner inside
                            acc[j][h].y = __fmaf_rn(i8[2 * h], I[h], acc[j][h].y);
                        }
                    } else if constexpr (V == 9) {
                        // V1 with the packed FMAs replaced by scalar ones
                        float2 ind[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            ind[h] = make_float2(ind_lt(fabsf(d.x), sc[j]), ind_lt(fabsf(d.y), sc[j]));
                        }
                        float2 s2 = add2(add2(ind[0], ind[1]), add2(ind[2], ind[3]));
                        const float sigma = s2.x + s2.y;
                        float all, ge;
                        asm("add.sat.f32 %0, %1, 0fC0E00000;" : "=f"(all) : "f"(sigma));
                        asm("add.sat.f32 %0, %1, 0fC0C00000;" : "=f"(ge) : "f"(sigma));
                        const float e = ge - all;
                        ae[j].x += all; ae[j].y += e;
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            acc[j][h].x = __fmaf_rn(e, ind[h].x, acc[j][h].x);
                            acc[j][h].y = __fmaf_rn(e, ind[h].y, acc[j][h].y);
                        }
                    } else {
                        float2 ind[4];
                        float p[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            const float ax = fabsf(d.x), ay = fabsf(d.y);
                            p[h] = fmaxf(ax, ay);
                            ind[h] = make_float2(ind_lt(ax, sc[j]), ind_lt(ay, sc[j]));
                        }
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float it = ind_lt(max3f(p[(h + 1) & 3], p[(h + 2) & 3], p[(h + 3) & 3]), sc[j]);
                            acc[j][h] = fma2(ind[h], make_float2(it, it), acc[j][h]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < LQ; ++j)
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float lo = acc[j][s / 2].x, hi = acc[j][s / 2].y;
                    if constexpr (V == 0 || V == 8) cnt[j][s] += (int)((s & 1) ? hi : lo);
                    else if constexpr (V == 1 || V == 9) cnt[j][s] += (int)(ae[j].x + (ae[j].y - ((s & 1) ? hi : lo)));
                    else if constexpr (V == 7) cnt[j][s] += (int)(ae[j].x + ((s & 1) ? hi : lo));
                    else cnt[j][s] += (int)((s & 1) ? lo : hi);
                }
        } else if constexpr (V == 3 || V == 4) {
#pragma unroll 2
            for (int c = 0; c < LTILE; ++c) {
                float2 cc[4];
                const float4 lo4 = *reinterpret_cast<const float4*>(tile + c * LDP);
                const float4 hi4 = *reinterpret_cast<const float4*>(tile + c * LDP + 4);
                cc[0] = make_float2(lo4.x, lo4.y); cc[1] = make_float2(lo4.z, lo4.w);
                cc[2] = make_float2(hi4.x, hi4.y); cc[3] = make_float2(hi4.z, hi4.w);
#pragma unroll
                for (int j = 0; j < LQ; ++j) {
                    float a[8], m[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float2 d = sub2(q[j][h], cc[h]);
                        a[2 * h] = fabsf(d.x); a[2 * h + 1] = fabsf(d.y);
                    }
                    loo8_maxima(a, m);
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const unsigned x = __float_as_uint(m[s] - T[j]);  // sign set <=> m < T
                        if constexpr (V == 3) cnt[j][s] += (int)(x >> 31);
                        else cnt[j][s] = (int)(__umulhi(x, 2u) + (unsigned)cnt[j][s]);
                    }
                }
            }
        } else {
            unsigned wa[LQ], wb[LQ];
            for (int c0 = 0; c0 < LTILE; c0 += 128) {
#pragma unroll
                for (int j = 0; j < LQ; ++j) wa[j] = wb[j] = 0u;
#pragma unroll 2
                for (int c = c0; c < c0 + 128; ++c) {
                    float2 cc[4];
                    const float4 lo4 = *reinterpret_cast<const float4*>(tile + c * LDP);
                    const float4 hi4 = *reinterpret_cast<const float4*>(tile + c * LDP + 4);
                    cc[0] = make_float2(lo4.x, lo4.y); cc[1] = make_float2(lo4.z, lo4.w);
                    cc[2] = make_float2(hi4.x, hi4.y); cc[3] = make_float2(hi4.z, hi4.w);
#pragma unroll
                    for (int j = 0; j < LQ; ++j) {
                        unsigned mask = 0u;  // bit d set <=> coordinate d inside
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 d = sub2(q[j][h], cc[h]);
                            const unsigned x = __float_as_uint(fabsf(d.x) - T[j]);
                            const unsigned y = __float_as_uint(fabsf(d.y) - T[j]);
                            if constexpr (V == 5) {
                                mask = __funnelshift_l(x, mask, 1);  // (mask << 1) | sign(x): coordinate 2h ends at bit 7 - 2h
                                mask = __funnelshift_l(y, mask, 1);
                            } else {
                                mask = __umulhi(x, 2u << (7 - 2 * h)) + mask;  // sign(x) << (7 - 2h), others spill below: masked next
                                mask = __umulhi(y, 2u << (6 - 2 * h)) + mask;
                            }
                        }
                        // (V == 6: umulhi(x, 2^(k+1)) = x >> (31 - k) keeps lower bits of x too; it only serves as a cost model)
                        const unsigned out = ~mask & 0xffu;                 // coordinates outside
                        const unsigned single = out & (out - 1u);           // 0 <=> at most one outside
                        unsigned contrib = out ? out : 0xffu;               // the subspaces this candidate counts in
                        contrib = single ? 0u : contrib;
                        wa[j] += ((contrib & 0xfu) * 0x00204081u) & 0x01010101u;
                        wb[j] += ((contrib >> 4) * 0x00204081u) & 0x01010101u;
                    }
                }
#pragma unroll
                for (int j = 0; j < LQ; ++j)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        // bit b of the mask is coordinate 7 - b: low nibble = coordinates 7..4, high nibble = 3..0
                        cnt[j][7 - s] += (int)((wa[j] >> (8 * s)) & 0xffu);
                        cnt[j][3 - s] += (int)((wb[j] >> (8 * s)) & 0xffu);
                    }
            }
        }
    }
    for (int j = 0; j < LQ; ++j) {
        const int qi = (blockIdx.x * LQ + j) * LT + threadIdx.x;
        for (int s = 0; s < 8; ++s) counts[qi * 8 + s] = cnt[j][s];
    }
}

template <int V>
static double run_loop(const char* name, int blocks, int reps, const float* tile, const float* qs, const float* Ts,
                       int* counts, std::vector<int>& host) {
    loo_loop<V><<<blocks, LT>>>(tile, qs, Ts, 2, counts);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    loo_loop<V><<<blocks, LT>>>(tile, qs, Ts, reps, counts);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s failed: %s\n", name, cudaGetErrorString(e)); return 0; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double pairs = (double)blocks * LT * LQ * LTILE * reps;
    host.resize((size_t)blocks * LT * LQ * 8);
    cudaMemcpy(host.data(), counts, host.size() * sizeof(int), cudaMemcpyDeviceToHost);
    printf("%-60s %8.3f ms  %7.3f Tpairs/s  (n = 250000 leave-one-out pass: %6.1f ms)\n", name, ms, pairs / ms * 1e-9,
           62.5e9 / (pairs / ms));
    return ms;
}

int main(int argc, char** argv) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("SMs %d, max SM clock %d kHz\n", sms, khz);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * THREADS);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    run("fadd", k_fadd, 1, sms, out, cyc);
    run("fmnmx", k_fmnmx, 1, sms, out, cyc);
    run("ffma 3 regs", k_ffma3, 1, sms, out, cyc);
    run("ffma.sat 3 regs", k_ffma3_sat, 1, sms, out, cyc);
    run("ffma.sat |a| 3 regs", k_ffma3_sat_abs, 1, sms, out, cyc);
    run("ffma imm", k_ffma_imm, 1, sms, out, cyc);
    run("ffma.sat imm", k_ffma_imm_sat, 1, sms, out, cyc);
    run("fmul.sat imm", k_fmul_imm_sat, 1, sms, out, cyc);
    run("fadd.sat", k_fadd_sat, 1, sms, out, cyc);
    run("fadd |a| - b", k_fadd_abs, 1, sms, out, cyc);
    run("ffma2 (x2)", k_ffma2, 2, sms, out, cyc);
    run("imad", k_imad, 1, sms, out, cyc);
    run("imad.hi", k_imad_hi, 1, sms, out, cyc);
    run("shr+add", k_shr_add, 1, sms, out, cyc);
    run("lop3", k_lop3, 1, sms, out, cyc);
    run("iadd3", k_iadd3, 1, sms, out, cyc);
    run("fadd+imad.hi", k_fadd_imadhi, 2, sms, out, cyc);
    run("fadd+shr+add", k_fadd_shradd, 2, sms, out, cyc);

    // ---- part 2
    const int blocks = sms * 12, reps = 64;
    const int nq = blocks * LT * LQ;
    std::vector<float> h_tile(LTILE * LDP), h_q((size_t)nq * LDP), h_T(nq);
    srand(7);
    auto rnd = []() { return (float)rand() / RAND_MAX; };
    for (auto& v : h_tile) v = rnd();
    for (auto& v : h_q) v = rnd();
    for (auto& v : h_T) v = 0.35f + 0.3f * rnd();  // ~1 % of the candidates inside all eight coordinates
    float *tile, *qs, *Ts; int* counts;
    cudaMalloc(&tile, h_tile.size() * 4); cudaMalloc(&qs, h_q.size() * 4); cudaMalloc(&Ts, h_T.size() * 4);
    cudaMalloc(&counts, (size_t)nq * 8 * 4);
    cudaMemcpy(tile, h_tile.data(), h_tile.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(qs, h_q.data(), h_q.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(Ts, h_T.data(), h_T.size() * 4, cudaMemcpyHostToDevice);
    std::vector<int> ref, got;
    run_loop<0>("V0 maxima + FFMA.SAT indicators + FADD2 counters", blocks, reps, tile, qs, Ts, counts, ref);
    long long total = 0;
    for (int v : ref) total += v;
    printf("   (mean count per subspace and query: %.1f of %d candidates)\n", (double)total / ref.size(), LTILE * reps);
    auto check = [&](const char* name) {
        size_t bad = 0;
        for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
        printf("   %s: %zu of %zu counts differ from V0\n", name, bad, ref.size());
    };
    run_loop<1>("V1 per-coordinate indicators, packed sums (FMA pipe only)", blocks, reps, tile, qs, Ts, counts, got); check("V1");
    run_loop<2>("V2 pair maxima (ALU) + 12 indicators + FFMA2", blocks, reps, tile, qs, Ts, counts, got); check("V2");
    run_loop<7>("V7 outside indicators, sigma, scalar FFMA counters", blocks, reps, tile, qs, Ts, counts, got); check("V7");
    run_loop<8>("V8 inside indicators, product tree, scalar FFMA counters", blocks, reps, tile, qs, Ts, counts, got); check("V8");
    run_loop<9>("V9 = V1 with scalar FFMA counters", blocks, reps, tile, qs, Ts, counts, got); check("V9");
    run_loop<3>("V3 maxima + FADD sign + shift/add integer counters", blocks, reps, tile, qs, Ts, counts, got); check("V3");
    run_loop<4>("V4 maxima + FADD sign + IMAD.HI integer counters", blocks, reps, tile, qs, Ts, counts, got); check("V4");
    run_loop<5>("V5 FADD signs -> SHF mask -> contribution byte -> IMAD spread", blocks, reps, tile, qs, Ts, counts, got); check("V5");
    run_loop<6>("V6 cost model: as V5, mask by IMAD.HI (counts not meaningful)", blocks, reps, tile, qs, Ts, counts, got);
    return 0;
}

// Shared helpers for the prosper_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/prosper_b200.h"

namespace pet {

void set_error(const char *fmt, ...);
extern thread_local int64_t g_launches;   // kernels launched on this thread (all engines)

#define PET_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            pet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                           __FILE__, __LINE__);                                            \
            return PET_ECUDA;                                                              \
        }                                                                                  \
    } while (0)

#define PET_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        pet::g_launches++;                                                                 \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            pet::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),     \
                           __FILE__, __LINE__);                                            \
            return PET_ECUDA;                                                              \
        }                                                                                  \
    } while (0)

#define PET_CHECK(expr)                                                                    \
    do {                                                                                   \
        int _r = (expr);                                                                   \
        if (_r != PET_OK) return _r;                                                       \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---- device helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared; src_bytes in {0,8,16}, the rest is zero-filled.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// FP64 tensor-core tile: D(8x8) = A(8x4, row) * B(4x8, col) + C.  Fragment ownership
// (PTX ISA, mma.m8n8k4 .f64): a0 = A[lane/4][lane%4]; b0 = B[lane%4][lane/4];
// c0,c1 = C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// exp(x) for -700 < x <= ~0.5 (callers pass log-joint minus the running maximum, cut off at -100): round-to-nearest
// range reduction x = n ln2 + r, |r| <= 0.347, degree-13 Taylor (truncation 4e-18), exponent added to the high
// word.  ~1 ulp, 19 instructions and no branches (libdevice exp is ~45 with its range checks).  The coefficients
// sit in constant memory so that every DFMA takes its constant as a c[][] operand; as immediates each one costs
// two extra MOVs per evaluation (measured: 57 instead of 19 instructions per call).
static __constant__ double c_expk[18] = {
    6755399441055744.0,              // [0] 2^52 + 2^51
    1.4426950408889634074,           // [1] log2(e)
    -6.93147180369123816490e-01,     // [2] -ln2 (high part)
    -1.90821492927058770002e-10,     // [3] -ln2 (low part)
    1.6059043836821613e-10,          // [4] 1/13!
    2.0876756987868100e-09, 2.5052108385441720e-08, 2.7557319223985890e-07, 2.7557319223985893e-06,
    2.4801587301587302e-05, 1.9841269841269841e-04, 1.3888888888888889e-03, 8.3333333333333332e-03,
    4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};

__device__ __forceinline__ double exp_nonpos(double x) {
    const double t = fma(x, c_expk[1], c_expk[0]);
    const int n = __double2loint(t);
    const double fn = t - c_expk[0];
    double r = fma(fn, c_expk[2], x);
    r = fma(fn, c_expk[3], r);
    double p = c_expk[4];
#pragma unroll
    for (int i = 5; i < 18; ++i) p = fma(p, r, c_expk[i]);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// exp(x) for -700 < x <= 0 with a 32-entry table: x = (32 k + j) ln2/32 + r, |r| <= ln2/64, exp = 2^k T[j] e^r with a
// degree-6 Taylor polynomial (truncation 3.5e-18): 11 FP64 instructions instead of the 17 of exp_nonpos, ~1.5 ulp.
static __constant__ double c_exp32[12] = {
    6755399441055744.0,               // 2^52 + 2^51
    46.166241308446828384,            // 32 / ln2
    -0.021660849393811077,            // -ln2/32, high part (33 significant bits: n * hi is exact)
    1.312785960212839e-12,            // -ln2/32, low part
    1.0 / 720, 1.0 / 120, 1.0 / 24, 1.0 / 6, 0.5, 1.0, 1.0, 0.0};
__device__ __forceinline__ double exp_tab32(double x, const double *tab) {
    const double t = fma(x, c_exp32[1], c_exp32[0]);
    const int n = __double2loint(t);
    const double fn = t - c_exp32[0];
    double r = fma(fn, c_exp32[2], x);
    r = fma(fn, c_exp32[3], r);
    double p = c_exp32[4];
#pragma unroll
    for (int i = 5; i < 11; ++i) p = fma(p, r, c_exp32[i]);
    p *= tab[n & 31];
    return __hiloint2double(__double2hiint(p) + ((n >> 5) << 20), __double2loint(p));
}
// fills the 32-entry table 2^(j/32) (shared memory; the callers synchronise before the first use)
__device__ __forceinline__ void exp_tab32_init(double *tab) {
    if (threadIdx.x < 32) tab[threadIdx.x] = exp2(double(threadIdx.x) * 0.03125);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace pet

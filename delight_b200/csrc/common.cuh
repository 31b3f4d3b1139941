// Shared device helpers for the delight_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define DL_OK 0
#define DL_ERR_INVALID 1
#define DL_ERR_CUDA 2
#define DL_ERR_NOMEM 3
#define DL_ERR_UNSUPPORTED 4

void dl_set_error(const char* fmt, ...);
void dl_count_launch(int n = 1);                    // bookkeeping: kernels launched by this library
bool dl_timing_enabled();
void dl_timing_mark(cudaStream_t stream, int which);  // 0 = before, 1 = after the fused kernel

#define DL_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            dl_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                         cudaGetErrorString(_e));                                        \
            return DL_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define DL_REQUIRE(cond, msg)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            dl_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);        \
            return DL_ERR_INVALID;                                                       \
        }                                                                                \
    } while (0)

#define DL_TWO_PI 6.283185307179586476925286766559

namespace dl {

// ------------------------------------------------------------------ FP64 fast-path arithmetic
// MUFU.RCP64H seed + ONE third-order step x1 = x0 (1 + e + e^2), e = 1 - d x0: 3 dependent FP64
// ops.  The relative error is e^3; with the seed's measured |e| <= 2^-19.9 (tests/test_gpu_parity.py::
// test_fast_math_accuracy) that is < 2^-59, i.e. limited by the final rounding.  Valid for
// normal, positive finite inputs (variances, sums of squares); no special-case branches.
__device__ __forceinline__ double rcp_fast(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double e = fma(-d, x, 1.0);
    const double t = fma(e, e, e);
    return fma(x, t, x);
}

// MUFU.RSQ64H seed + one third-order step y1 = y0 (1 + h/2 + 3 h^2 / 8), h = 1 - d y0^2
// (5 dependent FP64 ops, error ~ h^3 / 3).  Same validity domain as rcp_fast.
__device__ __forceinline__ double rsqrt_fast(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double h = fma(-t, y, 1.0);
    const double q = fma(h, 0.375, 0.5);
    return fma(y, h * q, y);
}

// exp(x) for x <= 0 (-inf / x < -708 -> 0: results that would be denormal are flushed; NaN
// propagates).  Range reduction x = n ln2 + r with a single-constant ln2 (|n| <= 1022, so the
// reduction error is < 2e-14 absolute), |r| <= ln2/2, degree-9 Taylor polynomial (truncation
// 7e-12 relative -- two orders below the 1e-9 parity contract) evaluated as even + r * odd parts
// in r^2 (two independent Horner chains of depth 4), scaling through the exponent field.
// 14 FP64-pipe instructions.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double L2E = 1.4426950408889634074;
    const double LN2 = 6.93147180559945286227e-01;
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
    double t = fma(x, L2E, SHIFT);
    int n = __double2loint(t);
    double fn = t - SHIFT;
    const double r = fma(-fn, LN2, x);
    const double r2 = r * r;
    double pe = 2.48015873015873015873e-05;            // 1/8!
    double po = 2.75573192239858906526e-06;            // 1/9!
    pe = fma(pe, r2, 1.38888888888888888889e-03);      // 1/6!
    po = fma(po, r2, 1.98412698412698412698e-04);      // 1/7!
    pe = fma(pe, r2, 4.16666666666666666667e-02);      // 1/4!
    po = fma(po, r2, 8.33333333333333333333e-03);      // 1/5!
    pe = fma(pe, r2, 0.5);
    po = fma(po, r2, 1.66666666666666666667e-01);      // 1/3!
    pe = fma(pe, r2, 1.0);
    po = fma(po, r2, 1.0);
    const double p = fma(po, r, pe);
    // scale by 2^n through the exponent bits; results below 2^-1021 flush to zero
    int hi = __double2hiint(p) + (n << 20);
    double res = __hiloint2double(hi, __double2loint(p));
    // x >= -708 is false for NaN: return the NaN itself then (detected on the exponent bits by
    // the integer pipe), 0 for underflow / -inf
    return (x >= -708.0) ? res : ((__double2hiint(x) & 0x7fffffff) > 0x7ff00000 ? x : 0.0);
}

// Table variant of the same function: x = (N k + j) ln2/N + r, |r| <= ln2/(2N), so a short
// polynomial is enough and exp(x) = 2^k * tab[j] * p(r) with tab[j] = 2^(j/N) in shared memory.
//   N = 32:  degree 4 (truncation r^5/120 <= 1.3e-12 relative), 8 FP64-pipe instructions
//   N = 128: degree 3 (truncation r^4/24  <= 2.3e-12 relative), 7 FP64-pipe instructions
// (N = 1024 with degree 2 measured no faster: 1.268 s.)
// instead of 14; the range test runs on the integer pipe.
#ifndef DL_EXPTAB_BITS
#define DL_EXPTAB_BITS 7
#endif
constexpr int EXP_TAB = 1 << DL_EXPTAB_BITS;
__device__ __forceinline__ void exp_tab_init(double* tab, int tid, int nthreads) {
    for (int j = tid; j < EXP_TAB; j += nthreads) tab[j] = exp2((double)j / EXP_TAB);
}
constexpr unsigned EXP_UNDER = 0xFFF00000u - 0xC0862000u;      // see `under` below
constexpr unsigned EXP_ZERO = 0xFFFFFFFFu;                    // as `lim`: the result is 0 for every x
__device__ __forceinline__ double exp_nonpos_tab(double x, const double* tab,
                                                 unsigned lim = EXP_UNDER) {
    const double L2EN = 1.4426950408889634074 * EXP_TAB;           // N / ln 2
    const double LN2N = 6.93147180559945286227e-01 / EXP_TAB;      // ln 2 / N (exact scaling)
    const double SHIFT = 6755399441055744.0;                       // 1.5 * 2^52
    const double t = fma(x, L2EN, SHIFT);
    const int n = __double2loint(t);
    const double fn = t - SHIFT;
    const double r = fma(-fn, LN2N, x);
#if DL_EXPTAB_BITS >= 7
    double p = fma(r, 1.66666666666666666667e-01, 0.5);
#else
    double p = fma(r, 4.16666666666666666667e-02, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
#endif
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= tab[n & (EXP_TAB - 1)];
    const int hi = __double2hiint(p) + ((n >> DL_EXPTAB_BITS) << 20);
    const double res = __hiloint2double(hi, __double2loint(p));
    // x < -708 (incl. -inf) -> 0, decided on the high word (integer pipe).  A NaN needs no test:
    // every NaN the GPU produces or NumPy hands over has a zero low word (0x7ff8... / 0xfff8...),
    // so n = 0, the exponent is left alone and p (= NaN) comes back as it is.
    const unsigned hx = (unsigned)__double2hiint(x);
    // (`lim` = EXP_ZERO is the caller's own reason for a 0 -- a padded cell -- at no extra cost)
    const bool under = hx - 0xC0862000u <= lim;                           // -708 >= x >= -inf
    return under ? 0.0 : res;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// ------------------------------------------------------------------ mbarrier + TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifndef DL_WAIT_HINT
#define DL_WAIT_HINT 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// Wait with back-off: a waiter that just re-issues try_wait competes for issue slots with the
// very warps (on its own scheduler) it is waiting for; __nanosleep takes it out of the way.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if DL_WAIT_HINT
    // try_wait with a suspend-time hint: the hardware parks the warp (no issue slots) until the
    // phase completes or the hint expires, instead of a software spin with nanosleep
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)DL_WAIT_HINT)
            : "memory");
    } while (!done);
#else
    if (mbar_try_wait(bar, parity)) return;
    unsigned ns = 32;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
    }
#endif
}

// One lane of a converged warp (elect.sync): code under this predicate is known to run in exactly
// one thread, so the uniform-datapath instructions below it (UBLKCP, SYNCS) need no
// "for every active thread" loop around them.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "elect.sync _|P1, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem,
                                             uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace dl

// Does a warp that waits on an mbarrier slow down the warps that still compute on the same SM?
// CTA = 8 warps (2 CTAs/SM as in the fused kernel).  NWORK warps run a fixed FP64 + LDS
// workload, the others wait for them on an mbarrier with one of several waiting methods.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t d;
    asm volatile("{.reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0,1,0,P1;}"
                 : "=r"(d) : "r"(smem_u32(b)), "r"(par) : "memory");
    return d;
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t par) {
    uint32_t d;
    asm volatile("{.reg .pred P1; mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0,1,0,P1;}"
                 : "=r"(d) : "r"(smem_u32(b)), "r"(par) : "memory");
    return d;
}

// MODE 0: waiters exit at once; 1: try_wait + nanosleep back-off (the fused kernel's mbar_wait);
// 2: try_wait tight loop; 3: test_wait + fixed nanosleep(256); 4: bar.sync with the workers;
// 5: try_wait by lane 0 only, other lanes parked at __syncwarp
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(int iters, int nwork, int rounds, double* out, long long* cyc) {
    __shared__ uint64_t bar;
    __shared__ double tab[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, nwork); asm volatile("fence.mbarrier_init.release.cluster;"); }
    tab[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
    __syncthreads();
    double a0 = 1.0 + lane * 1e-3, a1 = 2.0 + lane * 1e-3, s = 0.0;
    const long long t0 = clock64();
    long long tcomp = 0; int ndiv = 0;
    for (int r = 0; r < rounds; ++r) {
        if (warp < nwork) {
            const long long tc0 = clock64();
            ndiv += __popc(__activemask()) != 32;
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                const double m = tab[(it + lane) & 255];
#pragma unroll
                for (int q = 0; q < 8; ++q) { a0 = fma(a0, m, 1e-9); a1 = fma(a1, m, 1e-9); }
            }
            __syncwarp();
            tcomp += clock64() - tc0;
            if (lane == 0) mbar_arrive(&bar);
            if (MODE == 4) asm volatile("bar.sync 1, 256;");
        } else {
            if (MODE == 0) continue;
            if (MODE == 4) asm volatile("bar.sync 1, 256;");
        }
        if (MODE == 0 && warp >= nwork) continue;
        const uint32_t par = r & 1;
        if (MODE == 1) {
            if (!mbar_try(&bar, par)) {
                unsigned ns = 32;
                while (!mbar_try(&bar, par)) { __nanosleep(ns); if (ns < 1024) ns <<= 1; }
            }
        } else if (MODE == 2) {
            while (!mbar_try(&bar, par)) {}
        } else if (MODE == 3) {
            while (!mbar_test(&bar, par)) __nanosleep(256);
        } else if (MODE == 5) {
            if (lane == 0) while (!mbar_try(&bar, par)) {}
            __syncwarp();
        } else if (MODE == 0) {
            while (!mbar_try(&bar, par)) {}
        }
    }
    const long long t1 = clock64();
    s = a0 + a1;
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0) { cyc[3 * blockIdx.x] = t1 - t0; cyc[3 * blockIdx.x + 1] = tcomp; cyc[3 * blockIdx.x + 2] = ndiv; }
}

template <int MODE>
void run(int nwork, double* out, long long* cyc, const char* name) {
    const int blocks = 296, iters = 400, rounds = 50;
    k<MODE><<<blocks, 256>>>(10, nwork, 2, out, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(iters, nwork, rounds, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per worker warp: iters*rounds*16 dependent-pair DFMAs
    const double dfma_per_warp = (double)iters * rounds * 16;
    long long h[3 * 296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double tot = 0, comp = 0, nd = 0;
    for (int b = 0; b < 296; ++b) { tot += h[3 * b]; comp += h[3 * b + 1]; nd += h[3 * b + 2]; }
    printf("%-34s workers/CTA %d: %7.3f ms  -> %.2f cyc per DFMA per worker warp | warp0: compute %.0f%% of lifetime, diverged rounds %.1f of %d\n",
           name, nwork, ms, ms * 1e-3 * 1.965e9 / dfma_per_warp, 100 * comp / tot, nd / 296, rounds);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 1024);
    for (int nw : {1, 2, 8}) {
        run<0>(nw, out, cyc, "waiters exit");
        run<1>(nw, out, cyc, "try_wait + nanosleep back-off");
        run<2>(nw, out, cyc, "try_wait tight loop");
        run<3>(nw, out, cyc, "test_wait + nanosleep(256)");
        run<4>(nw, out, cyc, "bar.sync");
        run<5>(nw, out, cyc, "lane-0 try_wait, rest syncwarp");
    }
    return 0;
}

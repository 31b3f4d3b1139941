// Do DMMA (mma.sync m8n8k4 f64) and DFMA share one execution pipe on B200?
// Three kernels with the same loop structure: NF dependent-chain DFMAs, ND DMMAs, or both per
// iteration.  If t(both) ~ max(t(dfma), t(dmma)) the pipes are separate; ~ sum: shared.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int ND>
__global__ void k(int iters, double* out) {
    double f[8], c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = 1.0 + threadIdx.x * 1e-3 + i; c0[i] = i; c1[i] = -i; }
    const double m = 0.999999, c = 1e-9, a = 1e-3 * (threadIdx.x & 3), b = 1e-3 * (threadIdx.x >> 2);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < ND) dmma(c0[i], c1[i], a, b);
#pragma unroll
                for (int j = 0; j < NF; ++j)
                    if (i * NF + j < 8 * NF) f[(i * NF + j) % 8] = fma(f[(i * NF + j) % 8], m, c);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template <int NF, int ND>
float run(int warps_per_sm, double* out, const char* name) {
    const int blocks = 148, threads = warps_per_sm * 32, iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NF, ND><<<blocks, threads>>>(10, out);
    cudaEventRecord(e0);
    k<NF, ND><<<blocks, threads>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double nfma = (double)blocks * threads * iters * 4.0 * 8 * NF;          // thread FMAs
    const double nmma = (double)blocks * warps_per_sm * iters * 4.0 * ND;          // warp DMMAs
    printf("%-22s warps/SM %2d: %8.3f ms   DFMA %6.2f TFLOP/s   DMMA %6.2f TFLOP/s  (%.1f cyc/DMMA/SMSP)\n",
           name, warps_per_sm, ms, 2 * nfma / (ms * 1e-3) * 1e-12, 512 * nmma / (ms * 1e-3) * 1e-12,
           nmma > 0 ? ms * 1e-3 * 1.965e9 / (nmma / (148.0 * 4)) : 0.0);
    return ms;
}

int main() {
    double* out; cudaMalloc(&out, 8);
    for (int w : {8, 16, 32}) {
        float a = run<4, 0>(w, out, "DFMA only (32/iter)");
        float b = run<0, 8>(w, out, "DMMA only (8/iter)");
        float c = run<4, 8>(w, out, "both (32 DFMA + 8 DMMA)");
        float d = run<4, 2>(w, out, "both (32 DFMA + 2 DMMA)");
        float e = run<0, 2>(w, out, "DMMA only (2/iter)");
        printf("   -> both/(dfma+dmma) = %.2f, both/max = %.2f ; light mix: both/(sum) = %.2f both/max = %.2f\n",
               c / (a + b), c / (a > b ? a : b), d / (a + e), d / (a > e ? a : e));
    }
    return 0;
}

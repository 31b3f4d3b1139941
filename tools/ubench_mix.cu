// Can other instructions issue in the second cycle of an FP64 instruction?  Per loop iteration:
// 16 independent-chain DFMAs + M integer (or LDS / FFMA) instructions per warp, 16 warps/SM.
// If t(M) stays flat until M ~ 16: yes (dual-rate pipe leaves issue slots free); if t grows as
// 2*16 + M from M = 0: an FP64 instruction holds the issue port for both of its cycles.
#include <cstdio>
#include <cuda_runtime.h>

template <int M, int KIND>
__global__ void __launch_bounds__(256, 2) k(int iters, double* out, const double* in) {
    __shared__ double sm[256];
    sm[threadIdx.x] = in[threadIdx.x & 7];
    __syncthreads();
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    int q[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = threadIdx.x + i; f[i] = 1.0f + i; }
    double ld = 0.0;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fma(a[i], m, c);
#pragma unroll
                for (int j = 0; j < (M + 15 - (r * 8 + i)) / 16; ++j) {
                    if (KIND == 0) q[(i + j) & 7] = q[(i + j) & 7] * 3 + it;          // IMAD
                    if (KIND == 1) f[(i + j) & 7] = fmaf(f[(i + j) & 7], 0.999f, 1e-3f);  // FFMA
                    if (KIND == 2) ld += sm[(threadIdx.x + q[0] + i * 8 + j) & 255];      // LDS (+DADD!)
                }
            }
        }
    }
    double s = ld;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + q[i] + f[i];
    if (s == 123.456) out[0] = s;
}

template <int M, int KIND>
void run(double* out, const double* in, const char* name) {
    const int blocks = 296, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<M, KIND><<<blocks, 256>>>(10, out, in);
    cudaEventRecord(e0);
    k<M, KIND><<<blocks, 256>>>(iters, out, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // 4 warps per SMSP, each 16 DFMA + M others per iteration
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 4;   // cycles per warp-iteration per SMSP
    printf("%-6s M=%2d: %7.3f ms  %6.1f cycles per (16 DFMA + %2d other)  [pipe bound 32, 2N+M = %d]\n",
           name, M, ms, cyc, M, 32 + M);
}

int main() {
    double *out, *in; cudaMalloc(&out, 8); cudaMalloc(&in, 64); cudaMemset(in, 0, 64);
    run<0, 0>(out, in, "IMAD"); run<4, 0>(out, in, "IMAD"); run<8, 0>(out, in, "IMAD");
    run<16, 0>(out, in, "IMAD"); run<24, 0>(out, in, "IMAD"); run<32, 0>(out, in, "IMAD");
    run<8, 1>(out, in, "FFMA"); run<16, 1>(out, in, "FFMA"); run<32, 1>(out, in, "FFMA");
    return 0;
}

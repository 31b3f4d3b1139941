// FP64 pipe microbenchmark: throughput vs (warps per SM, independent chains per thread)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(int iters, double* out) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}
template <int ILP>
void run(int warps_per_sm, double* out) {
    int blocks = 148, threads = warps_per_sm * 32;
    int iters = 20000 / ILP;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, threads>>>(10, out);
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * threads * iters * 16.0 * ILP;
    double tf = 2 * fmas / (ms * 1e-3) * 1e-12;
    // cycles per dependent step per warp at 1.965 GHz: time per (iters*16) steps
    double cyc_per_step = ms * 1e-3 * 1.965e9 / (iters * 16.0);
    printf("warps/SM %2d (%.0f/SMSP) ILP %d : %6.2f TFLOP/s  %5.1f cyc per chain-step (=%.1f cyc/DFMA/warp)\n",
           warps_per_sm, warps_per_sm / 4.0, ILP, tf, cyc_per_step, cyc_per_step / ILP);
}
int main() {
    double* out; cudaMalloc(&out, 8);
    int ws[] = {4, 8, 16, 24, 32, 64};
    for (int w : ws) { run<1>(w, out); run<2>(w, out); run<4>(w, out); run<8>(w, out); }
    return 0;
}

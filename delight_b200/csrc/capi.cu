// C-ABI glue: error reporting, device query, and HOST-pointer forms of the five reference
// operators (the entry points a foreign-function binding of delight.utils_cy /
// delight.photoz_kernels_cy would bind).  Device-pointer forms live next to their kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "../../include/delight_b200.h"

static thread_local char g_err[1024] = "";

void dl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        if (bytes == 0) bytes = 8;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            dl_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            return DL_ERR_NOMEM;
        }
        return DL_OK;
    }
    int upload(const void* h, size_t bytes) {
        int rc = alloc(bytes);
        if (rc) return rc;
        if (bytes) DL_CUDA_CHECK(cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice));
        return DL_OK;
    }
    template <class T> T* as() { return static_cast<T*>(p); }
};

}  // namespace

#define TRY(x) do { int _rc = (x); if (_rc != DL_OK) return _rc; } while (0)

// host-side range check of index arrays (the reference runs with boundscheck=False and would
// read out of bounds; on a GPU that is a fault, so the host entry points refuse instead)
static bool in_range(const long long* v, long long n, long long lo, long long hi) {
    for (long long i = 0; i < n; ++i)
        if (v[i] < lo || v[i] >= hi) return false;
    return true;
}

extern "C" {

const char* dl_last_error(void) { return g_err; }

int dl_version(void) { return 100; }

int dl_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* global_mem) {
    int dev = 0;
    DL_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DL_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (global_mem) *global_mem = (long long)prop.totalGlobalMem;
    return DL_OK;
}

// ---------------------------------------------------------------- host-pointer operator forms
int dl_approx_flux_likelihood_host(double* like, long long nz, long long nt, long long nf,
                                   const double* f_obs, const double* f_obs_var,
                                   const double* f_mod, const double* f_mod_covar,
                                   const double* ell_hat, const double* ell_var) {
    DL_REQUIRE(nz >= 0 && nt >= 0 && nf >= 0, "negative size");
    if (nz * nt == 0) return DL_OK;
    DevBuf dlike, dobs, dvar, dmod, dcov, deh, dev_, scr;
    const size_t cells = (size_t)nz * nt;
    TRY(dlike.alloc(cells * 8));
    TRY(dobs.upload(f_obs, nf * 8));
    TRY(dvar.upload(f_obs_var, nf * 8));
    TRY(dmod.upload(f_mod, cells * nf * 8));
    TRY(dcov.upload(f_mod_covar, cells * nf * 8));
    TRY(deh.upload(ell_hat, nz * 8));
    TRY(dev_.upload(ell_var, nz * 8));
    TRY(scr.alloc(8));
    TRY(dl_approx_flux_likelihood_dev(dlike.as<double>(), nz, nt, nf, dobs.as<double>(),
                                      dvar.as<double>(), dmod.as<double>(), dcov.as<double>(),
                                      deh.as<double>(), dev_.as<double>(), scr.p, nullptr));
    DL_CUDA_CHECK(cudaMemcpy(like, dlike.p, cells * 8, cudaMemcpyDeviceToHost));
    return DL_OK;
}

int dl_kernelparts_host(long long NO1, long long NO2, long long NC, long long NL, double alpha_C,
                        double alpha_L, long long nB, const double* amp, const double* mu,
                        const double* sig, const double* lines_mu, const double* norms,
                        const long long* b1, const double* fz1, const long long* b2,
                        const double* fz2, int grad_needed, double* KL, double* KC,
                        double* D_alpha_C, double* D_alpha_L, double* D_alpha_z) {
    DL_REQUIRE(NO1 >= 0 && NO2 >= 0 && nB >= 1, "bad sizes");
    if (NO1 * NO2 == 0) return DL_OK;
    DL_REQUIRE(in_range(b1, NO1, 0, nB) && in_range(b2, NO2, 0, nB), "band index out of range");
    DevBuf da, dm, ds, dl_, dn, db1, df1, db2, df2, o[5];
    TRY(da.upload(amp, nB * NC * 8));
    TRY(dm.upload(mu, nB * NC * 8));
    TRY(ds.upload(sig, nB * NC * 8));
    TRY(dl_.upload(lines_mu, NL * 8));
    TRY(dn.upload(norms, nB * 8));
    TRY(db1.upload(b1, NO1 * 8));
    TRY(df1.upload(fz1, NO1 * 8));
    TRY(db2.upload(b2, NO2 * 8));
    TRY(df2.upload(fz2, NO2 * 8));
    double* host[5] = {KL, KC, D_alpha_C, D_alpha_L, D_alpha_z};
    const size_t n = (size_t)NO1 * NO2 * 8;
    for (int k = 0; k < 5; ++k) TRY(o[k].upload(host[k], n));
    TRY(dl_kernelparts_dev(NO1, NO2, NC, NL, alpha_C, alpha_L, da.as<double>(), dm.as<double>(),
                           ds.as<double>(), dl_.as<double>(), dn.as<double>(),
                           db1.as<long long>(), df1.as<double>(), db2.as<long long>(),
                           df2.as<double>(), grad_needed, o[0].as<double>(), o[1].as<double>(),
                           o[2].as<double>(), o[3].as<double>(), o[4].as<double>(), NO2,
                           nullptr));
    const int nout = grad_needed ? 5 : 2;
    for (int k = 0; k < nout; ++k)
        DL_CUDA_CHECK(cudaMemcpy(host[k], o[k].p, n, cudaMemcpyDeviceToHost));
    return DL_OK;
}

int dl_kernelparts_diag_host(long long NO1, long long NC, long long NL, double alpha_C,
                             double alpha_L, long long nB, const double* amp, const double* mu,
                             const double* sig, const double* lines_mu, const double* norms,
                             const long long* b1, const double* fz1, int grad_needed, double* KL,
                             double* KC, double* D_alpha_C, double* D_alpha_L) {
    DL_REQUIRE(NO1 >= 0 && nB >= 1, "bad sizes");
    if (NO1 == 0) return DL_OK;
    DL_REQUIRE(in_range(b1, NO1, 0, nB), "band index out of range");
    DevBuf da, dm, ds, dl_, dn, db1, df1, o[4];
    TRY(da.upload(amp, nB * NC * 8));
    TRY(dm.upload(mu, nB * NC * 8));
    TRY(ds.upload(sig, nB * NC * 8));
    TRY(dl_.upload(lines_mu, NL * 8));
    TRY(dn.upload(norms, nB * 8));
    TRY(db1.upload(b1, NO1 * 8));
    TRY(df1.upload(fz1, NO1 * 8));
    double* host[4] = {KL, KC, D_alpha_C, D_alpha_L};
    for (int k = 0; k < 4; ++k) TRY(o[k].upload(host[k], NO1 * 8));
    TRY(dl_kernelparts_diag_dev(NO1, NC, NL, alpha_C, alpha_L, da.as<double>(), dm.as<double>(),
                                ds.as<double>(), dl_.as<double>(), dn.as<double>(),
                                db1.as<long long>(), df1.as<double>(), grad_needed,
                                o[0].as<double>(), o[1].as<double>(), o[2].as<double>(),
                                o[3].as<double>(), nullptr));
    const int nout = grad_needed ? 4 : 2;
    for (int k = 0; k < nout; ++k)
        DL_CUDA_CHECK(cudaMemcpy(host[k], o[k].p, NO1 * 8, cudaMemcpyDeviceToHost));
    return DL_OK;
}

int dl_find_positions_host(long long NO1, long long nz, const double* fz1, long long* p1s,
                           const double* fzGrid) {
    DL_REQUIRE(NO1 >= 0 && nz >= 0, "bad sizes");
    if (NO1 == 0 || nz < 2) return DL_OK;
    DevBuf df, dp, dg;
    TRY(df.upload(fz1, NO1 * 8));
    TRY(dp.upload(p1s, NO1 * 8));
    TRY(dg.upload(fzGrid, nz * 8));
    TRY(dl_find_positions_dev(NO1, nz, df.as<double>(), dp.as<long long>(), dg.as<double>(),
                              nullptr));
    DL_CUDA_CHECK(cudaMemcpy(p1s, dp.p, NO1 * 8, cudaMemcpyDeviceToHost));
    return DL_OK;
}

int dl_kernel_parts_interp_host(long long NO1, long long NO2, double* Kinterp,
                                const long long* b1, const double* fz1, const long long* p1s,
                                const long long* b2, const double* fz2, const long long* p2s,
                                const double* fzGrid, const double* Kgrid, long long nB,
                                long long nzg) {
    DL_REQUIRE(NO1 >= 0 && NO2 >= 0 && nB >= 1 && nzg >= 2, "bad sizes");
    if (NO1 * NO2 == 0) return DL_OK;
    DL_REQUIRE(in_range(b1, NO1, 0, nB) && in_range(b2, NO2, 0, nB), "band index out of range");
    DL_REQUIRE(in_range(p1s, NO1, 0, nzg - 1) && in_range(p2s, NO2, 0, nzg - 1),
               "bin position out of range");
    DevBuf dk, db1, df1, dp1, db2, df2, dp2, dg, dG;
    TRY(dk.alloc((size_t)NO1 * NO2 * 8));
    TRY(db1.upload(b1, NO1 * 8));
    TRY(df1.upload(fz1, NO1 * 8));
    TRY(dp1.upload(p1s, NO1 * 8));
    TRY(db2.upload(b2, NO2 * 8));
    TRY(df2.upload(fz2, NO2 * 8));
    TRY(dp2.upload(p2s, NO2 * 8));
    TRY(dg.upload(fzGrid, nzg * 8));
    TRY(dG.upload(Kgrid, (size_t)nB * nB * nzg * nzg * 8));
    TRY(dl_kernel_parts_interp_dev(NO1, NO2, dk.as<double>(), NO2, db1.as<long long>(),
                                   df1.as<double>(), dp1.as<long long>(), db2.as<long long>(),
                                   df2.as<double>(), dp2.as<long long>(), dg.as<double>(),
                                   dG.as<double>(), nB, nzg, nullptr));
    DL_CUDA_CHECK(cudaMemcpy(Kinterp, dk.p, (size_t)NO1 * NO2 * 8, cudaMemcpyDeviceToHost));
    return DL_OK;
}

int dl_bilininterp_precomputedbins_host(long long numBands, long long nobj, double* Kinterp,
                                        const double* v1s, const double* v2s,
                                        const long long* p1s, const long long* p2s,
                                        const double* grid1, const double* grid2,
                                        const double* Kgrid, long long n1, long long n2) {
    DL_REQUIRE(numBands >= 0 && nobj >= 0 && n1 >= 2 && n2 >= 2, "bad sizes");
    if (numBands * nobj == 0) return DL_OK;
    DL_REQUIRE(in_range(p1s, nobj, 0, n1 - 1) && in_range(p2s, nobj, 0, n2 - 1),
               "bin position out of range");
    DevBuf dk, dv1, dv2, dp1, dp2, dg1, dg2, dG;
    TRY(dk.alloc((size_t)numBands * nobj * 8));
    TRY(dv1.upload(v1s, nobj * 8));
    TRY(dv2.upload(v2s, nobj * 8));
    TRY(dp1.upload(p1s, nobj * 8));
    TRY(dp2.upload(p2s, nobj * 8));
    TRY(dg1.upload(grid1, n1 * 8));
    TRY(dg2.upload(grid2, n2 * 8));
    TRY(dG.upload(Kgrid, (size_t)numBands * n1 * n2 * 8));
    TRY(dl_bilininterp_precomputedbins_dev(numBands, nobj, dk.as<double>(), nobj, dv1.as<double>(),
                                           dv2.as<double>(), dp1.as<long long>(),
                                           dp2.as<long long>(), dg1.as<double>(), dg2.as<double>(),
                                           dG.as<double>(), n1, n2, nullptr));
    DL_CUDA_CHECK(cudaMemcpy(Kinterp, dk.p, (size_t)numBands * nobj * 8, cudaMemcpyDeviceToHost));
    return DL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ bookkeeping / measurement
#include <atomic>
#include <mutex>

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_timing{0};
static std::mutex g_tmutex;
static std::vector<cudaEvent_t> g_events;     // pairs (before, after) around the fused kernel
static size_t g_events_used = 0;

void dl_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool dl_timing_enabled() { return g_timing.load(std::memory_order_relaxed) != 0; }

void dl_timing_mark(cudaStream_t stream, int which) {
    std::lock_guard<std::mutex> lk(g_tmutex);
    if (which == 0 && (g_events_used & 1)) return;       // unmatched: ignore
    if (g_events_used >= g_events.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        g_events.push_back(e);
    }
    cudaEventRecord(g_events[g_events_used++], stream);
}

namespace {
// 8 independent FMA chains per thread: measures the FP64 vector-pipe peak of the device
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* out) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4,
           a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;     // never true; keeps the chains alive
}
__global__ void selftest_math_kernel(long long n, const double* x, double* out) {
    __shared__ double etab[dl::EXP_TAB];
    dl::exp_tab_init(etab, threadIdx.x, blockDim.x);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    out[i] = dl::rcp_fast(v);
    out[n + i] = dl::rsqrt_fast(v);
    out[2 * n + i] = dl::exp_nonpos_tab(-v, etab);
    double s0, s1;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s0) : "d"(v));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s1) : "d"(v));
    out[3 * n + i] = s0;
    out[4 * n + i] = s1;
}
}  // namespace

extern "C" {

long long dl_launch_count(void) { return g_launches.load(); }

// accuracy self-test of the fast FP64 helpers used by the fused kernel (device pointers):
// out[0][i] = rcp_fast(x[i]), out[1][i] = rsqrt_fast(x[i]), out[2][i] = exp_nonpos_tab(-x[i]),
// out[3][i] = raw MUFU reciprocal seed, out[4][i] = raw MUFU rsqrt seed
int dl_selftest_math_dev(long long n, const double* x, double* out, void* stream) {
    DL_REQUIRE(n > 0 && x && out, "bad arguments");
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, out);
    DL_CUDA_CHECK(cudaGetLastError());
    return DL_OK;
}

// enable/disable CUDA-event timing of the dominant (fused likelihood) kernel; resets the log
int dl_timing_enable(int on) {
    std::lock_guard<std::mutex> lk(g_tmutex);
    g_timing.store(on ? 1 : 0);
    g_events_used = 0;
    return DL_OK;
}

// synchronises, returns the summed duration (ms) and number of fused-kernel launches recorded
// since dl_timing_enable(1), and resets the log
int dl_timing_collect(double* total_ms, int* launches) {
    std::lock_guard<std::mutex> lk(g_tmutex);
    double tot = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < g_events_used; i += 2) {
        DL_CUDA_CHECK(cudaEventSynchronize(g_events[i + 1]));
        float ms = 0.f;
        DL_CUDA_CHECK(cudaEventElapsedTime(&ms, g_events[i], g_events[i + 1]));
        tot += ms;
        ++n;
    }
    g_events_used = 0;
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return DL_OK;
}

// measured FP64 FMA peak of the current device in TFLOP/s (dependent-chain FMA kernel filling
// every SM, best of `reps`)
int dl_fp64_peak(int reps, double* tflops) {
    int dev = 0, sms = 0;
    DL_CUDA_CHECK(cudaGetDevice(&dev));
    DL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    DL_CUDA_CHECK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    DL_CUDA_CHECK(cudaEventCreate(&e0));
    DL_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 8;
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        DL_CUDA_CHECK(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, 256>>>(iters, 1.0, out);
        DL_CUDA_CHECK(cudaEventRecord(e1));
        DL_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        DL_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
        if (r > 0 && flops / (ms * 1e-3) * 1e-12 > best) best = flops / (ms * 1e-3) * 1e-12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (tflops) *tflops = best;
    return DL_OK;
}

}  // extern "C"

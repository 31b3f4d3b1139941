// Non-fused likelihood: the drop-in for delight/utils_cy.pyx:83-128 (one target per call, the
// full like[nz][nt] array is returned in the reference's layout).  HBM-bound by construction:
// 2*nf*8 B read + 8 B written per cell (+ 16 B for the exp pass): one thread per (z, t) cell,
// nf contiguous doubles per thread, cells contiguous across the warp.
#include "common.cuh"

namespace dl {

__device__ __forceinline__ unsigned long long dkey(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// log-likelihood per cell (utils_cy.pyx:98-118) + global max (:120-124) via ordered atomicMax.
// The formulas are the reference's, evaluated with one IEEE reciprocal per band and pass (the
// literal code divides three times by the same variance) and ONE log per cell: the product of
// the variances is carried as mantissa x 2^exponent (renormalised every four bands, so any nf
// and any unit system stay in range) instead of nf + 2 logs.  That keeps the kernel on the HBM
// roofline (2 nf 8 B read + 8 B written per cell) instead of the FP64 pipe.
// 1 / var: MUFU seed + one third-order Newton step (common.cuh) for normal positive finite
// variances -- what a catalogue produces --, the IEEE intrinsic for anything else (zero,
// denormal, negative, inf, NaN keep the reference's arithmetic semantics).
__device__ __forceinline__ double rcp_var(double v) {
    const unsigned hi = (unsigned)__double2hiint(v);
    return (hi - 0x00100000u < 0x7fe00000u) ? rcp_fast(v) : __drcp_rn(v);
}

// prod = m x 2^e with m in [1, 2): strip the exponent of a positive NORMAL double through its
// bit pattern (4 integer instructions instead of the ilogb / scalbn calls); anything else
// (zero, denormal, inf, NaN) is left alone and reaches the log as it is.
__device__ __forceinline__ void strip_exponent(double& prod, int& pexp) {
    const int hi = __double2hiint(prod);
    const unsigned ex = ((unsigned)hi >> 20) & 0x7ffu;
    if (hi > 0 && ex - 1u < 0x7feu) {
        pexp += (int)ex - 1023;
        prod = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(prod));
    }
}

// VEC: nf even and 16-byte aligned grids -> the nf doubles of a cell are read as double2
// (half the load instructions and L1 wavefronts of the strided per-thread access).
template <bool VEC>
__global__ void __launch_bounds__(256) loglike_kernel(
    long long ncell, int nt, int nf, const double* __restrict__ f_obs,
    const double* __restrict__ f_obs_var, const double* __restrict__ f_mod,
    const double* __restrict__ f_mod_covar, const double* __restrict__ ell_hat,
    const double* __restrict__ ell_var, double* __restrict__ like,
    unsigned long long* __restrict__ maxkey) {
    extern __shared__ double sh[];                 // f_obs[nf], f_obs_var[nf], f_obs^2[nf]
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        sh[i] = f_obs[i]; sh[nf + i] = f_obs_var[i]; sh[2 * nf + i] = f_obs[i] * f_obs[i];
    }
    __syncthreads();
    const double LN2 = 0.693147180559945309417, LOG_2PI = 1.83787706640934548356;
    double best = -INFINITY;
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell;
         cell += (long long)gridDim.x * blockDim.x) {
        // (64-bit division is ~100 instructions: use the 32-bit one whenever the cell count allows)
        const long long iz = ncell < 0x7fffffffll ? (long long)((unsigned)cell / (unsigned)nt)
                                                  : cell / nt;
        const double eh = ell_hat[iz], ev = ell_var[iz];
        const double iev = rcp_var(ev);
        const double* m = f_mod + cell * nf;
        const double* c = f_mod_covar + cell * nf;
        double ell2 = 0.0, fot = 0.0, ftt = 0.0, foo = 0.0, prod = 1.0, rftt = 0.0;
        int pexp = 0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            fot = eh * iev; ftt = iev; foo = eh * eh * iev; prod = 1.0; pexp = 0;
            auto band = [&](int k, double mk, double ck) {
                const double var = fma(ell2, ck, sh[nf + k]);            // :107
                const double r = rcp_var(var);
                const double mr = mk * r;
                fot = fma(mr, sh[k], fot);                               // :108
                ftt = fma(mr, mk, ftt);                                  // :109
                foo = fma(sh[2 * nf + k], r, foo);                       // :110
                if (pass == 1) {                                         // :111-112
                    prod *= var;
                    if ((k & 3) == 3) strip_exponent(prod, pexp);
                }
            };
            if (VEC) {
                const double2* m2 = reinterpret_cast<const double2*>(m);
                const double2* c2 = reinterpret_cast<const double2*>(c);
                for (int k = 0; k < nf; k += 2) {
                    const double2 mm = m2[k >> 1], cc = c2[k >> 1];
                    band(k, mm.x, cc.x);
                    band(k + 1, mm.y, cc.y);
                }
            } else {
                for (int k = 0; k < nf; ++k) band(k, m[k], c[k]);
            }
            rftt = rcp_var(ftt);
            const double ell = fot * rftt;                               // :113
            ell2 = ell * ell;
        }
        const double chi2 = fma(-(fot * rftt), fot, foo);                // :115
        // logDenom = sum log(2 pi var) + log(2 pi ell_var) + log(FTT / 2 pi)     (:111-117)
        double logden;
        {
            strip_exponent(prod, pexp);
            double tm = ev * ftt;                                        // > 0
            strip_exponent(tm, pexp);
            logden = log(prod * tm) + pexp * LN2 + nf * LOG_2PI;
        }
        const double L = -0.5 * chi2 - 0.5 * logden;                     // :118
        like[cell] = L;
        best = fmax(best, L);
    }
    best = warp_max(best);
    __shared__ double wbest[8];
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = fmax(best, wbest[w]);
        atomicMax(maxkey, dkey(best));
    }
}

// like = exp(like - max)  (utils_cy.pyx:125-128)
__global__ void __launch_bounds__(256) expnorm_kernel(long long ncell, double* __restrict__ like,
                                                      const unsigned long long* __restrict__ maxkey) {
    const double top = dkey_inv(*maxkey);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell;
         i += (long long)gridDim.x * blockDim.x)
        like[i] = exp(like[i] - top);
}

}  // namespace dl

using namespace dl;

extern "C" {

// Device-pointer form of approx_flux_likelihood_cy.  `scratch` = 8 bytes of device memory.
int dl_approx_flux_likelihood_dev(double* like, long long nz, long long nt, long long nf,
                                  const double* f_obs, const double* f_obs_var,
                                  const double* f_mod, const double* f_mod_covar,
                                  const double* ell_hat, const double* ell_var, void* scratch,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DL_REQUIRE(nz >= 0 && nt >= 0 && nf >= 0, "negative size");
    DL_REQUIRE(nf <= 1024, "nf too large");
    const long long ncell = nz * nt;
    if (ncell == 0) return DL_OK;
    DL_REQUIRE(like && f_obs && f_obs_var && f_mod && f_mod_covar && ell_hat && ell_var && scratch,
               "null pointer");
    DL_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 8, stream));
    int dev = 0, sms = 148;
    DL_CUDA_CHECK(cudaGetDevice(&dev));
    DL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (ncell + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;   // grid-stride, 16 CTAs/SM
    const bool vec = (nf % 2 == 0) && (((uintptr_t)f_mod | (uintptr_t)f_mod_covar) & 15) == 0;
    auto kern = vec ? loglike_kernel<true> : loglike_kernel<false>;
    kern<<<(unsigned)blocks, 256, 3 * nf * sizeof(double), stream>>>(
        ncell, (int)nt, (int)nf, f_obs, f_obs_var, f_mod, f_mod_covar, ell_hat, ell_var, like,
        (unsigned long long*)scratch);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    expnorm_kernel<<<(unsigned)blocks, 256, 0, stream>>>(ncell, like,
                                                         (const unsigned long long*)scratch);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

}  // extern "C"

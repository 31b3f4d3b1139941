// Second ("compressed") pass of scripts/delight-apply.py:109-135: per target only the K training
// galaxies stored by the first pass enter the likelihood, which is the NumPy twin
// delight/utils.py:115-176 (marginalizeEll, renormalize; `normalized` selects the log-determinant
// terms, the script passes False).  K*nz cells per target with no sharing between targets, so
// this is a gather: one CTA per target, thread <-> redshift (coalesced rows of the fused grid
// layout [nT][2 nB][nzp]), literal formulas, two sweeps (max, then exp/accumulate) instead of a
// K*nz scratch array.  Cost at config 3: 7.5e8 cells, HBM/L2-bound on the gathered rows.
#include "common.cuh"

namespace dl {

constexpr int CP_THREADS = 256;
constexpr int CP_MAXNF = 64;

struct CompParams {
    const double* grid;     // [nT][2*nB][nzp]
    const double* prior;    // [nT][nzp]
    const double* obs;      // [nO][nf]
    const double* var;      // [nO][nf]
    const long long* sel;   // [nO][K]  training indices, < 0 = unused entry
    const double* zgrid;    // [nz]
    const double* ell_hat_z;  // [nz] or null
    const double* ell_var_z;  // [nz] or null
    double* pdf;            // [nO][ld_pdf]
    double* evid;           // [nO][K] or null
    long long ld_pdf;
    double ell_hat, ell_var, cov_scale;
    int bands[CP_MAXNF];
    int nB, nz, nzp, nT, nf, K, normalized;
};

// log-likelihood of one (target, training galaxy, z) cell: utils.py:141-174
__device__ __forceinline__ double twin_loglike(const CompParams& p, const double* sobs,
                                               const double* svar, const double* gt, int z,
                                               double eh, double ev) {
    double ell = 0.0, fot = 0.0, ftt = 0.0, foo = 0.0, logdet = 0.0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fot = eh / ev; ftt = 1.0 / ev; foo = eh * eh / ev; logdet = 0.0;
        for (int f = 0; f < p.nf; ++f) {
            const double m = gt[(size_t)p.bands[f] * p.nzp + z];
            const double c = gt[(size_t)(p.nB + p.bands[f]) * p.nzp + z] * p.cov_scale;
            const double v = svar[f] + ell * ell * c;
            fot += m * sobs[f] / v;
            ftt += m * m / v;
            foo += sobs[f] * sobs[f] / v;
            if (pass == 1 && p.normalized) logdet += log(v);
        }
        ell = fot / ftt;
    }
    const double chi2 = foo - fot * fot / ftt;
    double logden = log(ftt) - log(DL_TWO_PI);
    if (p.normalized) logden += logdet + p.nf * log(DL_TWO_PI) + log(DL_TWO_PI * ev);
    return -0.5 * chi2 - 0.5 * logden;
}

__global__ void __launch_bounds__(CP_THREADS) compressed_kernel(const CompParams p) {
    extern __shared__ double sm[];
    double* sobs = sm;                    // [nf]
    double* svar = sobs + p.nf;           // [nf]
    double* spdf = svar + p.nf;           // [nz]
    __shared__ double red[CP_THREADS / 32];
    __shared__ double bcast;
    const int o = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.nf; i += CP_THREADS) {
        sobs[i] = p.obs[(size_t)o * p.nf + i];
        svar[i] = p.var[(size_t)o * p.nf + i];
    }
    for (int z = tid; z < p.nz; z += CP_THREADS) spdf[z] = 0.0;
    __syncthreads();
    const long long* sel = p.sel + (size_t)o * p.K;

    // sweep 1: max log-likelihood over the selected (training galaxy, z) cells (utils.py:175)
    double best = -INFINITY;
    for (int k = 0; k < p.K; ++k) {
        const long long t = sel[k];
        if (t < 0 || t >= p.nT) continue;
        const double* gt = p.grid + (size_t)t * 2 * p.nB * p.nzp;
        for (int z = tid; z < p.nz; z += CP_THREADS) {
            const double eh = p.ell_hat_z ? p.ell_hat_z[z] : p.ell_hat;
            const double ev = p.ell_var_z ? p.ell_var_z[z] : p.ell_var;
            best = fmax(best, twin_loglike(p, sobs, svar, gt, z, eh, ev));
        }
    }
    best = warp_max(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < CP_THREADS / 32; ++w) best = fmax(best, red[w]);
        bcast = best;
    }
    __syncthreads();
    const double top = bcast;
    __syncthreads();

    // sweep 2: exp(L - max) * prior, pdf = sum over galaxies, evidence = trapezoid over z
    // (delight-apply.py:135,145-146)
    for (int k = 0; k < p.K; ++k) {
        const long long t = (sel[k] >= 0 && sel[k] < p.nT) ? sel[k] : -1;
        double e = 0.0;
        if (t >= 0) {
            const double* gt = p.grid + (size_t)t * 2 * p.nB * p.nzp;
            const double* pr = p.prior + (size_t)t * p.nzp;
            for (int z = tid; z < p.nz; z += CP_THREADS) {
                const double eh = p.ell_hat_z ? p.ell_hat_z[z] : p.ell_hat;
                const double ev = p.ell_var_z ? p.ell_var_z[z] : p.ell_var;
                const double v = exp(twin_loglike(p, sobs, svar, gt, z, eh, ev) - top) * pr[z];
                spdf[z] += v;                 // z is owned by this thread
                const double zl = p.zgrid[z > 0 ? z - 1 : 0];
                const double zr = p.zgrid[z + 1 < p.nz ? z + 1 : p.nz - 1];
                e = fma(v, 0.5 * (zr - zl), e);
            }
        }
        if (p.evid) {
            e = warp_sum(e);
            if (lane == 0) red[warp] = e;
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < CP_THREADS / 32; ++w) e += red[w];
                p.evid[(size_t)o * p.K + k] = t >= 0 ? e : 0.0;
            }
            __syncthreads();
        }
    }
    for (int z = tid; z < p.nz; z += CP_THREADS) p.pdf[(size_t)o * p.ld_pdf + z] = spdf[z];
}

}  // namespace dl

using namespace dl;

extern "C" {

long long dl_fused_padded_nz(long long nz);

int dl_like_compressed_dev(long long nO, long long nT, long long nz, long long nB, int nf,
                           const int* bands, const double* obs, const double* var,
                           const double* grid, const double* prior, const double* zgrid,
                           double ell_hat, double ell_var, const double* ell_hat_z,
                           const double* ell_var_z, double cov_scale, int K,
                           const long long* sel, int normalized, double* pdf, long long ld_pdf,
                           double* evid, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DL_REQUIRE(nO >= 0 && nT > 0 && nz > 1 && nB > 0, "bad shape");
    DL_REQUIRE(nf >= 1 && nf <= CP_MAXNF && nf <= nB, "nf out of range");
    DL_REQUIRE(K >= 1, "K must be positive");
    DL_REQUIRE(ld_pdf >= nz, "ld_pdf < nz");
    DL_REQUIRE((ell_hat_z == nullptr) == (ell_var_z == nullptr), "ell_hat_z / ell_var_z: both or none");
    DL_REQUIRE(ell_hat_z || ell_var > 0.0, "ell_var must be positive");
    DL_REQUIRE(bands && obs && var && grid && prior && zgrid && sel && pdf, "null pointer");
    if (nO == 0) return DL_OK;
    CompParams p;
    for (int f = 0; f < nf; ++f) {
        DL_REQUIRE(bands[f] >= 0 && bands[f] < nB, "band index out of range");
        p.bands[f] = bands[f];
    }
    p.grid = grid; p.prior = prior; p.obs = obs; p.var = var; p.sel = sel; p.zgrid = zgrid;
    p.ell_hat_z = ell_hat_z; p.ell_var_z = ell_var_z; p.pdf = pdf; p.evid = evid;
    p.ld_pdf = ld_pdf; p.ell_hat = ell_hat; p.ell_var = ell_var; p.cov_scale = cov_scale;
    p.nB = (int)nB; p.nz = (int)nz; p.nzp = (int)dl_fused_padded_nz(nz); p.nT = (int)nT;
    p.nf = nf; p.K = K; p.normalized = normalized;
    const size_t smem = (2 * (size_t)nf + (size_t)nz) * sizeof(double);
    DL_REQUIRE(smem <= 200 * 1024, "nz too large for the compressed pass");
    if (smem > 48 * 1024)
        DL_CUDA_CHECK(cudaFuncSetAttribute(compressed_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    compressed_kernel<<<(unsigned)nO, CP_THREADS, smem, stream>>>(p);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

}  // extern "C"

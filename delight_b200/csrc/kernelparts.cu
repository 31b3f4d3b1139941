// Physical GP kernel from the filters' Gaussian-mixture coefficients (subsystem 2).
//
// Device counterparts of delight/photoz_kernels_cy.pyx:
//     kernelparts (:103-177), kernelparts_diag (:38-100), kernel_parts_interp (:9-35)
// and delight/utils_cy.pyx: find_positions (:10-22); plus a one-launch builder for the
// (nB, nB, nzGP, nzGP) grids of Photoz_kernel.construct_interpolators (photoz_kernels.py:342-416).
//
// One thread per (o1, o2) pair.  Continuum term KC: NC^2 Gaussians in (o1, o2) -- not separable
// because sigma mixes both redshifts.  Line term KL: the exponent of pyx:160 is a sum of a term
// in (o1, i, l1), a term in (o2, j, l2) and a term in (l1, l2) only, so
//     KL = sum_{l2<=l1} w(l1,l2) exp(-((mu_l1-mu_l2)/alpha_L)^2/2) A1[l1] A2[l2],
//     A1[l] = sum_i amp1_i exp(-((mu1_i - x1 mu_l)/sig1_i)^2/2)   (same for A2 with band 2)
// which needs 2*NL*NC exponentials per pair instead of NC^2*NL*(NL+1)/2.
#include "common.cuh"

namespace dl {

constexpr int KP_MAXNL = 8;

struct KPArgs {
    int NO1, NO2, NC, NL;
    double alpha_C, alpha_L;
    const double* amp;       // [nB][NC]
    const double* mu;
    const double* sig;
    const double* lines_mu;  // [NL]
    const double* norms;     // [nB]
};

__device__ __forceinline__ void line_sums(const KPArgs& a, int b, double x, double* A) {
    for (int l = 0; l < a.NL; ++l) A[l] = 0.0;
    for (int i = 0; i < a.NC; ++i) {
        const double am = a.amp[b * a.NC + i], m = a.mu[b * a.NC + i];
        const double is = 1.0 / a.sig[b * a.NC + i];
        for (int l = 0; l < a.NL; ++l) {
            const double e = (m - x * a.lines_mu[l]) * is;
            A[l] += am * exp(-0.5 * e * e);
        }
    }
}

// returns un-normalised sums for one pair
template <bool DIAG>
__device__ __forceinline__ void pair_sums(const KPArgs& a, int b1, double x1, int b2, double x2,
                                          bool grad, double& kc, double& kl, double& dc,
                                          double& dl_, double& dz) {
    kc = kl = dc = dl_ = dz = 0.0;
    const double aC = a.alpha_C;
    const double q3 = x1 * x2 * aC;
    for (int i = 0; i < a.NC; ++i) {
        const double am1 = a.amp[b1 * a.NC + i], mu1 = a.mu[b1 * a.NC + i],
                     sg1 = a.sig[b1 * a.NC + i];
        const double q2 = x2 * sg1;
        for (int j = 0; j < a.NC; ++j) {
            const double am2 = a.amp[b2 * a.NC + j], mu2 = a.mu[b2 * a.NC + j],
                         sg2 = a.sig[b2 * a.NC + j];
            const double q1 = x1 * sg2;
            const double s2 = q1 * q1 + q2 * q2 + q3 * q3;          // pyx:147
            const double is2 = 1.0 / s2;
            const double is = sqrt(is2);
            const double d = x1 * mu2 - x2 * mu1;
            const double g = am1 * am2 * DL_TWO_PI * sg1 * sg2 * exp(-0.5 * d * d * is2) * is;
            kc += aC * g;                                           // :149
            if (grad) {
                const double u = aC * d * x1 * x2;
                dc += g * (1.0 - q3 * q3 * is2 + u * u * is2 * is2);                  // :151
                if (!DIAG)
                    dz += aC * g * ((sg2 * sg2 * x1 + x1 * x2 * x2 * aC * aC) *
                                    (d * d * is2 * is2 - is2) - mu2 * d * is2);        // :152-153
            }
        }
    }
    if (a.NL > 0) {
        double A1[KP_MAXNL], A2[KP_MAXNL];
        line_sums(a, b1, x1, A1);
        line_sums(a, b2, x2, A2);
        const double iaL = 1.0 / a.alpha_L;
        for (int l1 = 0; l1 < a.NL; ++l1) {
            for (int l2 = 0; l2 <= l1; ++l2) {                       // :156-169
                const double dm = a.lines_mu[l1] - a.lines_mu[l2];
                const double e3 = dm * iaL;
                const double w = (l2 < l1 ? 2.0 : 1.0) * exp(-0.5 * e3 * e3) * A1[l1] * A2[l2];
                kl += w;
                if (grad) dl_ += w * dm * dm * iaL * iaL * iaL;
            }
        }
    }
}

// kernelparts: accumulate into the outputs, then divide the whole element by the norms
// (photoz_kernels_cy.pyx:149-177).  Outputs have leading dimension ld (elements).
__global__ void __launch_bounds__(128) kernelparts_kernel(
    KPArgs a, const long long* __restrict__ b1, const double* __restrict__ fz1,
    const long long* __restrict__ b2, const double* __restrict__ fz2, int grad,
    double* KL, double* KC, double* DaC, double* DaL, double* Daz, long long ld) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.NO1 * a.NO2) return;
    const int o1 = (int)(idx / a.NO2), o2 = (int)(idx % a.NO2);
    const int bb1 = (int)b1[o1], bb2 = (int)b2[o2];
    double kc, kl, dc, dl_, dz;
    pair_sums<false>(a, bb1, fz1[o1], bb2, fz2[o2], grad != 0, kc, kl, dc, dl_, dz);
    const double nn = a.norms[bb1] * a.norms[bb2];
    const long long k = (long long)o1 * ld + o2;
    KC[k] = (KC[k] + kc) / nn;
    KL[k] = (KL[k] + kl) / nn;
    if (grad) {
        DaC[k] = (DaC[k] + dc) / nn;
        DaL[k] = (DaL[k] + dl_) / nn;
        Daz[k] = (Daz[k] + dz) / nn;
    }
}

// kernelparts_diag: KC/KL overwritten (:61-62), gradients accumulate (:71,83,90), no D_alpha_z
__global__ void __launch_bounds__(128) kernelparts_diag_kernel(
    KPArgs a, const long long* __restrict__ b1, const double* __restrict__ fz1, int grad,
    double* KL, double* KC, double* DaC, double* DaL) {
    const int o1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (o1 >= a.NO1) return;
    const int bb = (int)b1[o1];
    double kc, kl, dc, dl_, dz;
    pair_sums<true>(a, bb, fz1[o1], bb, fz1[o1], grad != 0, kc, kl, dc, dl_, dz);
    const double nn = a.norms[bb] * a.norms[bb];
    KC[o1] = kc / nn;
    KL[o1] = kl / nn;
    if (grad) {
        DaC[o1] = (DaC[o1] + dc) / nn;
        DaL[o1] = (DaL[o1] + dl_) / nn;
    }
}

// All band pairs of the interpolation grids in one launch:
//   KCg, KLg [nB][nB][nzg][nzg] on fz = 1 + redshiftGridGP   (photoz_kernels.py:350-370)
//   KCd, KLd [nB][nzg] diagonals                              (photoz_kernels.py:378-416)
__global__ void __launch_bounds__(128) kernel_grid_kernel(KPArgs a, int nB, int nzg,
                                                          const double* __restrict__ fz,
                                                          double* __restrict__ KCg,
                                                          double* __restrict__ KLg,
                                                          double* __restrict__ KCd,
                                                          double* __restrict__ KLd) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npair = (long long)nB * nB * nzg * nzg;
    double kc, kl, dc, dl_, dz;
    if (idx < npair) {
        const int z2 = (int)(idx % nzg);
        const int z1 = (int)((idx / nzg) % nzg);
        const int bb2 = (int)((idx / ((long long)nzg * nzg)) % nB);
        const int bb1 = (int)(idx / ((long long)nzg * nzg * nB));
        pair_sums<false>(a, bb1, fz[z1], bb2, fz[z2], false, kc, kl, dc, dl_, dz);
        const double nn = a.norms[bb1] * a.norms[bb2];
        KCg[idx] = kc / nn;
        KLg[idx] = kl / nn;
    } else if (idx < npair + (long long)nB * nzg) {
        const long long d = idx - npair;
        const int bb = (int)(d / nzg), z1 = (int)(d % nzg);
        pair_sums<true>(a, bb, fz[z1], bb, fz[z1], false, kc, kl, dc, dl_, dz);
        const double nn = a.norms[bb] * a.norms[bb];
        KCd[d] = kc / nn;
        KLd[d] = kl / nn;
    }
}

// find_positions (utils_cy.pyx:10-22): first bin p with grid[p] <= v <= grid[p+1]; values outside
// the grid leave p1s untouched.  Binary search on the sorted grid, then step back over equal
// nodes so the FIRST matching bin is returned like the reference's linear scan.
__global__ void find_positions_kernel(int NO1, int nz, const double* __restrict__ fz1,
                                      long long* __restrict__ p1s,
                                      const double* __restrict__ grid) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= NO1 || nz < 2) return;
    const double v = fz1[o];
    if (!(v >= grid[0] && v <= grid[nz - 1])) return;
    int lo = 0, hi = nz - 1;                  // invariant: grid[lo] <= v <= grid[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (grid[mid] < v) lo = mid; else hi = mid;   // smallest hi with grid[hi] >= v
    }
    // lo = largest node strictly below v (or 0): grid[lo] <= v <= grid[lo+1], and no earlier
    // bin can contain v
    const int p = lo;
    p1s[o] = p;
}

// kernel_parts_interp (photoz_kernels_cy.pyx:9-35)
__global__ void kernel_parts_interp_kernel(int NO1, int NO2, double* __restrict__ Kinterp,
                                           long long ld, const long long* __restrict__ b1,
                                           const double* __restrict__ fz1,
                                           const long long* __restrict__ p1s,
                                           const long long* __restrict__ b2,
                                           const double* __restrict__ fz2,
                                           const long long* __restrict__ p2s,
                                           const double* __restrict__ fzGrid,
                                           const double* __restrict__ Kgrid, int nB, int nzg) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)NO1 * NO2) return;
    const int o1 = (int)(idx / NO2), o2 = (int)(idx % NO2);
    const long long p1 = p1s[o1], p2 = p2s[o2];
    const double x = fz1[o1], xl = fzGrid[p1], xh = fzGrid[p1 + 1];
    const double y = fz2[o2], yl = fzGrid[p2], yh = fzGrid[p2 + 1];
    const double* G = Kgrid + ((b1[o1] * nB + b2[o2]) * (long long)nzg) * nzg;
    const double inv = 1.0 / (xh - xl) / (yh - yl);
    Kinterp[(long long)o1 * ld + o2] = inv * ((xh - x) * (yh - y) * G[p1 * nzg + p2] +
                                              (x - xl) * (yh - y) * G[(p1 + 1) * nzg + p2] +
                                              (xh - x) * (y - yl) * G[p1 * nzg + p2 + 1] +
                                              (x - xl) * (y - yl) * G[(p1 + 1) * nzg + p2 + 1]);
}

// bilininterp_precomputedbins (utils_cy.pyx:25-50): bilinear interpolation of nbands grids
// Kgrid[b][n1][n2] at the points (v1s[o], v2s[o]) with precomputed bins -> Kinterp[b][o]
__global__ void bilininterp_kernel(int numBands, int nobj, double* __restrict__ Kinterp,
                                   long long ld, const double* __restrict__ v1s,
                                   const double* __restrict__ v2s,
                                   const long long* __restrict__ p1s,
                                   const long long* __restrict__ p2s,
                                   const double* __restrict__ grid1,
                                   const double* __restrict__ grid2,
                                   const double* __restrict__ Kgrid, int n1, int n2) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)numBands * nobj) return;
    const int b = (int)(idx / nobj), o = (int)(idx % nobj);
    const long long p1 = p1s[o], p2 = p2s[o];
    const double v1 = v1s[o], v2 = v2s[o];
    const double dzm2 = 1. / (grid1[p1 + 1] - grid1[p1]) / (grid2[p2 + 1] - grid2[p2]);
    const double* G = Kgrid + (long long)b * n1 * n2;
    Kinterp[(long long)b * ld + o] = dzm2 * (
        (grid1[p1 + 1] - v1) * (grid2[p2 + 1] - v2) * G[p1 * n2 + p2]
        + (v1 - grid1[p1]) * (grid2[p2 + 1] - v2) * G[(p1 + 1) * n2 + p2]
        + (grid1[p1 + 1] - v1) * (v2 - grid2[p2]) * G[p1 * n2 + p2 + 1]
        + (v1 - grid1[p1]) * (v2 - grid2[p2]) * G[(p1 + 1) * n2 + p2 + 1]);
}

}  // namespace dl

using namespace dl;

extern "C" {

int dl_bilininterp_precomputedbins_dev(long long numBands, long long nobj, double* Kinterp,
                                       long long ld, const double* v1s, const double* v2s,
                                       const long long* p1s, const long long* p2s,
                                       const double* grid1, const double* grid2,
                                       const double* Kgrid, long long n1, long long n2,
                                       void* stream) {
    DL_REQUIRE(numBands >= 0 && nobj >= 0 && n1 >= 2 && n2 >= 2 && ld >= nobj, "bad sizes");
    if (numBands * nobj == 0) return DL_OK;
    const long long n = numBands * nobj;
    bilininterp_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (int)numBands, (int)nobj, Kinterp, ld, v1s, v2s, p1s, p2s, grid1, grid2, Kgrid, (int)n1,
        (int)n2);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

static int kp_args(KPArgs& a, long long NO1, long long NO2, long long NC, long long NL,
                   double alpha_C, double alpha_L, const double* amp, const double* mu,
                   const double* sig, const double* lines_mu, const double* norms) {
    DL_REQUIRE(NO1 >= 0 && NO2 >= 0 && NC >= 1, "bad sizes");
    DL_REQUIRE(NL >= 0 && NL <= KP_MAXNL, "0 <= NL <= 8 supported");
    DL_REQUIRE(amp && mu && sig && norms && (NL == 0 || lines_mu), "null pointer");
    a.NO1 = (int)NO1; a.NO2 = (int)NO2; a.NC = (int)NC; a.NL = (int)NL;
    a.alpha_C = alpha_C; a.alpha_L = alpha_L;
    a.amp = amp; a.mu = mu; a.sig = sig; a.lines_mu = lines_mu; a.norms = norms;
    return DL_OK;
}

int dl_kernelparts_dev(long long NO1, long long NO2, long long NC, long long NL, double alpha_C,
                       double alpha_L, const double* amp, const double* mu, const double* sig,
                       const double* lines_mu, const double* norms, const long long* b1,
                       const double* fz1, const long long* b2, const double* fz2,
                       int grad_needed, double* KL, double* KC, double* D_alpha_C,
                       double* D_alpha_L, double* D_alpha_z, long long ld, void* stream) {
    KPArgs a;
    int rc = kp_args(a, NO1, NO2, NC, NL, alpha_C, alpha_L, amp, mu, sig, lines_mu, norms);
    if (rc) return rc;
    if (NO1 * NO2 == 0) return DL_OK;
    DL_REQUIRE(ld >= NO2, "ld < NO2");
    const long long n = NO1 * NO2;
    kernelparts_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        a, b1, fz1, b2, fz2, grad_needed, KL, KC, D_alpha_C, D_alpha_L, D_alpha_z, ld);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

int dl_kernelparts_diag_dev(long long NO1, long long NC, long long NL, double alpha_C,
                            double alpha_L, const double* amp, const double* mu,
                            const double* sig, const double* lines_mu, const double* norms,
                            const long long* b1, const double* fz1, int grad_needed, double* KL,
                            double* KC, double* D_alpha_C, double* D_alpha_L, void* stream) {
    KPArgs a;
    int rc = kp_args(a, NO1, 1, NC, NL, alpha_C, alpha_L, amp, mu, sig, lines_mu, norms);
    if (rc) return rc;
    if (NO1 == 0) return DL_OK;
    kernelparts_diag_kernel<<<(unsigned)((NO1 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        a, b1, fz1, grad_needed, KL, KC, D_alpha_C, D_alpha_L);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

int dl_kernel_grids_dev(long long nB, long long nzg, long long NC, long long NL, double alpha_C,
                        double alpha_L, const double* amp, const double* mu, const double* sig,
                        const double* lines_mu, const double* norms, const double* fzGrid,
                        double* KC_grid, double* KL_grid, double* KC_diag, double* KL_diag,
                        void* stream) {
    KPArgs a;
    int rc = kp_args(a, nzg, nzg, NC, NL, alpha_C, alpha_L, amp, mu, sig, lines_mu, norms);
    if (rc) return rc;
    DL_REQUIRE(nB >= 1 && nzg >= 2, "bad sizes");
    const long long n = nB * nB * nzg * nzg + nB * nzg;
    kernel_grid_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        a, (int)nB, (int)nzg, fzGrid, KC_grid, KL_grid, KC_diag, KL_diag);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

int dl_find_positions_dev(long long NO1, long long nz, const double* fz1, long long* p1s,
                          const double* fzGrid, void* stream) {
    DL_REQUIRE(NO1 >= 0 && nz >= 0, "bad sizes");
    if (NO1 == 0 || nz < 2) return DL_OK;
    find_positions_kernel<<<(unsigned)((NO1 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (int)NO1, (int)nz, fz1, p1s, fzGrid);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

int dl_kernel_parts_interp_dev(long long NO1, long long NO2, double* Kinterp, long long ld,
                               const long long* b1, const double* fz1, const long long* p1s,
                               const long long* b2, const double* fz2, const long long* p2s,
                               const double* fzGrid, const double* Kgrid, long long nB,
                               long long nzg, void* stream) {
    DL_REQUIRE(NO1 >= 0 && NO2 >= 0 && nB >= 1 && nzg >= 2 && ld >= NO2, "bad sizes");
    if (NO1 * NO2 == 0) return DL_OK;
    const long long n = NO1 * NO2;
    kernel_parts_interp_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (int)NO1, (int)NO2, Kinterp, ld, b1, fz1, p1s, b2, fz2, p2s, fzGrid, Kgrid, (int)nB,
        (int)nzg);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

}  // extern "C"

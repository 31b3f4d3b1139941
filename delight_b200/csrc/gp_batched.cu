// Batched per-training-galaxy Gaussian process (subsystem 3).
//
//   dl_gp_fit_dev      replaces the per-galaxy body of scripts/delight-learn.py:61-90:
//                      template fluxes at z (delight/io.py:263-277 interpolators), best template
//                      and luminosity scale (delight/utils.py:179-193), K(X) + diag(Yvar) from
//                      the interpolated kernel grids (delight/photoz_kernels.py:221-231,
//                      277-340), lower Cholesky and residuals (delight/photoz_gp.py:186-219),
//                      serialised exactly like getCore + the script's row (photoz_gp.py:221-234,
//                      delight-learn.py:85-90).  One warp per galaxy, lane j <-> band row j.
//   dl_gp_predict_dev  replaces scripts/delight-apply.py:89-98 per training galaxy:
//                      setCore (photoz_gp.py:236-254) + predictAndInterpolate (:263-347),
//                      writing model_mean / model_var straight into the fused likelihood's
//                      HBM layout [nT][2 nB][nzp].  One CTA per galaxy.
//
// Both are tiny (<= 1 MFLOP per galaxy); they exist to keep the model grid device-resident
// (no 2 x nz x nT x nB host array, no per-galaxy Python loop), not for roofline reasons.
#include "common.cuh"

namespace dl {

constexpr int GP_MAXB = 32;

struct GPModel {
    const double* KCg;     // [nB][nB][nzg][nzg]
    const double* KLg;
    const double* KCd;     // [nB][nzg]
    const double* KLd;
    const double* zGP;     // [nzg]   redshiftGridGP
    const double* fzGrid;  // [nzg]   1 + redshiftGridGP
    const double* tmpl;    // [nt][nB][nz]  <sed>_fluxredshiftmod tables on zgrid
    const double* zgrid;   // [nz]    fine redshift grid
    int nB, nzg, nt, nz;
    double V_C, V_L;
    // galaxy-independent tables of dl_gp_tables_dev (NULL: recomputed per galaxy)
    const double* dls_tab;  // [nzg]  approx_DL(zGP)
    const int* p1_tab;      // [nzg]  find_pos(fzGrid, 1 + zGP)
    const int* tlo_tab;     // [nzg]  bin_left(zgrid, zGP)      (template tables)
    const int* flo_tab;     // [nz]   bin_left(zGP, zgrid)      (coarse -> fine interpolation)
};

// delight/utils.py:12-13
__device__ __forceinline__ double approx_DL(double z) { return exp(30.5 * pow(z, 0.04) - 21.7); }

// scipy interp1d(kind='linear', fill_value='extrapolate'): searchsorted-left, clip to [1, n-1]
__device__ __forceinline__ int bin_left(const double* __restrict__ x, int n, double v) {
    int lo = 0, hi = n;            // first index with x[idx] >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] < v) lo = mid + 1; else hi = mid;
    }
    int idx = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
    return idx - 1;
}

__device__ __forceinline__ double lin_interp(const double* __restrict__ x,
                                             const double* __restrict__ y, int lo, double v) {
    const double slope = (y[lo + 1] - y[lo]) / (x[lo + 1] - x[lo]);
    return slope * (v - x[lo]) + y[lo];
}

// find_positions semantics (delight/utils_cy.pyx:10-22): first bin containing v, 0 if outside
__device__ __forceinline__ int find_pos(const double* __restrict__ g, int n, double v) {
    if (!(v >= g[0] && v <= g[n - 1])) return 0;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] < v) lo = mid; else hi = mid;
    }
    return lo;
}

// kernel_parts_interp for one element (delight/photoz_kernels_cy.pyx:25-35)
__device__ __forceinline__ double bilin(const double* __restrict__ G, int nzg,
                                        const double* __restrict__ fz, int p1, double x, int p2,
                                        double y) {
    const double xl = fz[p1], xh = fz[p1 + 1], yl = fz[p2], yh = fz[p2 + 1];
    const double inv = 1.0 / (xh - xl) / (yh - yl);
    return inv * ((xh - x) * (yh - y) * G[p1 * nzg + p2] + (x - xl) * (yh - y) * G[(p1 + 1) * nzg + p2] +
                  (xh - x) * (y - yl) * G[p1 * nzg + p2 + 1] +
                  (x - xl) * (y - yl) * G[(p1 + 1) * nzg + p2 + 1]);
}

// ------------------------------------------------------------------------------------------ fit
constexpr int FIT_WARPS = 4;

__global__ void __launch_bounds__(FIT_WARPS * 32) gp_fit_kernel(
    GPModel gm, int nT, const long long* __restrict__ nb, const long long* __restrict__ bands,
    const double* __restrict__ zs, const double* __restrict__ Y, const double* __restrict__ Yvar,
    double* __restrict__ rows, int numCol) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * FIT_WARPS + warp;
    if (gi >= nT) return;
    const int nB = gm.nB, nt = gm.nt;
    double* A = sm + (size_t)warp * (GP_MAXB * (GP_MAXB + 1) + (size_t)nt * GP_MAXB);  // [B][33]
    double* tm = A + GP_MAXB * (GP_MAXB + 1);                                         // [nt][32]
    const int B = (int)nb[gi];
    const double z = zs[gi];
    const long long* bnd = bands + (size_t)gi * nB;
    double* row = rows + (size_t)gi * numCol;
    for (int c = lane; c < numCol; c += 32) row[c] = 0.0;
    if (B < 1 || B > GP_MAXB) return;      // caller validates; row stays zero

    // template fluxes at z in the galaxy's bands (delight-learn.py:67-70)
    const int zlo = bin_left(gm.zgrid, gm.nz, z);
    for (int idx = lane; idx < nt * B; idx += 32) {
        const int t = idx / B, j = idx % B;
        const double* tab = gm.tmpl + ((size_t)t * nB + bnd[j]) * gm.nz;
        tm[t * GP_MAXB + j] = lin_interp(gm.zgrid, tab, zlo, z);
    }
    __syncwarp();
    // best template and scale (delight/utils.py:179-193; returnChi2 -> chi2 + FTT, :191)
    double bestv = INFINITY, bestell = 0.0;
    int bestt = 0x7fffffff;
    for (int t = lane; t < nt; t += 32) {
        double fot = 0, ftt = 0, foo = 0;
        for (int j = 0; j < B; ++j) {
            const double y = Y[(size_t)gi * nB + j], v = Yvar[(size_t)gi * nB + j];
            const double w = (y / v < 1e-6) ? 0.0 : 1.0 / v;
            const double f = tm[t * GP_MAXB + j];
            fot += f * y * w; ftt += f * f * w; foo += w * y * y;
        }
        const double val = foo - fot * fot / ftt + ftt;
        if (val < bestv) { bestv = val; bestt = t; bestell = fot / ftt; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bestv, s);
        const int ot = __shfl_xor_sync(0xffffffffu, bestt, s);
        const double oe = __shfl_xor_sync(0xffffffffu, bestell, s);
        if (ov < bestv || (ov == bestv && ot < bestt)) { bestv = ov; bestt = ot; bestell = oe; }
    }
    // all scores NaN (e.g. non-positive variances): np.argmin would return 0; keep indices in
    // range and let the NaN propagate into the row (the host reports it)
    if (bestt >= nt) { bestt = 0; bestell = nan(""); }
    const double ell = bestell;
    // A = K(X) + diag(Yvar)   (photoz_gp.py:205-206, photoz_kernels.py:221-231, 338-340)
    const double fz = 1.0 + z;
    const int p = find_pos(gm.fzGrid, gm.nzg, fz);
    const double dl = approx_DL(z);
    const double zp = (1.0 + z) * (1.0 + z) / (4.0 * M_PI * 1.0 * dl * dl);
    const double pref = zp * zp * ell * ell;
    if (lane < B) {
        const int bj = (int)bnd[lane];
        for (int k = 0; k < B; ++k) {
            const int bk = (int)bnd[k];
            const size_t g = ((size_t)bj * nB + bk) * gm.nzg * gm.nzg;
            const double kc = bilin(gm.KCg + g, gm.nzg, gm.fzGrid, p, fz, p, fz);
            const double kl = bilin(gm.KLg + g, gm.nzg, gm.fzGrid, p, fz, p, fz);
            double a = pref * (gm.V_C * kc + gm.V_L * kl);
            if (k == lane) a += Yvar[(size_t)gi * nB + lane];
            A[lane * (GP_MAXB + 1) + k] = a;
        }
    }
    __syncwarp();
    // lower Cholesky, right-looking (scipy.linalg.cholesky(lower=True), photoz_gp.py:209)
    for (int k = 0; k < B; ++k) {
        const double d = sqrt(A[k * (GP_MAXB + 1) + k]);
        __syncwarp();
        if (lane == k) A[k * (GP_MAXB + 1) + k] = d;
        if (lane > k && lane < B) A[lane * (GP_MAXB + 1) + k] /= d;
        __syncwarp();
        if (lane > k && lane < B) {
            const double ljk = A[lane * (GP_MAXB + 1) + k];
            for (int mcol = k + 1; mcol <= lane; ++mcol)
                A[lane * (GP_MAXB + 1) + mcol] -= ljk * A[mcol * (GP_MAXB + 1) + k];
        }
        __syncwarp();
    }
    // row = [B, z, ell, bands(B), betas(nt), L_tril(B(B+1)/2), D(B), 0...]  (delight-learn.py:85-90)
    if (lane == 0) {
        row[0] = (double)B; row[1] = z; row[2] = ell;
        row[3 + B + bestt] = 1.0;
    }
    if (lane < B) {
        row[3 + lane] = (double)bnd[lane];
        double* Lt = row + 3 + B + nt;
        for (int c = 0; c <= lane; ++c) Lt[lane * (lane + 1) / 2 + c] = A[lane * (GP_MAXB + 1) + c];
        double h = ell * tm[bestt * GP_MAXB + lane];
        if (!isfinite(h)) h = 0.0;                                  // photoz_gp.py:217
        row[3 + B + nt + B * (B + 1) / 2 + lane] = Y[(size_t)gi * nB + lane] - h;
    }
}

// -------------------------------------------------------------------------------------- predict
constexpr int PRED_THREADS = 128;

__global__ void __launch_bounds__(PRED_THREADS) gp_predict_kernel(
    GPModel gm, int nT, const double* __restrict__ rows, int numCol, int nzp,
    double* __restrict__ grid, double* __restrict__ train_z, const double* __restrict__ zeval,
    double* __restrict__ at_out) {
    extern __shared__ double sm[];
    const int nB = gm.nB, nzg = gm.nzg, nt = gm.nt, nz = gm.nz;
    double* L = sm;                                   // [32][33]
    double* beta = L + GP_MAXB * (GP_MAXB + 1);       // [32]
    double* nmean = beta + GP_MAXB;                   // [nB][nzg]
    double* nvar = nmean + (size_t)nB * nzg;          // [nB][nzg]
    double* wsc = nvar + (size_t)nB * nzg;            // [B][PRED_THREADS] forward-substitution
    __shared__ int s_B, s_best, s_p2, s_band[GP_MAXB];
    __shared__ double s_z, s_ell;

    const int gi = blockIdx.x;
    const double* row = rows + (size_t)gi * numCol;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int B = (int)row[0];
        s_B = B; s_z = row[1]; s_ell = row[2];
        for (int j = 0; j < B && j < GP_MAXB; ++j) s_band[j] = (int)row[3 + j];
        // bestType = argmax(betas), first maximum (photoz_gp.py:249)
        int best = 0;
        for (int t = 1; t < nt; ++t) if (row[3 + B + t] > row[3 + B + best]) best = t;
        s_best = best;
        s_p2 = find_pos(gm.fzGrid, nzg, 1.0 + row[1]);
        if (train_z) train_z[gi] = row[1];
    }
    __syncthreads();
    const int B = s_B;
    const double z = s_z, ell = s_ell;
    double* g_mean = grid ? grid + (size_t)gi * 2 * nB * nzp : nullptr;
    double* g_var = grid ? g_mean + (size_t)nB * nzp : nullptr;
    if (B < 1 || B > GP_MAXB) {           // invalid core: neutral rows
        if (grid)
            for (int i = tid; i < nB * nzp; i += PRED_THREADS) { g_mean[i] = 0.0; g_var[i] = 1.0; }
        if (at_out)
            for (int i = tid; i < 2 * nB; i += PRED_THREADS)
                at_out[(size_t)gi * 2 * nB + i] = i < nB ? 0.0 : 1.0;
        return;
    }
    // L from the row-major lower triangle (photoz_gp.py:251-253)
    const double* Lt = row + 3 + B + nt;
    for (int i = tid; i < B * B; i += PRED_THREADS) {
        const int r = i / B, c = i % B;
        L[r * (GP_MAXB + 1) + c] = c <= r ? Lt[r * (r + 1) / 2 + c] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {   // beta = cho_solve((L, lower), D)  (photoz_gp.py:254)
        const double* D = Lt + B * (B + 1) / 2;
        double w[GP_MAXB];
        for (int j = 0; j < B; ++j) {
            double s = D[j];
            for (int mcol = 0; mcol < j; ++mcol) s -= L[j * (GP_MAXB + 1) + mcol] * w[mcol];
            w[j] = s / L[j * (GP_MAXB + 1) + j];
        }
        for (int j = B - 1; j >= 0; --j) {
            double s = w[j];
            for (int mcol = j + 1; mcol < B; ++mcol) s -= L[mcol * (GP_MAXB + 1) + j] * beta[mcol];
            beta[j] = s / L[j * (GP_MAXB + 1) + j];
        }
    }
    __syncthreads();
    // GP on the coarse nodes (photoz_gp.py:263-291, 306-318; X_pred row = b * nzg + iz)
    const double fz = 1.0 + z;
    const double dlz = approx_DL(z);
    const int p2 = s_p2, best = s_best;
    // galaxy-side half of the bilinear weights (photoz_kernels_cy.pyx:25-35), once per galaxy
    const double yl = gm.fzGrid[p2], yh = gm.fzGrid[p2 + 1];
    const double wyl = yh - fz, wyh = fz - yl, dy = yh - yl;
    for (int node = tid; node < nB * nzg; node += PRED_THREADS) {
        const int bs = node / nzg, iz = node % nzg;
        const double zs = gm.zGP[iz];
        const double xs = 1.0 + zs;
        const int p1 = gm.p1_tab ? gm.p1_tab[iz] : find_pos(gm.fzGrid, nzg, xs);
        const double dls = gm.dls_tab ? gm.dls_tab[iz] : approx_DL(zs);
        // Zprefac (photoz_kernels.py:338-340), g_AB = 1
        const double zp = (1.0 + zs) * (1.0 + z) / (4.0 * M_PI * 1.0 * dls * dlz);
        const double pref = zp * zp * ell * ell;
        // node-side half of the bilinear weights; same operation order as bilin()
        const double xl = gm.fzGrid[p1], xh = gm.fzGrid[p1 + 1];
        const double inv = 1.0 / (xh - xl) / dy;
        const double w00 = (xh - xs) * wyl, w10 = (xs - xl) * wyl;
        const double w01 = (xh - xs) * wyh, w11 = (xs - xl) * wyh;
        const size_t o00 = (size_t)p1 * nzg + p2, o10 = o00 + nzg;
        double acc = 0.0, q = 0.0;
        for (int j = 0; j < B; ++j) {
            const size_t g = ((size_t)bs * nB + s_band[j]) * nzg * nzg;
            const double* C = gm.KCg + g;
            const double* Lg = gm.KLg + g;
            const double kc = inv * (w00 * C[o00] + w10 * C[o10] + w01 * C[o00 + 1] + w11 * C[o10 + 1]);
            const double kl = inv * (w00 * Lg[o00] + w10 * Lg[o10] + w01 * Lg[o00 + 1] + w11 * Lg[o10 + 1]);
            const double kj = pref * (gm.V_C * kc + gm.V_L * kl);
            acc = fma(kj, beta[j], acc);
            double s = kj;          // forward substitution: w = L^-1 k ; k^T A^-1 k = |w|^2
            for (int mcol = 0; mcol < j; ++mcol)
                s -= L[j * (GP_MAXB + 1) + mcol] * wsc[mcol * PRED_THREADS + tid];
            s /= L[j * (GP_MAXB + 1) + j];
            wsc[j * PRED_THREADS + tid] = s;
            q = fma(s, s, q);
        }
        // Kdiag (photoz_kernels.py:212-219, 239-240) with the per-band diagonal grids
        const double zpd = (1.0 + zs) * (1.0 + zs) / (4.0 * M_PI * 1.0 * dls * dls);
        const double kd = zpd * zpd * ell * ell *
                          (gm.V_C * gm.KCd[bs * nzg + iz] + gm.V_L * gm.KLd[bs * nzg + iz]);
        // mean function: ell * T[bestType, b*](z*)  (photoz_kernels.py:32-55)
        const double* tab = gm.tmpl + ((size_t)best * nB + bs) * nz;
        const int tlo = gm.tlo_tab ? gm.tlo_tab[iz] : bin_left(gm.zgrid, nz, zs);
        const double mf = ell * lin_interp(gm.zgrid, tab, tlo, zs);
        nmean[node] = acc + mf;
        nvar[node] = kd - q;
    }
    __syncthreads();
    if (at_out) {
        // predictAndInterpolate(np.array([z_i])) of the cross-validation in delight-learn.py:92-94:
        // the same node values interpolated at ONE redshift per galaxy -> [mean(nB), var(nB)]
        const double zv = zeval[gi];
        const int lo = bin_left(gm.zGP, nzg, zv);
        for (int b = tid; b < nB; b += PRED_THREADS) {
            at_out[(size_t)gi * 2 * nB + b] = lin_interp(gm.zGP, nmean + b * nzg, lo, zv);
            at_out[(size_t)gi * 2 * nB + nB + b] = lin_interp(gm.zGP, nvar + b * nzg, lo, zv);
        }
    }
    if (!grid) return;
    // coarse -> fine grid, linear (photoz_gp.py:321-335), into the fused layout
    for (int i = tid; i < nB * nzp; i += PRED_THREADS) {
        const int b = i / nzp, zf = i % nzp;
        double mval = 0.0, vval = 0.0;       // padding: mean 0, model variance 0
        if (zf < nz) {
            const double zv = gm.zgrid[zf];
            const int lo = gm.flo_tab ? gm.flo_tab[zf] : bin_left(gm.zGP, nzg, zv);
            mval = lin_interp(gm.zGP, nmean + b * nzg, lo, zv);
            vval = lin_interp(gm.zGP, nvar + b * nzg, lo, zv);
        }
        g_mean[i] = mval;
        g_var[i] = vval;
    }
}

// galaxy-independent lookups of gp_predict_kernel, once per model: without them every galaxy
// repeats 150 pow/exp calls and ~10k binary searches with dependent global loads
__global__ void gp_tables_kernel(int nzg, int nz, const double* __restrict__ zGP,
                                 const double* __restrict__ fzGrid,
                                 const double* __restrict__ zgrid, double* __restrict__ dls,
                                 int* __restrict__ p1, int* __restrict__ tlo,
                                 int* __restrict__ flo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nzg) {
        const double zs = zGP[i];
        dls[i] = approx_DL(zs);
        p1[i] = find_pos(fzGrid, nzg, 1.0 + zs);
        tlo[i] = bin_left(zgrid, nz, zs);
    }
    if (i < nz) flo[i] = bin_left(zGP, nzg, zgrid[i]);
}

}  // namespace dl

using namespace dl;

extern "C" {

long long dl_fused_padded_nz(long long nz);

// bytes of the table buffer of dl_gp_tables_dev
long long dl_gp_tables_bytes(long long nzg, long long nz) {
    return ((nzg * 8 + 255) & ~255ll) + 3 * ((nzg * 4 + 255) & ~255ll) + ((nz * 4 + 255) & ~255ll);
}

static void gp_set_tables(GPModel& gm, const void* tables) {
    gm.dls_tab = nullptr; gm.p1_tab = nullptr; gm.tlo_tab = nullptr; gm.flo_tab = nullptr;
    if (!tables) return;
    const char* w = (const char*)tables;
    const size_t a8 = ((size_t)gm.nzg * 8 + 255) & ~(size_t)255, a4 = ((size_t)gm.nzg * 4 + 255) & ~(size_t)255;
    gm.dls_tab = (const double*)w; w += a8;
    gm.p1_tab = (const int*)w; w += a4;
    gm.tlo_tab = (const int*)w; w += a4;
    w += a4;                                   // (spare)
    gm.flo_tab = (const int*)w;
}

// Galaxy-independent lookup tables of the GP predict kernels for one model (grids as in
// dl_gp_predict_dev): tables = device buffer of dl_gp_tables_bytes(nzg, nz) bytes.
int dl_gp_tables_dev(long long nzg, long long nz, const double* zGP, const double* fzGrid,
                     const double* zgrid, void* tables, void* stream) {
    DL_REQUIRE(nzg >= 2 && nz >= 2 && zGP && fzGrid && zgrid && tables, "bad arguments");
    GPModel gm;
    gm.nzg = (int)nzg; gm.nz = (int)nz;
    gp_set_tables(gm, tables);
    const long long n = nzg > nz ? nzg : nz;
    gp_tables_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (int)nzg, (int)nz, zGP, fzGrid, zgrid, const_cast<double*>(gm.dls_tab),
        const_cast<int*>(gm.p1_tab), const_cast<int*>(gm.tlo_tab), const_cast<int*>(gm.flo_tab));
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

static int gp_model(GPModel& gm, long long nB, long long nzg, long long nt, long long nz,
                    const double* KCg, const double* KLg, const double* KCd, const double* KLd,
                    const double* zGP, const double* fzGrid, const double* tmpl,
                    const double* zgrid, double V_C, double V_L) {
    DL_REQUIRE(nB >= 1 && nzg >= 2 && nt >= 1 && nz >= 2, "bad sizes");
    DL_REQUIRE(KCg && KLg && zGP && fzGrid && tmpl && zgrid, "null pointer");
    gm.KCg = KCg; gm.KLg = KLg; gm.KCd = KCd; gm.KLd = KLd; gm.zGP = zGP; gm.fzGrid = fzGrid;
    gm.tmpl = tmpl; gm.zgrid = zgrid;
    gm.nB = (int)nB; gm.nzg = (int)nzg; gm.nt = (int)nt; gm.nz = (int)nz;
    gm.V_C = V_C; gm.V_L = V_L;
    gp_set_tables(gm, nullptr);
    return DL_OK;
}

// rows [nT][numCol], numCol >= 3 + nB + nt + nB(nB+1)/2 + nB (delight-learn.py:46)
int dl_gp_fit_dev(long long nT, long long nB, long long nt, long long nzg, long long nz,
                  const double* KC_grid, const double* KL_grid, const double* zGP,
                  const double* fzGrid, const double* tmpl, const double* zgrid, double V_C,
                  double V_L, const long long* nbands, const long long* bands, const double* z,
                  const double* Y, const double* Yvar, double* rows, long long numCol,
                  void* stream) {
    GPModel gm;
    int rc = gp_model(gm, nB, nzg, nt, nz, KC_grid, KL_grid, nullptr, nullptr, zGP, fzGrid, tmpl,
                      zgrid, V_C, V_L);
    if (rc) return rc;
    if (nT == 0) return DL_OK;
    DL_REQUIRE(nB <= GP_MAXB, "at most 32 bands");
    DL_REQUIRE(numCol >= 3 + nB + nt + nB * (nB + 1) / 2 + nB, "numCol too small");
    DL_REQUIRE(nbands && bands && z && Y && Yvar && rows, "null pointer");
    const size_t smem = sizeof(double) * FIT_WARPS * (GP_MAXB * (GP_MAXB + 1) + (size_t)nt * GP_MAXB);
    DL_REQUIRE(smem <= 200 * 1024, "too many templates");
    DL_CUDA_CHECK(cudaFuncSetAttribute(gp_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    gp_fit_kernel<<<(unsigned)((nT + FIT_WARPS - 1) / FIT_WARPS), FIT_WARPS * 32, smem,
                    (cudaStream_t)stream>>>(gm, (int)nT, nbands, bands, z, Y, Yvar, rows,
                                            (int)numCol);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

// grid [nT][2 nB][nzp] (nzp = dl_fused_padded_nz(nz)); train_z [nT] or NULL
int dl_gp_predict_dev(long long nT, long long nB, long long nt, long long nzg, long long nz,
                      const double* KC_grid, const double* KL_grid, const double* KC_diag,
                      const double* KL_diag, const double* zGP, const double* fzGrid,
                      const double* tmpl, const double* zgrid, double V_C, double V_L,
                      const double* rows, long long numCol, double* grid, double* train_z,
                      const void* tables, void* stream) {
    GPModel gm;
    int rc = gp_model(gm, nB, nzg, nt, nz, KC_grid, KL_grid, KC_diag, KL_diag, zGP, fzGrid, tmpl,
                      zgrid, V_C, V_L);
    if (rc) return rc;
    gp_set_tables(gm, tables);
    if (nT == 0) return DL_OK;
    DL_REQUIRE(KC_diag && KL_diag && rows && grid, "null pointer");
    DL_REQUIRE(nB <= GP_MAXB, "at most 32 bands");
    DL_REQUIRE(numCol >= 3 + nB + nt + nB * (nB + 1) / 2 + nB, "numCol too small");
    const int nzp = (int)dl_fused_padded_nz(nz);
    const size_t smem = sizeof(double) * (GP_MAXB * (GP_MAXB + 1) + GP_MAXB + 2 * (size_t)nB * nzg +
                                          (size_t)GP_MAXB * PRED_THREADS);
    DL_REQUIRE(smem <= 220 * 1024, "nB * nzGP too large for shared memory");
    DL_CUDA_CHECK(cudaFuncSetAttribute(gp_predict_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gp_predict_kernel<<<(unsigned)nT, PRED_THREADS, smem, (cudaStream_t)stream>>>(
        gm, (int)nT, rows, (int)numCol, nzp, grid, train_z, nullptr, nullptr);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

// Same GP, evaluated at ONE redshift per galaxy: z_eval [nT] -> out [nT][2 nB] (band means, then
// band variances).  delight-learn.py:92-94 calls predictAndInterpolate(np.array([z]), ell=ell)
// with the galaxy's own redshift for the cross-validation chi2 (:95-100).
int dl_gp_predict_at_dev(long long nT, long long nB, long long nt, long long nzg, long long nz,
                         const double* KC_grid, const double* KL_grid, const double* KC_diag,
                         const double* KL_diag, const double* zGP, const double* fzGrid,
                         const double* tmpl, const double* zgrid, double V_C, double V_L,
                         const double* rows, long long numCol, const double* z_eval, double* out,
                         const void* tables, void* stream) {
    GPModel gm;
    int rc = gp_model(gm, nB, nzg, nt, nz, KC_grid, KL_grid, KC_diag, KL_diag, zGP, fzGrid, tmpl,
                      zgrid, V_C, V_L);
    if (rc) return rc;
    gp_set_tables(gm, tables);
    if (nT == 0) return DL_OK;
    DL_REQUIRE(KC_diag && KL_diag && rows && z_eval && out, "null pointer");
    DL_REQUIRE(nB <= GP_MAXB, "at most 32 bands");
    DL_REQUIRE(numCol >= 3 + nB + nt + nB * (nB + 1) / 2 + nB, "numCol too small");
    const size_t smem = sizeof(double) * (GP_MAXB * (GP_MAXB + 1) + GP_MAXB + 2 * (size_t)nB * nzg +
                                          (size_t)GP_MAXB * PRED_THREADS);
    DL_REQUIRE(smem <= 220 * 1024, "nB * nzGP too large for shared memory");
    DL_CUDA_CHECK(cudaFuncSetAttribute(gp_predict_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gp_predict_kernel<<<(unsigned)nT, PRED_THREADS, smem, (cudaStream_t)stream>>>(
        gm, (int)nT, rows, (int)numCol, 0, nullptr, nullptr, z_eval, out);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

}  // extern "C"

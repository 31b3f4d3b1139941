/*
 * delight_b200 -- C ABI of the B200-native photo-z inference hot path of ixkael/Delight.
 *
 * libdelight_b200.so exports exactly these symbols.  Conventions
 *   - every function returns 0 (DL_OK) or a DL_ERR_* code; dl_last_error() gives the message
 *     (thread local).  The reference's Cython operators have no error channel at all
 *     (boundscheck=False, photoz_kernels_cy.pyx:1); shape checks that Cython's memoryview
 *     coercion performs (ValueError) are done by the Python shim modules.
 *   - all floating point data is IEEE double, all indices are 64-bit signed (Cython `long`).
 *   - `_host` entry points take HOST pointers and do their own transfers (these are what an FFI
 *     binding of the reference operators binds);  `_dev` entry points take DEVICE pointers and a
 *     cudaStream_t passed as void* (NULL = default stream) and never synchronise.
 *   - arrays are C-contiguous unless a leading dimension `ld` (in elements) is given.
 *
 * References are to files under the reference repository root.
 */
#ifndef DELIGHT_B200_H
#define DELIGHT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DL_OK 0
#define DL_ERR_INVALID 1
#define DL_ERR_CUDA 2
#define DL_ERR_NOMEM 3
#define DL_ERR_UNSUPPORTED 4

const char* dl_last_error(void);
int dl_version(void);
int dl_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* global_mem);

/* ------------------------------------------------------------------------------------------
 * (1) approx_flux_likelihood_cy          replaces delight/utils_cy.pyx:83-128
 *     like [nz][nt] (out), f_obs/f_obs_var [nf], f_mod/f_mod_covar [nz][nt][nf],
 *     ell_hat/ell_var [nz].  Result = exp(L - max L) over the whole call, as the reference.
 * ---------------------------------------------------------------------------------------- */
int dl_approx_flux_likelihood_host(double* like, long long nz, long long nt, long long nf,
                                   const double* f_obs, const double* f_obs_var,
                                   const double* f_mod, const double* f_mod_covar,
                                   const double* ell_hat, const double* ell_var);
int dl_approx_flux_likelihood_dev(double* like, long long nz, long long nt, long long nf,
                                  const double* f_obs, const double* f_obs_var,
                                  const double* f_mod, const double* f_mod_covar,
                                  const double* ell_hat, const double* ell_var,
                                  void* scratch8, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2) kernelparts                        replaces delight/photoz_kernels_cy.pyx:103-177
 *     Outputs [NO1][NO2] are ACCUMULATED into and then divided by norms[b1]*norms[b2]
 *     (pyx:149-177); gradients are only touched when grad_needed.  Output order KL, KC as in
 *     the reference signature.  lines_sig of the reference signature is unused there (never
 *     read) and is not part of this ABI.  0 <= NL <= 8.
 * ---------------------------------------------------------------------------------------- */
int dl_kernelparts_host(long long NO1, long long NO2, long long NC, long long NL, double alpha_C,
                        double alpha_L, long long nB, const double* fcoefs_amp,
                        const double* fcoefs_mu, const double* fcoefs_sig,
                        const double* lines_mu, const double* norms, const long long* b1,
                        const double* fz1, const long long* b2, const double* fz2,
                        int grad_needed, double* KL, double* KC, double* D_alpha_C,
                        double* D_alpha_L, double* D_alpha_z);
int dl_kernelparts_dev(long long NO1, long long NO2, long long NC, long long NL, double alpha_C,
                       double alpha_L, const double* fcoefs_amp, const double* fcoefs_mu,
                       const double* fcoefs_sig, const double* lines_mu, const double* norms,
                       const long long* b1, const double* fz1, const long long* b2,
                       const double* fz2, int grad_needed, double* KL, double* KC,
                       double* D_alpha_C, double* D_alpha_L, double* D_alpha_z, long long ld,
                       void* stream);

/* (3) kernelparts_diag                   replaces delight/photoz_kernels_cy.pyx:38-100
 *     KL/KC [NO1] overwritten, D_alpha_C/D_alpha_L accumulated then normalised. */
int dl_kernelparts_diag_host(long long NO1, long long NC, long long NL, double alpha_C,
                             double alpha_L, long long nB, const double* fcoefs_amp,
                             const double* fcoefs_mu, const double* fcoefs_sig,
                             const double* lines_mu, const double* norms, const long long* b1,
                             const double* fz1, int grad_needed, double* KL, double* KC,
                             double* D_alpha_C, double* D_alpha_L);
int dl_kernelparts_diag_dev(long long NO1, long long NC, long long NL, double alpha_C,
                            double alpha_L, const double* fcoefs_amp, const double* fcoefs_mu,
                            const double* fcoefs_sig, const double* lines_mu,
                            const double* norms, const long long* b1, const double* fz1,
                            int grad_needed, double* KL, double* KC, double* D_alpha_C,
                            double* D_alpha_L, void* stream);

/* (4) find_positions                     replaces delight/utils_cy.pyx:10-22
 *     p1s[o] = first p with fzGrid[p] <= fz1[o] <= fzGrid[p+1]; untouched when outside. */
int dl_find_positions_host(long long NO1, long long nz, const double* fz1, long long* p1s,
                           const double* fzGrid);
int dl_find_positions_dev(long long NO1, long long nz, const double* fz1, long long* p1s,
                          const double* fzGrid, void* stream);

/* (5) kernel_parts_interp                replaces delight/photoz_kernels_cy.pyx:9-35
 *     Kgrid [nB][nB][nzg][nzg], Kinterp [NO1][NO2] overwritten. */
int dl_kernel_parts_interp_host(long long NO1, long long NO2, double* Kinterp,
                                const long long* b1, const double* fz1, const long long* p1s,
                                const long long* b2, const double* fz2, const long long* p2s,
                                const double* fzGrid, const double* Kgrid, long long nB,
                                long long nzg);
int dl_kernel_parts_interp_dev(long long NO1, long long NO2, double* Kinterp, long long ld,
                               const long long* b1, const double* fz1, const long long* p1s,
                               const long long* b2, const double* fz2, const long long* p2s,
                               const double* fzGrid, const double* Kgrid, long long nB,
                               long long nzg, void* stream);

/* (6) bilininterp_precomputedbins        replaces delight/utils_cy.pyx:25-50 (used by the
 *     reference's tests only, tests/test_utils.py:89-115; here for a complete utils_cy surface)
 *     Kgrid [numBands][n1][n2], Kinterp [numBands][nobj] overwritten. */
int dl_bilininterp_precomputedbins_host(long long numBands, long long nobj, double* Kinterp,
                                        const double* v1s, const double* v2s,
                                        const long long* p1s, const long long* p2s,
                                        const double* grid1, const double* grid2,
                                        const double* Kgrid, long long n1, long long n2);
int dl_bilininterp_precomputedbins_dev(long long numBands, long long nobj, double* Kinterp,
                                       long long ld, const double* v1s, const double* v2s,
                                       const long long* p1s, const long long* p2s,
                                       const double* grid1, const double* grid2,
                                       const double* Kgrid, long long n1, long long n2,
                                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched / fused entry points used by the new host path (device pointers only).
 * ---------------------------------------------------------------------------------------- */

/* Photoz_kernel.construct_interpolators  replaces delight/photoz_kernels.py:342-416
 * (nB^2 kernelparts calls + nB kernelparts_diag calls) by one launch.
 * KC_grid/KL_grid [nB][nB][nzg][nzg], KC_diag/KL_diag [nB][nzg], fzGrid = 1 + redshiftGridGP. */
int dl_kernel_grids_dev(long long nB, long long nzg, long long NC, long long NL, double alpha_C,
                        double alpha_L, const double* fcoefs_amp, const double* fcoefs_mu,
                        const double* fcoefs_sig, const double* lines_mu, const double* norms,
                        const double* fzGrid, double* KC_grid, double* KL_grid, double* KC_diag,
                        double* KL_diag, void* stream);

/* Model-grid layout of the fused likelihood: [nT][2*nB][nzp], rows 0..nB-1 = band means,
 * nB..2nB-1 = band variances, z contiguous, nzp = dl_fused_padded_nz(nz) (padding: mean 0,
 * variance 0).  Priors are [nT][nzp] (padding 0). */
int dl_fused_chunk(void);                      /* nz is padded to a multiple of this (128) */
int dl_fused_targets_per_cta(int nf, int per_target_bands);   /* 32 or 64 (launch sizing) */
int dl_fused_ctas_per_sm(int nf, int per_target_bands);       /* resident CTAs per SM: 2 or 1 */
long long dl_fused_padded_nz(long long nz);
/* scratch bytes of dl_like_fused_dev: per target 8 B x (nzp + nT) + ~1.5 KB; the nT term drops
 * when the caller passes an evidence array or asks for no evidences (K = 0, evid = NULL) */
long long dl_fused_workspace_bytes(long long nO, long long nT, long long nz,
                                   int evid_is_output_or_unused);

/* legacy [nz][nT][nB] arrays of scripts/delight-apply.py:84-85 -> fused layout */
int dl_relayout_grid_dev(long long nz, long long nT, long long nB, const double* model_mean,
                         const double* model_covar, double* grid, void* stream);
/* explicit prior [nz][nT] (e.g. scripts/templateFitting.py:79-82) -> [nT][nzp] */
int dl_relayout_prior_dev(long long nz, long long nT, const double* prior_zt, double* prior_tz,
                          void* stream);
/* Gaussian redshift prior of scripts/delight-apply.py:104 -> [nT][nzp] */
int dl_zprior_dev(long long nT, long long nz, const double* zgrid, const double* ztrain,
                  double sigma, double* prior_tz, void* stream);

/* Template prior of scripts/templateFitting.py:79-83 in the fused layout [nt][nzp]:
 * p(z,t) = p_t[t] z / p_z_t[t]^2 exp(-z^2 / (2 p_z_t[t]^2)); p_t, p_z_t device arrays [nt]. */
int dl_template_prior_dev(long long nt, long long nz, const double* zgrid, const double* p_t,
                          const double* p_z_t, double* prior_tz, void* stream);

/* Fused likelihood + apply epilogue   replaces scripts/delight-apply.py:137-162 x nO targets
 * (approx_flux_likelihood_cy, like*=prior, pdf=sum_t, evid=trapz_z, top-K) for nO targets that
 * are observed in nf bands each (nf <= 32).
 *   union_bands  HOST array [n_union]: the band rows of the grid that are staged per step.
 *   target_rows  NULL: every target is observed in exactly union_bands (n_union == nf) -- the
 *                fastest form; else DEVICE int32 [nO][nf]: band f of target o is
 *                union_bands[target_rows[o][f]] (the valid-band list of io.py:331-337 per target,
 *                as delight-apply.py:141-142 selects model_mean[:, :, bands] per target), so a
 *                catalogue with missing bands needs one call per band COUNT, not per band set.
 *   obs/var [nO][nf] in the target's own band order; ell_hat/ell_var scalars, or device vectors
 *   ell_hat_z/ell_var_z [nz] (not together with target_rows)
 *   pdf [nO][ld_pdf] out; evid [nO][nT] out or NULL; top_idx [nO][K] / top_ev [nO][K] out
 *   (K = 0: skipped).  Ties: descending evidence, equal -> larger index first.
 *   cov_scale multiplies the model variance (1.0 for delight-apply; delight-optimize.py:122
 *   passes V_C * model_covar).
 * Stream-ordered, no host synchronisation.  PDFs are accumulated with floating-point atomics:
 * repeated calls agree to rounding (~1e-16 relative), not bit for bit. */
int dl_like_fused_dev(long long nO, long long nT, long long nz, long long nB, int nf,
                      int n_union, const int* union_bands, const int* target_rows,
                      const double* obs, const double* var,
                      const double* grid, const double* prior, const double* zgrid,
                      double ell_hat, double ell_var, const double* ell_hat_z,
                      const double* ell_var_z, double cov_scale, int K, double* pdf,
                      long long ld_pdf,
                      double* evid, long long* top_idx, double* top_ev, void* workspace,
                      long long workspace_bytes, void* stream);

/* Compressed second pass               replaces scripts/delight-apply.py:109-135,145-146 x nO
 * (useCompression with the compressIndices file of a first pass present): per target only the K
 * stored training galaxies sel[o][0..K) (device, int64; entries outside [0, nT) are skipped)
 * enter the likelihood, which is the NumPy twin delight/utils.py:115-176 with
 * marginalizeEll=True, renormalize=True and `normalized` as given (the script passes False: no
 * log-determinant terms); like *= prior; pdf = sum over the selected galaxies;
 * evid[o][k] = trapz_z of column sel[o][k] (0 for skipped entries; may be NULL).
 * Same grid / prior layout and band-group convention as dl_like_fused_dev; nf <= 64. */
int dl_like_compressed_dev(long long nO, long long nT, long long nz, long long nB, int nf,
                           const int* bands, const double* obs, const double* var,
                           const double* grid, const double* prior, const double* zgrid,
                           double ell_hat, double ell_var, const double* ell_hat_z,
                           const double* ell_var_z, double cov_scale, int K,
                           const long long* sel, int normalized, double* pdf, long long ld_pdf,
                           double* evid, void* stream);

/* computeMetrics core                  replaces delight/utils.py:213-227 (first 7 columns)
 * out [nO][8]: ztrue zmean std(zmean) zMAP std(zMAP) pdf(ztrue) cumpdf(ztrue) argmax-index */
int dl_metrics_dev(long long nO, long long nz, const double* pdf, long long ld_pdf,
                   const double* zgrid, const double* ztrue, double* out, void* stream);

/* CIlevel                               replaces delight/utils.py:196-205 for nfrac <= 8
 * fractions at once (computeMetrics passes 1 - confidenceLevels, utils.py:223-225).
 * fractions: device array [nfrac]; numlevels = 200 in the reference; out [nO][nfrac]. */
int dl_ci_levels_dev(long long nO, long long nz, const double* pdf, long long ld_pdf,
                     const double* zgrid, int nfrac, const double* fractions, int numlevels,
                     double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched per-training-galaxy GP (device pointers).  Model = kernel grids from
 * dl_kernel_grids_dev on zGP (redshiftGridGP) / fzGrid (= 1 + zGP), template flux tables
 * tmpl [nt][nB][nz] on the fine grid zgrid [nz] (<sed>_fluxredshiftmod.txt, io.py:263-277),
 * hyper-parameters V_C, V_L.  Rows are the reference's gpparams rows (delight-learn.py:46,85-90):
 *   [B, z, ell, bands(B), betas(nt), L_tril(B(B+1)/2 row-major lower), D(B), 0-padding]
 *
 * dl_gp_fit_dev      replaces the loop body of scripts/delight-learn.py:61-90
 *                    (scalefree_flux_likelihood utils.py:179-193 + PhotozGP.setData/getCore
 *                    photoz_gp.py:186-234).  nbands [nT], bands/Y/Yvar [nT][nB] left-packed.
 * dl_gp_predict_dev  replaces scripts/delight-apply.py:89-98 (PhotozGP.setCore +
 *                    predictAndInterpolate, photoz_gp.py:236-347); writes the fused layout
 *                    grid [nT][2 nB][nzp] and the training redshifts.
 * ---------------------------------------------------------------------------------------- */
int dl_gp_fit_dev(long long nT, long long nB, long long nt, long long nzg, long long nz,
                  const double* KC_grid, const double* KL_grid, const double* zGP,
                  const double* fzGrid, const double* tmpl, const double* zgrid, double V_C,
                  double V_L, const long long* nbands, const long long* bands, const double* z,
                  const double* Y, const double* Yvar, double* rows, long long numCol,
                  void* stream);
int dl_gp_predict_dev(long long nT, long long nB, long long nt, long long nzg, long long nz,
                      const double* KC_grid, const double* KL_grid, const double* KC_diag,
                      const double* KL_diag, const double* zGP, const double* fzGrid,
                      const double* tmpl, const double* zgrid, double V_C, double V_L,
                      const double* rows, long long numCol, double* grid, double* train_z,
                      const void* tables, void* stream);
/* Galaxy-independent lookup tables of the two predict entry points (distance factors, grid
 * positions and interpolation bins of the GP nodes and of the fine grid), computed once per
 * model: tables = device buffer of dl_gp_tables_bytes(nzg, nz) bytes.  The predict calls accept
 * tables = NULL (everything recomputed per galaxy; several times slower). */
long long dl_gp_tables_bytes(long long nzg, long long nz);
int dl_gp_tables_dev(long long nzg, long long nz, const double* zGP, const double* fzGrid,
                     const double* zgrid, void* tables, void* stream);

/* The same GP evaluated at ONE redshift per training galaxy   replaces delight-learn.py:92-94
 * (predictAndInterpolate(np.array([z]), ell=ell) of the cross-validation): z_eval [nT] ->
 * out [nT][2 nB] = band means then band variances. */
int dl_gp_predict_at_dev(long long nT, long long nB, long long nt, long long nzg, long long nz,
                         const double* KC_grid, const double* KL_grid, const double* KC_diag,
                         const double* KL_diag, const double* zGP, const double* fzGrid,
                         const double* tmpl, const double* zgrid, double V_C, double V_L,
                         const double* rows, long long numCol, const double* z_eval, double* out,
                         const void* tables, void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (used by bench.py; no effect on results)
 * ---------------------------------------------------------------------------------------- */
long long dl_launch_count(void);                 /* kernels launched by this library so far */
int dl_timing_enable(int on);                    /* CUDA-event timing of the fused kernel */
int dl_timing_collect(double* total_ms, int* launches);
int dl_fp64_peak(int reps, double* tflops);      /* measured FP64 FMA peak of the device */
/* accuracy self-test of the kernel's fast reciprocal / rsqrt / exp: out [5][n] (device) */
int dl_selftest_math_dev(long long n, const double* x, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DELIGHT_B200_H */

"""CPU ORACLE (Python side) -- TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH.

ctypes bindings for oracle/delight_oracle.c plus NumPy restatements of the Python-level algebra
on Delight's hot path (GP fit / predict, metrics, template-fitting likelihood).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; nothing under
delight_b200/ does.

Parity status: PINNED against the compiled reference (oracle/_ref) through the fixtures in
tests/golden/ (see tests/test_oracle.py) and the reference's own golden vectors.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_lp = ctypes.POINTER(ctypes.c_long)


def build():
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "delight_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fPIC", "-fopenmp", "-shared", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _l(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_lp)


# ------------------------------------------------------------------------------------------------
# likelihood  (delight/utils_cy.pyx:83-128)
# ------------------------------------------------------------------------------------------------
def flux_loglike(f_obs, f_obs_var, f_mod, f_mod_covar, ell_hat, ell_var):
    nz, nt, nf = f_mod.shape
    out = np.zeros((nz, nt))
    keep = [_d(f_obs), _d(f_obs_var), _d(f_mod), _d(f_mod_covar), _d(ell_hat), _d(ell_var)]
    lib().orc_flux_loglike(out.ctypes.data_as(_dp), ctypes.c_long(nz), ctypes.c_long(nt),
                           ctypes.c_long(nf), *[k[1] for k in keep])
    return out


def approx_flux_likelihood_cy(like, nz, nt, nf, f_obs, f_obs_var, f_mod, f_mod_covar,
                              ell_hat, ell_var):
    """Same signature and in-place contract as delight/utils_cy.pyx:83-94."""
    out = np.zeros((nz, nt))
    keep = [_d(f_obs), _d(f_obs_var), _d(f_mod), _d(f_mod_covar), _d(ell_hat), _d(ell_var)]
    lib().orc_approx_flux_likelihood(out.ctypes.data_as(_dp), ctypes.c_long(nz),
                                     ctypes.c_long(nt), ctypes.c_long(nf),
                                     *[k[1] for k in keep])
    like[...] = out


def apply_fused(fluxes, fluxes_var, nf, bands, model_mean, model_covar, prior, zgrid,
                ell_hat, ell_var, K, want_evid=False):
    """Batch of targets through likelihood + the apply epilogue (delight-apply.py:137-162).

    fluxes, fluxes_var, bands: [nO, nB] packed left (first nf[o] entries valid).
    model_mean/model_covar: [nz, nT, nB];  prior: [nz, nT].
    Returns dict(pdf [nO,nz], top_idx [nO,K], top_ev [nO,K], like_max_pos [nO], evid [nO,nT]|None).
    """
    nz, nT, nB = model_mean.shape
    nO = fluxes.shape[0]
    pdf = np.zeros((nO, nz))
    top_idx = np.zeros((nO, max(K, 1)), dtype=np.int64)
    top_ev = np.zeros((nO, max(K, 1)))
    pos = np.zeros(nO, dtype=np.int64)
    evid = np.zeros((nO, nT)) if want_evid else None
    keep = [_l(nf), _l(bands), _d(fluxes), _d(fluxes_var), _d(model_mean), _d(model_covar),
            _d(prior), _d(zgrid), _d(ell_hat), _d(ell_var)]
    lib().orc_apply_fused(ctypes.c_long(nO), ctypes.c_long(nT), ctypes.c_long(nz),
                          ctypes.c_long(nB), ctypes.c_long(K), *[k[1] for k in keep],
                          pdf.ctypes.data_as(_dp),
                          evid.ctypes.data_as(_dp) if want_evid else None,
                          top_idx.ctypes.data_as(_lp), top_ev.ctypes.data_as(_dp),
                          pos.ctypes.data_as(_lp))
    return dict(pdf=pdf, top_idx=top_idx[:, :K], top_ev=top_ev[:, :K], like_max_pos=pos,
                evid=evid)


# ------------------------------------------------------------------------------------------------
# compressed second pass  (scripts/delight-apply.py:109-135,145-147; delight/utils.py:115-176)
# ------------------------------------------------------------------------------------------------
def flux_like_twin(f_obs, f_obs_var, f_mod, f_mod_covar, ell_hat, ell_var, normalized):
    """delight.utils.approx_flux_likelihood (utils.py:115-176) for the call the compressed branch
    makes: 3-D covariance (two iterations), marginalizeEll=True, renormalize=True, all
    ell_var > 0.  ell_hat / ell_var: scalars or [nz] arrays (broadcast over training galaxies).
    `normalized` switches the log-determinant terms (:165-168); delight-apply.py:131 passes False.
    Returns like[nz, nt] = exp(L - max L)."""
    f_obs = np.asarray(f_obs, dtype=np.float64)
    f_obs_var = np.asarray(f_obs_var, dtype=np.float64)
    nz, nt, nf = f_mod.shape
    eh = np.broadcast_to(np.asarray(ell_hat, dtype=np.float64), (nz,))[:, None]
    ev = np.broadcast_to(np.asarray(ell_var, dtype=np.float64), (nz,))[:, None]
    assert np.all(ev > 0)
    ellML = 0.0
    for _ in range(2):                                                    # utils.py:141-160
        var = f_obs_var[None, None, :] + ellML ** 2 * f_mod_covar
        invvar = 1.0 / var
        FOT = np.sum(f_mod * f_obs[None, None, :] * invvar, axis=2) + eh / ev
        FTT = np.sum(f_mod ** 2 * invvar, axis=2) + 1.0 / ev
        FOO = np.sum(f_obs[None, None, :] ** 2 * invvar, axis=2) + eh ** 2 / ev
        log_sigma_det = np.sum(np.log(var), axis=2)
        ellML = (FOT / FTT)[:, :, None]
    chi2 = FOO - FOT ** 2 / FTT                                           # :163
    logDenom = 0.0
    if normalized:                                                        # :167-170
        logDenom = logDenom + log_sigma_det + nf * np.log(2 * np.pi) + np.log(2 * np.pi * ev)
    logDenom = logDenom + np.log(FTT) - np.log(2 * np.pi)                 # :171-174
    like = -0.5 * chi2 - 0.5 * logDenom
    like -= like.max()                                                    # :175-176
    return np.exp(like)


def apply_compressed(fluxes, fluxes_var, nf, bands, model_mean, model_covar, prior, zgrid,
                     ell_hat, ell_var, indices, normalized=False):
    """The second ("compressed") pass of delight-apply.py:109-135 as its author evidently meant
    it: per target only the training galaxies whose indices were stored by the first pass
    (compressIndicesFile) enter the likelihood.  The reference assigns the result to
    `like_grid2` and then keeps using the stale `like_grid` (:126,:135) and hands 1-D [nz]
    ell arrays to the NumPy twin, which cannot broadcast them against [nz, K]; this restatement
    uses the computed array and broadcasts ell over the training-galaxy axis.
    `sel = np.in1d(targetIndices, indices)` (:125) is a mask, so the selected columns come in
    ascending training index order, each once.
    indices: [nO, K] int (entries outside [0, nT) are ignored, like the mask does).
    Returns dict(pdf [nO, nz], sel list of index arrays, evid list of [len(sel)] arrays)."""
    nz, nT, nB = model_mean.shape
    nO = fluxes.shape[0]
    pdf = np.zeros((nO, nz))
    sels, evids = [], []
    for o in range(nO):
        b = np.asarray(bands[o, :nf[o]], dtype=np.int64)
        sel = np.unique(indices[o][(indices[o] >= 0) & (indices[o] < nT)])
        if sel.size == 0:          # nothing stored (the reference would fail on like.max())
            sels.append(sel)
            evids.append(np.zeros(0))
            continue
        like = flux_like_twin(fluxes[o, :nf[o]], fluxes_var[o, :nf[o]],
                              model_mean[:, sel, :][:, :, b], model_covar[:, sel, :][:, :, b],
                              ell_hat, ell_var, normalized)
        like = like * prior[:, sel]                                       # :135
        pdf[o] = like.sum(axis=1)                                         # :145
        evids.append(np.trapezoid(like, x=zgrid, axis=0) if hasattr(np, 'trapezoid')
                     else np.trapz(like, x=zgrid, axis=0))                     # :146
        sels.append(sel)
    return dict(pdf=pdf, sel=sels, evid=evids)


# ------------------------------------------------------------------------------------------------
# GP kernel pieces  (delight/photoz_kernels_cy.pyx)
# ------------------------------------------------------------------------------------------------
def kernelparts(NO1, NO2, NC, NL, alpha_C, alpha_L, fcoefs_amp, fcoefs_mu, fcoefs_sig,
                lines_mu, lines_sig, norms, b1, fz1, b2, fz2, grad_needed,
                KL, KC, D_alpha_C, D_alpha_L, D_alpha_z):
    """Same signature / accumulate-then-normalise contract as photoz_kernels_cy.pyx:103-122."""
    outs = [np.ascontiguousarray(a, dtype=np.float64).copy()
            for a in (KL, KC, D_alpha_C, D_alpha_L, D_alpha_z)]
    keep = [_d(fcoefs_amp), _d(fcoefs_mu), _d(fcoefs_sig), _d(lines_mu), _d(norms)]
    kb1, kf1, kb2, kf2 = _l(b1), _d(fz1), _l(b2), _d(fz2)
    lib().orc_kernelparts(int(NO1), int(NO2), int(NC), int(NL), ctypes.c_double(alpha_C),
                          ctypes.c_double(alpha_L), *[k[1] for k in keep],
                          kb1[1], kf1[1], kb2[1], kf2[1], int(bool(grad_needed)),
                          *[o.ctypes.data_as(_dp) for o in outs])
    for dst, src in zip((KL, KC, D_alpha_C, D_alpha_L, D_alpha_z), outs):
        dst[...] = src


def kernelparts_diag(NO1, NC, NL, alpha_C, alpha_L, fcoefs_amp, fcoefs_mu, fcoefs_sig,
                     lines_mu, lines_sig, norms, b1, fz1, grad_needed,
                     KL, KC, D_alpha_C, D_alpha_L):
    """photoz_kernels_cy.pyx:38-54."""
    outs = [np.ascontiguousarray(a, dtype=np.float64).copy()
            for a in (KL, KC, D_alpha_C, D_alpha_L)]
    keep = [_d(fcoefs_amp), _d(fcoefs_mu), _d(fcoefs_sig), _d(lines_mu), _d(norms)]
    kb1, kf1 = _l(b1), _d(fz1)
    lib().orc_kernelparts_diag(int(NO1), int(NC), int(NL), ctypes.c_double(alpha_C),
                               ctypes.c_double(alpha_L), *[k[1] for k in keep],
                               kb1[1], kf1[1], int(bool(grad_needed)),
                               *[o.ctypes.data_as(_dp) for o in outs])
    for dst, src in zip((KL, KC, D_alpha_C, D_alpha_L), outs):
        dst[...] = src


def find_positions(NO1, nz, fz1, p1s, fzGrid):
    """utils_cy.pyx:10-22 (entries outside the grid are left untouched)."""
    p = np.ascontiguousarray(p1s, dtype=np.int64).copy()
    kf, kg = _d(fz1), _d(fzGrid)
    lib().orc_find_positions(int(NO1), int(nz), kf[1], p.ctypes.data_as(_lp), kg[1])
    p1s[...] = p


def kernel_parts_interp(NO1, NO2, Kinterp, b1, fz1, p1s, b2, fz2, p2s, fzGrid, Kgrid):
    """photoz_kernels_cy.pyx:9-35."""
    out = np.zeros((NO1, NO2))
    nB, _, nzg, _ = Kgrid.shape
    keep = [_l(b1), _d(fz1), _l(p1s), _l(b2), _d(fz2), _l(p2s), _d(fzGrid), _d(Kgrid)]
    lib().orc_kernel_parts_interp(int(NO1), int(NO2), out.ctypes.data_as(_dp),
                                  *[k[1] for k in keep], ctypes.c_long(nB), ctypes.c_long(nzg))
    Kinterp[...] = out


# ------------------------------------------------------------------------------------------------
# NumPy restatements of the Python-level algebra
# ------------------------------------------------------------------------------------------------
def bilininterp_precomputedbins(numBands, nobj, Kinterp, v1s, v2s, p1s, p2s, grid1, grid2, Kgrid):
    """delight/utils_cy.pyx:25-50 (NumPy restatement)."""
    p1, p2 = np.asarray(p1s[:nobj]), np.asarray(p2s[:nobj])
    v1, v2 = np.asarray(v1s[:nobj]), np.asarray(v2s[:nobj])
    dzm2 = 1. / (grid1[p1 + 1] - grid1[p1]) / (grid2[p2 + 1] - grid2[p2])
    for b in range(numBands):
        G = Kgrid[b]
        Kinterp[b, :nobj] = dzm2 * (
            (grid1[p1 + 1] - v1) * (grid2[p2 + 1] - v2) * G[p1, p2]
            + (v1 - grid1[p1]) * (grid2[p2 + 1] - v2) * G[p1 + 1, p2]
            + (grid1[p1 + 1] - v1) * (v2 - grid2[p2]) * G[p1, p2 + 1]
            + (v1 - grid1[p1]) * (v2 - grid2[p2]) * G[p1 + 1, p2 + 1])


def approx_DL(z):
    """delight/utils.py:12-13."""
    return np.exp(30.5 * np.asarray(z, dtype=float) ** 0.04 - 21.7)


def zeta(z, g_AB=1.0):
    """sqrt of the kernel prefactor: Zprefac(z,z') = zeta(z) zeta(z')  (photoz_kernels.py:338-340)."""
    z = np.asarray(z, dtype=float)
    return (1.0 + z) / (np.sqrt(4 * np.pi * g_AB) * approx_DL(z))


def kernel_grids(amp, mu, sig, lines_mu, alpha_C, alpha_L, redshiftGridGP):
    """Photoz_kernel.construct_interpolators (photoz_kernels.py:342-416): KC/KL on the
    (nB,nB,nzGP,nzGP) grid and the per-band diagonals (nB,nzGP)."""
    nB, NC = amp.shape
    NL = len(lines_mu)
    nzg = len(redshiftGridGP)
    norms = np.sqrt(2 * np.pi) * np.sum(amp * sig, axis=1)
    fz = 1.0 + np.asarray(redshiftGridGP, dtype=float)
    KC = np.zeros((nB, nB, nzg, nzg))
    KL = np.zeros_like(KC)
    scratch = [np.zeros((nzg, nzg)) for _ in range(3)]
    for i in range(nB):
        for j in range(nB):
            kc = np.zeros((nzg, nzg))
            kl = np.zeros((nzg, nzg))
            for s in scratch:
                s[...] = 0
            kernelparts(nzg, nzg, NC, NL, alpha_C, alpha_L, amp, mu, sig, lines_mu, None, norms,
                        np.full(nzg, i), fz, np.full(nzg, j), fz, True, kl, kc, *scratch)
            KC[i, j], KL[i, j] = kc, kl
    KCd = np.zeros((nB, nzg))
    KLd = np.zeros((nB, nzg))
    for i in range(nB):
        d = [np.zeros(nzg) for _ in range(4)]
        kernelparts_diag(nzg, NC, NL, alpha_C, alpha_L, amp, mu, sig, lines_mu, None, norms,
                         np.full(nzg, i), fz, True, *d)
        KLd[i], KCd[i] = d[0], d[1]
    return dict(KC=KC, KL=KL, KCd=KCd, KLd=KLd, norms=norms, fzGrid=fz)


def _bilinear(G, fzGrid, b1, x1, b2, x2):
    p1 = np.zeros(len(x1), dtype=np.int64)
    p2 = np.zeros(len(x2), dtype=np.int64)
    find_positions(len(x1), len(fzGrid), x1, p1, fzGrid)
    find_positions(len(x2), len(fzGrid), x2, p2, fzGrid)
    out = np.zeros((len(x1), len(x2)))
    kernel_parts_interp(len(x1), len(x2), out, b1, x1, p1, b2, x2, p2, fzGrid, G)
    return out


def interp_lin_extrap(xg, yg, x):
    """scipy interp1d(kind='linear', fill_value='extrapolate') on a sorted grid."""
    xg = np.asarray(xg, dtype=float)
    yg = np.asarray(yg, dtype=float)
    x = np.atleast_1d(np.asarray(x, dtype=float))
    i = np.clip(np.searchsorted(xg, x, side="left") - 1, 0, len(xg) - 2)
    # scipy: searchsorted(x_new) then clip to [1, len-1]; lo = that - 1
    i = np.clip(np.searchsorted(xg, x), 1, len(xg) - 1) - 1
    slope = (yg[i + 1] - yg[i]) / (xg[i + 1] - xg[i])
    return slope * (x - xg[i]) + yg[i]


def scalefree_best_type(fluxes, fluxes_var, tmpl):
    """delight-learn.py:67-79 with utils.py:179-193: tmpl [nt, B] = template fluxes at the
    galaxy's z in its valid bands.  Returns (bestType, ell)."""
    w = np.where(fluxes / fluxes_var < 1e-6, 0.0, 1.0 / fluxes_var)
    FOT = np.sum(tmpl * fluxes * w, axis=1)
    FTT = np.sum(tmpl ** 2 * w, axis=1)
    FOO = np.dot(w, fluxes ** 2)
    chi2 = FOO - FOT ** 2 / FTT
    best = int(np.argmin(chi2 + FTT))            # utils.py:191 returns chi2 + FTT
    return best, (FOT / FTT)[best]


def gp_fit(grids, V_C, V_L, tmpl_table, zfine, bands, z, fluxes, fluxes_var):
    """Per-training-galaxy fit (photoz_gp.py:186-234, delight-learn.py:61-90).

    tmpl_table [nt, nB, nzfine]: template fluxes on the fine grid `zfine`
    (<sed>_fluxredshiftmod.txt); evaluated at z by linear interpolation with extrapolation.
    Returns dict(bestType, ell, L [B,B], D [B], beta [B], core flat array).
    """
    import scipy.linalg
    bands = np.asarray(bands, dtype=np.int64)
    B = len(bands)
    nt = tmpl_table.shape[0]
    tm = np.array([[interp_lin_extrap(zfine, tmpl_table[t, b], z)[0] for b in bands]
                   for t in range(nt)])
    best, ell = scalefree_best_type(fluxes, fluxes_var, tm)
    x = np.full(B, 1.0 + z)
    kc = _bilinear(grids["KC"], grids["fzGrid"], bands, x, bands, x)
    kl = _bilinear(grids["KL"], grids["fzGrid"], bands, x, bands, x)
    pref = zeta(z) ** 4 * ell * ell
    A = pref * (V_C * kc + V_L * kl) + np.diag(fluxes_var)
    L = scipy.linalg.cholesky(A, lower=True)
    hx = ell * tm[best]
    hx = np.where(np.isfinite(hx), hx, 0.0)
    D = fluxes - hx
    beta = scipy.linalg.cho_solve((L, True), D)
    core = np.zeros(nt + B * (B + 1) // 2 + B)
    core[best] = 1.0
    core[nt:nt + B * (B + 1) // 2] = L[np.tril_indices(B)]
    core[nt + B * (B + 1) // 2:] = D
    return dict(bestType=best, ell=ell, L=L, D=D, beta=beta, core=core, A=A)


def gp_predict_interp(grids, V_C, V_L, tmpl_table, zfine, redshiftGridGP, bands, z, ell,
                      bestType, L, D, redshiftGrid):
    """setCore + predictAndInterpolate (photoz_gp.py:236-347) -> model_mean, model_var [nz, nB]."""
    import scipy.linalg
    nB = grids["KC"].shape[0]
    zg = np.asarray(redshiftGridGP, dtype=float)
    nzg = len(zg)
    bands = np.asarray(bands, dtype=np.int64)
    B = len(bands)
    beta = scipy.linalg.cho_solve((L, True), D)
    bstar = np.repeat(np.arange(nB), nzg)
    xstar = np.tile(1.0 + zg, nB)
    x = np.full(B, 1.0 + z)
    kc = _bilinear(grids["KC"], grids["fzGrid"], bstar, xstar, bands, x)
    kl = _bilinear(grids["KL"], grids["fzGrid"], bstar, xstar, bands, x)
    zs = np.tile(zg, nB)
    k = (zeta(zs) ** 2)[:, None] * zeta(z) ** 2 * ell * ell * (V_C * kc + V_L * kl)
    v = scipy.linalg.cho_solve((L, True), k.T)
    kd = zeta(zs) ** 4 * ell * ell * (V_C * grids["KCd"].ravel() + V_L * grids["KLd"].ravel())
    var = kd - np.einsum("ij,ji->i", k, v)
    mf = ell * np.concatenate([interp_lin_extrap(zfine, tmpl_table[bestType, b], zg)
                               for b in range(nB)])
    mean = k @ beta + mf
    model_mean = np.zeros((len(redshiftGrid), nB))
    model_var = np.zeros_like(model_mean)
    for b in range(nB):
        model_mean[:, b] = interp_lin_extrap(zg, mean[b * nzg:(b + 1) * nzg], redshiftGrid)
        model_var[:, b] = interp_lin_extrap(zg, var[b * nzg:(b + 1) * nzg], redshiftGrid)
    return model_mean, model_var


def z_prior(redshiftGrid, train_z, zPriorSigma):
    """delight-apply.py:104."""
    return np.exp(-0.5 * ((redshiftGrid[:, None] - train_z[None, :]) / zPriorSigma) ** 2)


def template_prior(redshiftGrid, p_t, p_z_t):
    """scripts/templateFitting.py:79-82."""
    b2 = np.asarray(p_z_t, dtype=float) ** 2.0
    zz = redshiftGrid[:, None]
    return np.asarray(p_t)[None, :] * zz / b2[None, :] * np.exp(-0.5 * zz ** 2 / b2[None, :])


def metrics_core(ztrue, redshiftGrid, pdf):
    """First seven columns of computeMetrics (utils.py:213-227)."""
    zmean = np.average(redshiftGrid, weights=pdf)
    imap = int(np.argmax(pdf))
    zmap = redshiftGrid[imap]
    s1 = np.sqrt(np.average((redshiftGrid - zmean) ** 2, weights=pdf))
    s2 = np.sqrt(np.average((redshiftGrid - zmap) ** 2, weights=pdf))
    return np.array([ztrue, zmean, s1, zmap, s2, np.interp(ztrue, redshiftGrid, pdf),
                     np.interp(ztrue, redshiftGrid, pdf.cumsum())]), imap


def ci_level(redshiftGrid, pdf, fraction, numlevels=200):
    """utils.py:196-205."""
    def trapz(y, x):
        return np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5) if len(x) > 1 else 0.0
    ev = trapz(pdf, redshiftGrid)
    for level in np.linspace(0, pdf.max(), num=numlevels):
        ind = np.where(pdf <= level)
        if trapz(pdf[ind], redshiftGrid[ind]) >= fraction * ev:
            return level
    return None


def gp_predict_interp_hp(grids, V_C, V_L, tmpl_table, zfine, redshiftGridGP, bands, z, ell,
                         bestType, L, D, redshiftGrid):
    """Higher-precision truth for gp_predict_interp: the same formulas (photoz_gp.py:236-347,
    SURVEY Appendix C) on the same float64 inputs, every operation in np.longdouble (x86: 64-bit
    mantissa, eps 1.1e-19) with explicit substitutions instead of LAPACK.  The model variance
    `Kdiag - k^T A^-1 k` is a cancelling difference; this function is the yardstick that shows how
    much the float64 implementations actually lose there: on the golden mock the reference's
    cho_solve + dot is 1.7e-14 away, the GPU's |L^-1 k|^2 form 1.1e-14 (tests/test_oracle.py,
    tools/var_check.py)."""
    ld = np.longdouble
    nB = grids["KC"].shape[0]
    zg = np.asarray(redshiftGridGP, dtype=float)
    nzg = len(zg)
    bands = np.asarray(bands, dtype=np.int64)
    B = len(bands)
    fz = grids["fzGrid"].astype(ld)
    KC, KL = grids["KC"], grids["KL"]

    def zeta_hp(zz):
        zz = ld(zz)
        dl = np.exp(ld(30.5) * np.power(zz, ld(0.04)) - ld(21.7))
        return (ld(1) + zz) / (np.sqrt(ld(4) * ld(np.pi) * ld(1.0)) * dl)

    def findpos(x):
        x = float(x)
        g = grids["fzGrid"]
        if not (x >= g[0] and x <= g[-1]):
            return 0
        for p in range(len(g) - 1):
            if g[p] <= x <= g[p + 1]:
                return p
        return 0

    def bilin(G, p1, x, p2, y):
        xl, xh, yl, yh = fz[p1], fz[p1 + 1], fz[p2], fz[p2 + 1]
        x, y = ld(x), ld(y)
        inv = ld(1) / (xh - xl) / (yh - yl)
        return inv * ((xh - x) * (yh - y) * ld(G[p1, p2]) + (x - xl) * (yh - y) * ld(G[p1 + 1, p2])
                      + (xh - x) * (y - yl) * ld(G[p1, p2 + 1])
                      + (x - xl) * (y - yl) * ld(G[p1 + 1, p2 + 1]))

    Lh = np.asarray(L, dtype=float).astype(ld)
    Dh = np.asarray(D, dtype=float).astype(ld)

    def fwd(b):
        w = np.zeros(B, dtype=ld)
        for j in range(B):
            s = b[j]
            for m in range(j):
                s = s - Lh[j, m] * w[m]
            w[j] = s / Lh[j, j]
        return w

    def bwd(w):
        x = np.zeros(B, dtype=ld)
        for j in range(B - 1, -1, -1):
            s = w[j]
            for m in range(j + 1, B):
                s = s - Lh[m, j] * x[m]
            x[j] = s / Lh[j, j]
        return x

    beta = bwd(fwd(Dh))
    ellh = ld(ell)
    zz = zeta_hp(z)
    x = 1.0 + z
    p2 = findpos(x)
    mean = np.zeros((nB, nzg), dtype=ld)
    var = np.zeros((nB, nzg), dtype=ld)
    VC, VL = ld(V_C), ld(V_L)
    for bs in range(nB):
        for iz in range(nzg):
            xs = 1.0 + zg[iz]
            p1 = findpos(xs)
            zs = zeta_hp(zg[iz])
            pref = zs * zs * zz * zz * ellh * ellh
            k = np.zeros(B, dtype=ld)
            for j in range(B):
                bj = int(bands[j])
                k[j] = pref * (VC * bilin(KC[bs, bj], p1, xs, p2, x)
                               + VL * bilin(KL[bs, bj], p1, xs, p2, x))
            v = bwd(fwd(k))
            kd = zs ** 4 * ellh * ellh * (VC * ld(grids["KCd"][bs, iz]) + VL * ld(grids["KLd"][bs, iz]))
            var[bs, iz] = kd - np.dot(k, v)
            tab = tmpl_table[bestType, bs]
            zf = np.asarray(zfine, dtype=float)
            i = int(np.clip(np.searchsorted(zf, zg[iz]), 1, len(zf) - 1) - 1)
            slope = (ld(tab[i + 1]) - ld(tab[i])) / (ld(zf[i + 1]) - ld(zf[i]))
            mean[bs, iz] = np.dot(k, beta) + ellh * (slope * (ld(zg[iz]) - ld(zf[i])) + ld(tab[i]))
    rg = np.asarray(redshiftGrid, dtype=float)
    ii = np.clip(np.searchsorted(zg, rg), 1, nzg - 1) - 1
    zgl, rgl = zg.astype(ld), rg.astype(ld)
    model_mean = np.zeros((len(rg), nB), dtype=ld)
    model_var = np.zeros_like(model_mean)
    for b in range(nB):
        for arr, out in ((mean, model_mean), (var, model_var)):
            slope = (arr[b, ii + 1] - arr[b, ii]) / (zgl[ii + 1] - zgl[ii])
            out[:, b] = slope * (rgl - zgl[ii]) + arr[b, ii]
    return model_mean, model_var

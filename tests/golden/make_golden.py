#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference here.

Needs /root/reference (data files) and oracle/_ref (python oracle/build_ref.py).  Run from the
repo root in the build container:

    python tests/golden/make_golden.py

Writes
    tests/golden/like_small.npz        approx_flux_likelihood_cy + NumPy twin, random inputs
    tests/golden/kernelparts_small.npz kernelparts / kernelparts_diag / kernel_parts_interp /
                                       find_positions on random inputs (reference test shapes)
    tests/golden/sdss_pipeline.npz     seeded mock -> reference learn -> predict -> apply epilogue
                                       (5 SDSS bands, training on 3, heterogeneous targets)
    tests/golden/templatefit.npz       templateFitting.py-equivalent (config 2 twin)
    tests/golden/cv_small.npz          the CV file the real delight-learn.py writes with
                                       crossValidate on (training catalogue of sdss_pipeline.npz;
                                       `python tests/golden/make_golden.py cv`, not in the default
                                       list so that the other fixtures stay byte-identical)
    tests/golden/compressed_small.npz  second ("compressed") pass of delight-apply.py:123-135:
                                       the reference's NumPy twin on the stored top-K columns
    delight_b200/data/lsst_cww.npz     LSST ugrizy Gaussian coefficients + CWW template flux
                                       tables on the 1500-point z grid (outputs of the
                                       reference's processFilters/processSEDs), used by bench.py
The NumPy global RNG is seeded explicitly everywhere (the reference scripts are unseeded).
"""
import os
import sys
import tempfile
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

from oracle import ref, pipeline  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def stable_topk(evid, K):
    idx = np.argsort(evid, kind="stable")[::-1][:K]
    return idx, evid[idx]


def gen_like_small():
    from delight.utils_cy import approx_flux_likelihood_cy
    from delight.utils import approx_flux_likelihood
    rng = np.random.default_rng(1234)
    out = {}
    cases = [("a", 30, 20, 5), ("b", 7, 3, 1), ("c", 1, 1, 6), ("d", 64, 33, 6), ("e", 5, 40, 3)]
    for name, nz, nt, nf in cases:
        f_obs = rng.uniform(1, 2, nf)
        f_var = rng.uniform(.1, .2, nf)
        f_mod = rng.uniform(1, 2, (nz, nt, nf))
        f_cov = rng.uniform(.1, .2, (nz, nt, nf))
        if name == "d":      # pipeline-like dynamic range + z-dependent ell prior
            f_obs = rng.uniform(1, 2, nf) * 4e-5
            f_var = (0.05 * f_obs) ** 2
            f_mod = rng.uniform(0.5, 2, (nz, nt, nf)) * 1e5 * rng.uniform(.1, 10, (nz, nt, 1))
            f_cov = (f_mod * rng.uniform(1e-4, 1e-2, (nz, nt, nf))) ** 2
            ell_hat = rng.uniform(.5, 2, nz)
            ell_var = rng.uniform(.1, .5, nz)
        else:
            ell_hat = np.ones(nz)
            ell_var = np.full(nz, 0.01 if name != "e" else 1e24)
        like = np.zeros((nz, nt))
        approx_flux_likelihood_cy(like, nz, nt, nf, f_obs, f_var, f_mod, f_cov, ell_hat, ell_var)
        twin = approx_flux_likelihood(f_obs, f_var, f_mod, f_mod_covar=f_cov,
                                      ell_hat=ell_hat[:, None], ell_var=ell_var[:, None],
                                      normalized=True, marginalizeEll=True, renormalize=True)
        for k, v in dict(f_obs=f_obs, f_var=f_var, f_mod=f_mod, f_cov=f_cov, ell_hat=ell_hat,
                         ell_var=ell_var, like=like, twin=twin).items():
            out[name + "_" + k] = v
    np.savez_compressed(os.path.join(GOLD, "like_small.npz"), **out)
    print("like_small: max |cy/twin-1| =",
          max(np.abs(out[c[0] + "_like"] / out[c[0] + "_twin"] - 1).max() for c in cases))


def gen_compressed_small():
    """delight-apply.py:123-135,145-146 with the computed array actually used (the script keeps
    the stale `like_grid`) and ell broadcast over the training axis (the script's 1-D arrays do
    not broadcast against [nz, K])."""
    from delight.utils import approx_flux_likelihood
    rng = np.random.default_rng(4321)
    nz, nT, nB, nO, K = 40, 30, 5, 6, 8
    zgrid = np.sort(rng.uniform(0.01, 3.0, nz))
    model_mean = rng.uniform(0.5, 2, (nz, nT, nB)) * 1e5 * rng.uniform(.1, 10, (1, nT, 1))
    model_covar = (model_mean * rng.uniform(1e-4, 1e-2, (nz, nT, nB))) ** 2
    train_z = rng.uniform(0.01, 3.0, nT)
    prior = np.exp(-0.5 * ((zgrid[:, None] - train_z[None, :]) / 0.5) ** 2)
    nf = np.array([5, 5, 3, 4, 2, 5])
    bands = np.zeros((nO, nB), dtype=np.int64)
    fluxes = np.zeros((nO, nB))
    fluxes_var = np.ones((nO, nB))
    for o in range(nO):
        b = np.sort(rng.choice(nB, nf[o], replace=False))
        bands[o, :nf[o]] = b
        t, iz = rng.integers(nT), rng.integers(nz)
        f = model_mean[iz, t, b] * 4e-10 * rng.uniform(0.9, 1.1, nf[o])
        fluxes[o, :nf[o]] = f
        fluxes_var[o, :nf[o]] = (0.05 * f) ** 2
    indices = np.stack([rng.choice(nT, K, replace=False) for _ in range(nO)]).astype(np.int64)
    indices[1, 3] = indices[1, 0]          # duplicate entry: the in1d mask keeps one column
    indices[2, 5] = nT + 7                 # not a training index: ignored by the mask
    ell_hat = rng.uniform(.5, 2, nz)
    ell_var = rng.uniform(.1, .5, nz)
    out = dict(zgrid=zgrid, model_mean=model_mean, model_covar=model_covar, prior=prior, nf=nf,
               bands=bands, fluxes=fluxes, fluxes_var=fluxes_var, indices=indices,
               ell_hat=ell_hat, ell_var=ell_var)
    targetIndices = np.arange(nT)
    for tag, normalized, eh, ev in (("n0", False, ell_hat, ell_var), ("n1", True, ell_hat, ell_var),
                                    ("s0", False, np.ones(nz), np.full(nz, 0.25))):
        pdfs = np.zeros((nO, nz))
        evid = np.zeros((nO, K))
        sel_sorted = -np.ones((nO, K), dtype=np.int64)
        for o in range(nO):
            b = bands[o, :nf[o]]
            sel = np.in1d(targetIndices, indices[o], assume_unique=True)       # :125
            like_grid = approx_flux_likelihood(                                 # :126-134
                fluxes[o, :nf[o]], fluxes_var[o, :nf[o]], model_mean[:, sel, :][:, :, b],
                f_mod_covar=model_covar[:, sel, :][:, :, b], marginalizeEll=True,
                normalized=normalized, ell_hat=eh[:, None], ell_var=ev[:, None])
            like_grid *= prior[:, sel]                                          # :135
            pdfs[o] = like_grid.sum(axis=1)                                     # :145
            e = np.trapz(like_grid, x=zgrid, axis=0)                            # :146
            evid[o, :e.size] = e
            sel_sorted[o, :e.size] = targetIndices[sel]
        out["pdf_" + tag], out["evid_" + tag], out["sel_" + tag] = pdfs, evid, sel_sorted
    np.savez_compressed(os.path.join(GOLD, "compressed_small.npz"), **out)
    print("compressed_small done")


def gen_kernelparts_small(amp, mu, sig):
    from delight.photoz_kernels_cy import kernelparts, kernelparts_diag, kernel_parts_interp
    from delight.utils_cy import find_positions
    rng = np.random.default_rng(99)
    nB, NC = amp.shape
    norms = np.sqrt(2 * np.pi) * np.sum(amp * sig, axis=1)
    lines_mu = np.array([6500, 5002.26, 3732.22])
    lines_sig = np.array([20.0, 20.0, 20.0])
    NL = 3
    aC, aL = 1e3, 1e2
    out = dict(amp=amp, mu=mu, sig=sig, norms=norms, lines_mu=lines_mu, lines_sig=lines_sig,
               alpha_C=aC, alpha_L=aL)
    NO1, NO2 = 13, 9
    b1 = rng.integers(0, nB, NO1)
    b2 = rng.integers(0, nB, NO2)
    fz1 = 1 + rng.uniform(0, 3, NO1)
    fz2 = 1 + rng.uniform(0, 3, NO2)
    outs = [np.zeros((NO1, NO2)) for _ in range(5)]
    kernelparts(NO1, NO2, NC, NL, aC, aL, amp, mu, sig, lines_mu, lines_sig, norms,
                b1, fz1, b2, fz2, True, *outs)
    out.update(b1=b1, b2=b2, fz1=fz1, fz2=fz2, KL=outs[0].copy(), KC=outs[1].copy(),
               DaC=outs[2].copy(), DaL=outs[3].copy(), Daz=outs[4].copy())
    # second call on the now non-zero buffers: pins the accumulate-then-normalise contract
    kernelparts(NO1, NO2, NC, NL, aC, aL, amp, mu, sig, lines_mu, lines_sig, norms,
                b1, fz1, b2, fz2, True, *outs)
    out.update(KL2=outs[0].copy(), KC2=outs[1].copy(), DaC2=outs[2].copy(),
               DaL2=outs[3].copy(), Daz2=outs[4].copy())
    # grad_needed False leaves gradients alone
    outs = [np.full((NO1, NO2), 0.5) for _ in range(5)]
    kernelparts(NO1, NO2, NC, NL, aC, aL, amp, mu, sig, lines_mu, lines_sig, norms,
                b1, fz1, b2, fz2, False, *outs)
    out.update(KL_ng=outs[0].copy(), KC_ng=outs[1].copy(), DaC_ng=outs[2].copy())
    # NL = 0
    outs = [np.zeros((NO1, NO2)) for _ in range(5)]
    kernelparts(NO1, NO2, NC, 0, aC, aL, amp, mu, sig, lines_mu[:0], lines_sig[:0], norms,
                b1, fz1, b2, fz2, True, *outs)
    out.update(KL_nl0=outs[0].copy(), KC_nl0=outs[1].copy())
    # diag
    d = [np.full(NO1, 0.25) for _ in range(4)]
    kernelparts_diag(NO1, NC, NL, aC, aL, amp, mu, sig, lines_mu, lines_sig, norms,
                     b1, fz1, True, *d)
    out.update(KLd=d[0], KCd=d[1], DaCd=d[2], DaLd=d[3])
    # reference golden vector (tests/test_photoz_kernels_cy.py:62-68)
    a = np.array([0., 1., 2., 3., 4.])
    b = np.array([0.5, 2.5, 3.0, 3.1, 4.0])
    pos = np.zeros(b.size, dtype=np.int64)
    find_positions(b.size, a.size, b, pos, a)
    assert list(pos) == [0, 2, 2, 3, 3]
    out.update(fp_grid=a, fp_vals=b, fp_pos=pos)
    # out-of-range values stay untouched
    vals = np.array([-1.0, 0.0, 4.0, 5.0, 2.0])
    pos2 = np.full(vals.size, 7, dtype=np.int64)
    find_positions(vals.size, a.size, vals, pos2, a)
    out.update(fp_vals2=vals, fp_pos2=pos2)
    # interp on a small grid
    nzg = 11
    fzg = 1 + np.logspace(np.log10(0.01), np.log10(3.03), nzg)
    Kgrid = rng.uniform(0, 1, (nB, nB, nzg, nzg))
    x1 = rng.uniform(fzg[0], fzg[-1], NO1)
    x2 = rng.uniform(fzg[0], fzg[-1], NO2)
    p1 = np.zeros(NO1, dtype=np.int64)
    p2 = np.zeros(NO2, dtype=np.int64)
    find_positions(NO1, nzg, x1, p1, fzg)
    find_positions(NO2, nzg, x2, p2, fzg)
    Ki = np.zeros((NO1, NO2))
    kernel_parts_interp(NO1, NO2, Ki, b1, x1, p1, b2, x2, p2, fzg, Kgrid)
    out.update(ki_fzg=fzg, ki_grid=Kgrid, ki_x1=x1, ki_x2=x2, ki_p1=p1, ki_p2=p2, ki_out=Ki)
    np.savez_compressed(os.path.join(GOLD, "kernelparts_small.npz"), **out)
    print("kernelparts_small done")


def gen_sdss_pipeline(wd):
    """Seeded C1-shaped mock through the reference's learn -> predict -> apply sequence,
    captured in memory (scripts/delight-learn.py:61-90, scripts/delight-apply.py:77-170)."""
    from delight.io import parseParamFile, readBandCoefficients, createGrids, readSEDs, \
        getDataFromFile
    from delight.utils import scalefree_flux_likelihood, computeMetrics
    from delight.utils_cy import approx_flux_likelihood_cy
    from delight.photoz_gp import PhotozGP

    nT, nO = 20, 24   # nO >= nT: delight-apply.py:79-80 caps the training set at numObjectsTarget
    cfg = pipeline.make_workdir(wd, bands=pipeline.SDSS,
                                train_bands=["U_SDSS", "G_SDSS", "I_SDSS"], refband="I_SDSS",
                                numObjects=max(nT, nO), dz=0.01, nzGP=40, Ncompress=8)
    pipeline.run_script("processFilters", cfg)
    pipeline.run_script("processSEDs", cfg)
    pipeline.run_script("simulateWithSEDs", cfg, seed=0)
    params = parseParamFile(cfg, verbose=False)
    # make the catalogues heterogeneous: knock out bands (io.py:331-337 drops them)
    rng = np.random.default_rng(5)
    tr = np.loadtxt(params["training_catFile"])[:nT]
    tr[3, 1] = -1.0            # G flux invalid -> 2 bands left
    tr[7, 0] = np.nan          # U invalid
    np.savetxt(params["training_catFile"], tr)
    ta = np.loadtxt(params["target_catFile"])[:nO]
    ta[2, 2] = -5.0            # R missing
    ta[5, 0] = 0.0             # U missing
    ta[5, 4] = -1.0            # Z missing
    ta[9, 8] = np.nan          # I variance invalid... I is the reference band flux column 3
    np.savetxt(params["target_catFile"], ta)

    amp, mu, sig, norms = readBandCoefficients(params)
    _, zgrid, zgridGP = createGrids(params)
    f_mod = readSEDs(params)
    nt = f_mod.shape[0]
    nB = amp.shape[0]
    nz = zgrid.size
    tmpl = np.stack([np.loadtxt(os.path.join(params["templates_directory"],
                                             s + "_fluxredshiftmod.txt")).T
                     for s in params["templates_names"]])          # [nt, nB, nz]
    gp = PhotozGP(f_mod, amp, mu, sig, params["lines_pos"], params["lines_width"],
                  params["V_C"], params["V_L"], params["alpha_C"], params["alpha_L"],
                  zgridGP, use_interpolators=True)
    out = dict(amp=amp, mu=mu, sig=sig, norms=norms, zgrid=zgrid, zgridGP=zgridGP, tmpl=tmpl,
               lines_pos=np.array(params["lines_pos"]), V_C=params["V_C"], V_L=params["V_L"],
               alpha_C=params["alpha_C"], alpha_L=params["alpha_L"],
               zPriorSigma=params["zPriorSigma"], ellPriorSigma=params["ellPriorSigma"],
               extraFrac=params["training_extraFracFluxError"],
               KC_grid=gp.kernel.KC_grid, KL_grid=gp.kernel.KL_grid,
               KCd=np.stack([gp.kernel.KC_diag_interp[b].y for b in range(nB)]),
               KLd=np.stack([gp.kernel.KL_diag_interp[b].y for b in range(nB)]))

    # ---- learn (delight-learn.py:61-90)
    numCol = 3 + nB + nB * (nB + 1) // 2 + nB + nt
    rows = np.zeros((nT, numCol))
    tr_bands = np.full((nT, nB), -1, dtype=np.int64)
    tr_nb = np.zeros(nT, dtype=np.int64)
    tr_flux = np.zeros((nT, nB))
    tr_var = np.zeros((nT, nB))
    tr_z = np.zeros(nT)
    it = getDataFromFile(params, 0, nT, prefix="training_", getXY=True, CV=False)
    for loc, (z, nrf, bands, fluxes, fluxesVar, _, _, _, X, Y, Yvar) in enumerate(it):
        themod = np.zeros((1, nt, bands.size))
        for t in range(nt):
            for ib, band in enumerate(bands):
                themod[0, t, ib] = f_mod[t, band](z)
        chi2_grid, ellMLs = scalefree_flux_likelihood(fluxes, fluxesVar, themod, returnChi2=True)
        bestType = np.argmin(chi2_grid)
        ell = ellMLs[0, bestType]
        X[:, 2] = ell
        gp.setData(X, Y, Yvar, bestType)
        lB = bands.size
        rows[loc, 0] = lB
        rows[loc, 1] = z
        rows[loc, 2] = ell
        rows[loc, 3:3 + lB] = bands
        rows[loc, 3 + lB:3 + nt + lB + lB * (lB + 1) // 2 + lB] = gp.getCore()
        tr_nb[loc] = lB
        tr_bands[loc, :lB] = bands
        tr_flux[loc, :lB] = fluxes
        tr_var[loc, :lB] = fluxesVar
        tr_z[loc] = z
    assert loc == nT - 1
    out.update(gp_rows=rows, tr_nb=tr_nb, tr_bands=tr_bands, tr_flux=tr_flux, tr_var=tr_var,
               tr_z=tr_z)

    # cross-check with the real script's text output (written with %.12e)
    pipeline.run_script("delight-learn", cfg)
    txt = np.loadtxt(params["training_paramFile"])
    assert np.allclose(txt, rows, rtol=2e-12, atol=1e-300), np.abs(txt - rows).max()

    # ---- model grid (delight-apply.py:89-98)
    model_mean = np.zeros((nz, nT, nB))
    model_covar = np.zeros((nz, nT, nB))
    for loc in range(nT):
        B = int(rows[loc, 0])
        z, ell = rows[loc, 1], rows[loc, 2]
        bands = rows[loc, 3:3 + B]
        flat = rows[loc, 3 + B:]
        X = np.zeros((B, 3))
        X[:, 0] = bands
        X[:, 1] = z
        X[:, 2] = ell
        gp.setCore(X, B, nt, flat[0:nt + B + B * (B + 1) // 2])
        model_mean[:, loc, :], model_covar[:, loc, :] = gp.predictAndInterpolate(zgrid, ell=ell)
    out.update(model_mean=model_mean, model_covar=model_covar)

    # ---- apply (delight-apply.py:104-170)
    K = params["Ncompress"]
    prior = np.exp(-0.5 * ((zgrid[:, None] - rows[:, 1][None, :]) / params["zPriorSigma"]) ** 2)
    pdfs = np.zeros((nO, nz))
    evid = np.zeros((nO, nT))
    top_idx = np.zeros((nO, K), dtype=np.int64)
    top_ev = np.zeros((nO, K))
    metrics = np.zeros((nO, 7 + len(params["confidenceLevels"])))
    ta_bands = np.full((nO, nB), -1, dtype=np.int64)
    ta_nb = np.zeros(nO, dtype=np.int64)
    ta_flux = np.zeros((nO, nB))
    ta_var = np.zeros((nO, nB))
    ta_z = np.zeros(nO)
    like0 = None
    it = getDataFromFile(params, 0, nO, prefix="target_", getXY=False, CV=False)
    n_read = 0
    for loc, (z, nrf, bands, fluxes, fluxesVar, _, _, _) in enumerate(it):
        ell_hat_z = np.ones(nz)
        like_grid = np.zeros((nz, nT))
        approx_flux_likelihood_cy(like_grid, nz, nT, bands.size, fluxes, fluxesVar,
                                  model_mean[:, :, bands], model_covar[:, :, bands],
                                  ell_hat=ell_hat_z,
                                  ell_var=(ell_hat_z * params["ellPriorSigma"]) ** 2)
        if loc == 0:
            like0 = like_grid.copy()
        like_grid *= prior
        pdfs[loc] = like_grid.sum(axis=1)
        ev = np.trapz(like_grid, x=zgrid, axis=0)
        evid[loc] = ev
        top_idx[loc], top_ev[loc] = stable_topk(ev, K)
        metrics[loc] = computeMetrics(z, zgrid, pdfs[loc], params["confidenceLevels"])
        ta_nb[loc] = bands.size
        ta_bands[loc, :bands.size] = bands
        ta_flux[loc, :bands.size] = fluxes
        ta_var[loc, :bands.size] = fluxesVar
        ta_z[loc] = z
        n_read += 1
    for a in (pdfs, evid, top_idx, top_ev, metrics, ta_bands, ta_nb, ta_flux, ta_var, ta_z):
        a.resize((n_read,) + a.shape[1:], refcheck=False)
    out.update(train_cat=tr, target_cat=ta,
               train_bandOrder=np.array(" ".join(params["training_bandOrder"])),
               target_bandOrder=np.array(" ".join(params["target_bandOrder"])),
               Ncompress=K)
    out.update(prior=prior, pdfs=pdfs, evid=evid, top_idx=top_idx, top_ev=top_ev,
               metrics=metrics, ta_nb=ta_nb, ta_bands=ta_bands, ta_flux=ta_flux, ta_var=ta_var,
               ta_z=ta_z, like0=like0, confidenceLevels=np.array(params["confidenceLevels"]))

    # cross-check with the real script's outputs (%.2e text)
    pipeline.run_script("delight-apply", cfg)
    txt = np.loadtxt(params["redshiftpdfFile"])
    assert np.allclose(txt, pdfs, rtol=6e-3, atol=1e-300), np.abs(txt / pdfs - 1).max()
    np.savez_compressed(os.path.join(GOLD, "sdss_pipeline.npz"), **out)
    print("sdss_pipeline: nT=%d nO=%d nz=%d; script outputs agree" % (nT, n_read, nz))
    return amp, mu, sig


def gen_templatefit(wd):
    """Config-2 twin: scripts/templateFitting.py:60-89 captured in memory (SDSS cfg above)."""
    from delight.io import parseParamFile, createGrids, getDataFromFile
    from delight.utils import approx_flux_likelihood, computeMetrics
    cfg = os.path.join(wd, "params.cfg")
    params = parseParamFile(cfg, verbose=False)
    _, zgrid, _ = createGrids(params)
    nz = zgrid.size
    f_mod = np.zeros((nz, len(params["templates_names"]), len(params["bandNames"])))
    for t, s in enumerate(params["templates_names"]):
        f_mod[:, t, :] = np.loadtxt(os.path.join(params["templates_directory"],
                                                 s + "_fluxredshiftmod.txt"))
    nO = 12
    pdfs = np.zeros((nO, nz))
    metrics = np.zeros((nO, 7 + len(params["confidenceLevels"])))
    likes = []
    it = getDataFromFile(params, 0, nO, prefix="target_", getXY=False)
    for loc, (z, nrf, bands, fluxes, fluxesVar, _, _, _) in enumerate(it):
        ell_hat_z = 1
        ellPriorSigma = 1e12
        like_grid = approx_flux_likelihood(fluxes, fluxesVar, f_mod[:, :, bands],
                                           normalized=True, marginalizeEll=True,
                                           ell_hat=ell_hat_z,
                                           ell_var=(ell_hat_z * ellPriorSigma) ** 2)
        b_in = np.array(params["p_t"])[None, :]
        beta2 = np.array(params["p_z_t"]) ** 2.0
        p_z = b_in * zgrid[:, None] / beta2[None, :] * \
            np.exp(-0.5 * zgrid[:, None] ** 2 / beta2[None, :])
        if loc == 0:
            likes = like_grid.copy()
        like_grid *= p_z
        pdfs[loc] = like_grid.sum(axis=1)
        metrics[loc] = computeMetrics(z, zgrid, pdfs[loc], params["confidenceLevels"])
    n = loc + 1
    np.savez_compressed(os.path.join(GOLD, "templatefit.npz"), f_mod=f_mod, p_z=p_z,
                        pdfs=pdfs[:n], metrics=metrics[:n], like0=likes,
                        p_t=np.array(params["p_t"]), p_z_t=np.array(params["p_z_t"]))
    print("templatefit: %d targets" % n)


def gen_cv(wd):
    """Cross-validation of delight-learn (delight-learn.py:50-57,92-100,131-132): the real script,
    crossValidate on, on the training catalogue of sdss_pipeline.npz (training on U, G, I;
    held-out bands R, Z) -> the CV file it writes."""
    g = np.load(os.path.join(GOLD, "sdss_pipeline.npz"))
    cfg = pipeline.make_workdir(wd, bands=pipeline.SDSS,
                                train_bands=["U_SDSS", "G_SDSS", "I_SDSS"], refband="I_SDSS",
                                numObjects=24, dz=0.01, nzGP=40, Ncompress=8, copy_data=False)
    pipeline.write_preprocessed(wd, pipeline.SDSS, g["amp"], g["mu"], g["sig"], g["tmpl"])
    # the simulator leaves the held-out columns at 0 (and the script then fails on the empty band
    # list): give R and Z deterministic fluxes; one galaxy keeps a missing R
    tr = g["train_cat"].copy()
    tr[:, 2] = np.sqrt(np.abs(tr[:, 1] * tr[:, 3])) * (1 + 0.03 * np.sin(np.arange(tr.shape[0])))
    tr[:, 4] = 1.1 * tr[:, 3] * (1 + 0.03 * np.cos(np.arange(tr.shape[0])))
    tr[:, 7] = (0.05 * tr[:, 2]) ** 2
    tr[:, 9] = (0.05 * tr[:, 4]) ** 2
    tr[3, 2] = np.sqrt(tr[3, 0] * tr[3, 3])       # row 3 has G = -1
    tr[3, 7] = (0.05 * tr[3, 2]) ** 2
    tr[5, 2] = 0.0                                # R missing for one galaxy
    np.savetxt(wd + "/galaxies-fluxredshifts.txt", tr)
    np.savetxt(wd + "/galaxies-fluxredshifts2.txt", g["target_cat"])
    txt = open(cfg).read().replace("crossValidate: False", "crossValidate: True")
    open(cfg, "w").write(txt)
    pipeline.run_script("delight-learn", cfg)
    cv = np.loadtxt(wd + "/galaxies-gpCV.txt", ndmin=2)
    rows = np.loadtxt(wd + "/galaxies-gpparams.txt")
    assert np.allclose(rows, g["gp_rows"], rtol=2e-12, atol=1e-300)
    assert cv.shape == (g["train_cat"].shape[0], 2) and np.all(cv <= 0) and np.any(cv < 0)
    np.savez_compressed(os.path.join(GOLD, "cv_small.npz"), cv=cv, train_cat=tr)
    print("cv_small:", cv.shape, "min", cv.min())


def gen_lsst_data(wd):
    """LSST ugrizy coefficients + CWW template tables on the 1500-point grid (bench inputs)."""
    from delight.io import parseParamFile, readBandCoefficients, createGrids
    cfg = pipeline.make_workdir(wd, bands=pipeline.LSST, refband="lsst_i", zmin=0.01, zmax=3.01,
                                dz=0.002, nzGP=150)
    pipeline.run_script("processFilters", cfg)
    pipeline.run_script("processSEDs", cfg)
    params = parseParamFile(cfg, verbose=False, catFilesNeeded=False)
    amp, mu, sig, norms = readBandCoefficients(params)
    _, zgrid, zgridGP = createGrids(params)
    tmpl = np.stack([np.loadtxt(os.path.join(params["templates_directory"],
                                             s + "_fluxredshiftmod.txt")).T
                     for s in params["templates_names"]])
    assert zgrid.size == 1500 and tmpl.shape == (8, 6, 1500)
    dst = os.path.join(REPO, "delight_b200", "data")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "lsst_cww.npz"), amp=amp, mu=mu, sig=sig, norms=norms,
                        zgrid=zgrid, zgridGP=zgridGP, tmpl=tmpl,
                        band_names=np.array(pipeline.LSST), sed_names=np.array(pipeline.CWW),
                        lines_pos=np.array(params["lines_pos"]),
                        p_t=np.array(params["p_t"]), p_z_t=np.array(params["p_z_t"]))
    print("lsst_cww done", tmpl.shape)


if __name__ == "__main__":
    ref.load()
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["like", "comp", "sdss", "lsst"]
    if "like" in which:
        gen_like_small()
    if "comp" in which:
        gen_compressed_small()
    if "sdss" in which:
        with tempfile.TemporaryDirectory(prefix="delight_gold_") as wd:
            amp, mu, sig = gen_sdss_pipeline(wd)
            gen_templatefit(wd)
            gen_kernelparts_small(amp, mu, sig)
    if "cv" in which:
        with tempfile.TemporaryDirectory(prefix="delight_gold_") as wd:
            gen_cv(wd)
    if "lsst" in which:
        with tempfile.TemporaryDirectory(prefix="delight_gold_") as wd:
            gen_lsst_data(wd)

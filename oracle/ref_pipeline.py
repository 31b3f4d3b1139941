"""Reference-side pipeline pieces for benchmarks and tests -- TEST INFRASTRUCTURE ONLY.

Drives the UNMODIFIED reference classes of oracle/_ref (delight.photoz_gp.PhotozGP,
delight.utils.scalefree_flux_likelihood) exactly as scripts/delight-learn.py:61-90 and
scripts/delight-apply.py:77-98 do, on in-memory catalogues (no cfg / text files).
"""
import numpy as np

from . import ref


def make_gp(data, hyper):
    """PhotozGP(...) as constructed by the scripts (delight-learn.py:38-42)."""
    ref.load()
    from scipy.interpolate import interp1d
    from delight.photoz_gp import PhotozGP
    tmpl, zgrid = data["tmpl"], data["zgrid"]
    nt, nB, _ = tmpl.shape
    f_mod = np.zeros((nt, nB), dtype=object)
    for t in range(nt):                       # io.readSEDs, io.py:263-277
        for b in range(nB):
            f_mod[t, b] = interp1d(zgrid, tmpl[t, b], kind="linear", bounds_error=False,
                                   fill_value="extrapolate")
    gp = PhotozGP(f_mod, data["amp"], data["mu"], data["sig"], hyper["lines_pos"],
                  hyper["lines_width"], hyper["V_C"], hyper["V_L"], hyper["alpha_C"],
                  hyper["alpha_L"], data["zgridGP"], use_interpolators=True)
    return gp, f_mod


def learn(gp, f_mod, cat_z, cat_bands, cat_nb, cat_flux, cat_var):
    """delight-learn.py:61-90 -> gpparams rows [nT, numCol]."""
    from delight.utils import scalefree_flux_likelihood
    nt, nB = f_mod.shape
    nT = len(cat_z)
    numCol = 3 + nB + nB * (nB + 1) // 2 + nB + nt
    rows = np.zeros((nT, numCol))
    for i in range(nT):
        B = int(cat_nb[i])
        bands = np.asarray(cat_bands[i, :B], dtype=int)
        z = cat_z[i]
        fluxes, fluxesVar = cat_flux[i, :B], cat_var[i, :B]
        themod = np.zeros((1, nt, B))
        for it in range(nt):
            for ib, band in enumerate(bands):
                themod[0, it, ib] = f_mod[it, band](z)
        chi2_grid, ellMLs = scalefree_flux_likelihood(fluxes, fluxesVar, themod, returnChi2=True)
        bestType = np.argmin(chi2_grid)
        ell = ellMLs[0, bestType]
        X = np.ones((B, 3))
        X[:, 0] = bands
        X[:, 1] = z
        X[:, 2] = ell
        gp.setData(X, fluxes.reshape(-1, 1).copy(), fluxesVar.reshape(-1, 1).copy(), bestType)
        rows[i, 0] = B
        rows[i, 1] = z
        rows[i, 2] = ell
        rows[i, 3:3 + B] = bands
        rows[i, 3 + B:3 + nt + B + B * (B + 1) // 2 + B] = gp.getCore()
    return rows


def model_grid(gp, rows, zgrid, nt):
    """delight-apply.py:84-98 -> model_mean, model_covar [nz, nT, nB], redshifts [nT]."""
    nT = rows.shape[0]
    nB = gp.bands.size
    model_mean = np.zeros((zgrid.size, nT, nB))
    model_covar = np.zeros((zgrid.size, nT, nB))
    for loc in range(nT):
        B = int(rows[loc, 0])
        z, ell = rows[loc, 1], rows[loc, 2]
        X = np.zeros((B, 3))
        X[:, 0] = rows[loc, 3:3 + B]
        X[:, 1] = z
        X[:, 2] = ell
        flat = rows[loc, 3 + B:]
        gp.setCore(X, B, nt, flat[0:nt + B + B * (B + 1) // 2])
        model_mean[:, loc, :], model_covar[:, loc, :] = gp.predictAndInterpolate(zgrid, ell=ell)
    return model_mean, model_covar, rows[:, 1].copy()

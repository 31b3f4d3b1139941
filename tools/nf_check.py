"""Fused kernel vs oracle for a list of band counts (debug helper): python tools/nf_check.py 22 28 32"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as orc
from delight_b200 import apply

for nf in [int(a) for a in sys.argv[1:]]:
    nO, nT, nz, nB = 5, 20, 129, max(nf, 6)
    rng = np.random.default_rng(nf)
    zgrid = np.linspace(0.01, 3.0, nz)
    mm = rng.uniform(0.5, 2.0, (1, nT, nB)) * (1 + np.sin(zgrid[:, None, None] * rng.uniform(1, 4, (1, nT, nB))) * 0.5) * 1e5
    mc = (mm * rng.uniform(1e-3, 5e-2, (nz, nT, nB))) ** 2
    zt = rng.uniform(0.01, 3.0, nT)
    prior = orc.z_prior(zgrid, zt, 0.5)
    bands = np.tile(np.arange(nB), (nO, 1))
    nfs = np.full(nO, nf)
    fl = np.zeros((nO, nB)); fv = np.zeros((nO, nB))
    for o in range(nO):
        true = mm[rng.integers(nz), rng.integers(nT), :nf] * 4e-10
        fv[o, :nf] = (0.05 * true) ** 2
        fl[o, :nf] = true * (1 + 0.05 * rng.standard_normal(nf))
    ref = orc.apply_fused(fl, fv, nfs, bands, mm, mc, prior, zgrid, np.ones(nz), np.full(nz, 0.25), 10, want_evid=True)
    grid = apply.ModelGrid.from_legacy(mm, mc)
    res = apply.apply_fused(grid, grid.gaussian_prior(zgrid, zt, 0.5), zgrid, apply.Targets(fl, fv, bands, nfs),
                            ell_hat=1.0, ell_var=0.25, K=10, want_evid=True)
    bad = ~np.isfinite(res["pdf"])
    rel = np.abs(res["pdf"] / np.where(ref["pdf"] > 1e-250 * ref["pdf"].max(axis=1, keepdims=True), ref["pdf"], np.nan) - 1)
    ebad = ~np.isfinite(res["evid"])
    erel = np.abs(res["evid"] / np.where(ref["evid"] > 1e-250 * ref["evid"].max(axis=1, keepdims=True), ref["evid"], np.nan) - 1)
    print("nf=%2d  non-finite pdf entries %d of %d   max rel err %.2e | non-finite evid %d of %d (galaxies %s)  max rel err %.2e" %
          (nf, bad.sum(), bad.size, np.nanmax(rel), ebad.sum(), ebad.size,
           np.unique(np.argwhere(ebad)[:, 1]).tolist()[:8], np.nanmax(erel)))
    if ebad.any():
        print("   evid non-finite mask [target x galaxy]:")
        for o in range(nO):
            print("   ", "".join("X" if b else "." for b in ebad[o]), " values:", res["evid"][o][ebad[o]][:3])

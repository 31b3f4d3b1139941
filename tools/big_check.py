"""One-off scale check (config-5 shaped, one GPU's share scaled down in targets): 100k training
galaxies x 2048 targets x 1500 z x 6 bands through fit -> predict -> fused apply."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from delight_b200 import apply, gp, workloads
data = workloads.load_lsst(); H = workloads.HYPER
zgrid = data["zgrid"]; nB = 6
nT, nO = 100000, 2048
gpm = gp.PhotozGPBatched(data["tmpl"], zgrid, data["amp"], data["mu"], data["sig"], H["lines_pos"], H["V_C"], H["V_L"], H["alpha_C"], H["alpha_L"], data["zgridGP"])
tr = workloads.mock_catalogue(nT, data, seed=1, ell=1e6)
torch.cuda.synchronize(); t0 = time.time()
rows = gpm.fit(gp.catalogue(tr["fluxes"], tr["fluxes_var"], np.tile(np.arange(nB), (nT, 1)), np.full(nT, nB), tr["z"]), out_device=True)
torch.cuda.synchronize(); t1 = time.time()
grid, tz = gpm.predict(rows)
torch.cuda.synchronize(); t2 = time.time()
print("fit %.3f s  predict %.3f s  grid %.1f GB" % (t1 - t0, t2 - t1, grid.grid.numel() * 8 / 1e9))
prior = grid.gaussian_prior(zgrid, tz, 0.5)
ta = workloads.mock_catalogue(nO, data, seed=7)
t = apply.Targets.full(ta["fluxes"], ta["fluxes_var"], ta["z"])
torch.cuda.synchronize(); t0 = time.time()
res = apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50, want_metrics=True)
torch.cuda.synchronize(); t1 = time.time()
print("apply %.3f s -> %.0f gal/s, %.3g cells/s, launches %d" % (t1 - t0, nO / (t1 - t0), nO * nT * 1500 / (t1 - t0), res["fused_launches"]))
m = res["metrics"]
print("finite:", np.isfinite(res["pdf"]).all(), " |zMAP - ztrue| median:", np.median(np.abs(m[:, 3] - m[:, 0])), " top idx range:", res["top_idx"].min(), res["top_idx"].max())
assert res["top_idx"].max() < nT and res["top_idx"].min() >= 0

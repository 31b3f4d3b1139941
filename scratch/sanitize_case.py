import sys
sys.path.insert(0, '.')
import numpy as np
from delight_b200 import apply, gp, workloads
data = workloads.load_lsst(); H = workloads.HYPER
zgrid = np.ascontiguousarray(data["zgrid"][::10]); tmpl = np.ascontiguousarray(data["tmpl"][:, :, ::10])
gpm = gp.PhotozGPBatched(tmpl, zgrid, data["amp"], data["mu"], data["sig"], H["lines_pos"], .5, .5, 1e3, 1e2, data["zgridGP"][::5])
sub = dict(data, zgrid=zgrid, tmpl=tmpl)
tr = workloads.mock_catalogue(40, sub, seed=1, ell=1e6)
rows = gpm.fit(gp.catalogue(tr["fluxes"], tr["fluxes_var"], np.tile(np.arange(6), (40, 1)), np.full(40, 6), tr["z"]), out_device=True)
grid, tz = gpm.predict(rows)
prior = grid.gaussian_prior(zgrid, tz, 0.5)
ta = workloads.mock_catalogue(37, sub, seed=2)
bands = np.tile(np.arange(6), (37, 1)); nf = np.full(37, 6); nf[::3] = 4
res = apply.apply_fused(grid, prior, zgrid, apply.Targets(ta["fluxes"], ta["fluxes_var"], bands, nf, ta["z"]), ell_var=0.25, K=10, want_evid=True, want_metrics=True, confidence_levels=[0.1, 0.5], nsplit=2)
print("ok", res["pdf"].shape, np.isfinite(res["pdf"]).all())

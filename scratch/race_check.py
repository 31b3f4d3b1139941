import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from delight_b200 import apply, gp, workloads
tag = sys.argv[1]
data = workloads.load_lsst(); H = workloads.HYPER
zgrid = data["zgrid"]; nB = 6
nT, nO = 3000, 4096
gpm = gp.PhotozGPBatched(data["tmpl"], zgrid, data["amp"], data["mu"], data["sig"], H["lines_pos"], H["V_C"], H["V_L"], H["alpha_C"], H["alpha_L"], data["zgridGP"])
tr = workloads.mock_catalogue(nT, data, seed=1, ell=1e6)
rows = gpm.fit(gp.catalogue(tr["fluxes"], tr["fluxes_var"], np.tile(np.arange(nB), (nT, 1)), np.full(nT, nB), tr["z"]), out_device=True)
grid, tz = gpm.predict(rows)
prior = grid.gaussian_prior(zgrid, tz, 0.5)
ta = workloads.mock_catalogue(nO, data, seed=7)
t = apply.Targets.full(ta["fluxes"], ta["fluxes_var"], ta["z"]).to_device()
res = apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50)
torch.cuda.synchronize(); t0 = time.time()
res = apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50)
torch.cuda.synchronize(); dt = time.time() - t0
np.save("/tmp/pdf_%s.npy" % tag, res["pdf"]); np.save("/tmp/idx_%s.npy" % tag, res["top_idx"])
print(tag, "time %.3f s" % dt, "finite", np.isfinite(res["pdf"]).all(), "sum", res["pdf"].sum())
if tag != "a":
    a = np.load("/tmp/pdf_a.npy"); ia = np.load("/tmp/idx_a.npy")
    rel = np.abs(res["pdf"] - a) / np.maximum(a, 1e-280 * a.max(axis=1, keepdims=True))
    print("   vs a: max rel diff", rel.max(), " idx equal:", (ia == res["top_idx"]).mean())

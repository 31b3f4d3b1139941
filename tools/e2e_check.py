"""Actual end-to-end errors of the device pipeline (fit -> predict -> fused apply) against the
reference's outputs on the golden SDSS mock (needs a GPU): the numbers behind the tolerances of
tests/test_gpu_parity.py::test_full_pipeline_learn_predict_apply."""
import sys
import numpy as np
sys.path.insert(0, ".")
from delight_b200 import apply, gp

g = np.load("tests/golden/sdss_pipeline.npz")
gpm = gp.PhotozGPBatched(g["tmpl"], g["zgrid"], g["amp"], g["mu"], g["sig"], g["lines_pos"],
                         float(g["V_C"]), float(g["V_L"]), float(g["alpha_C"]),
                         float(g["alpha_L"]), g["zgridGP"])
cat = gp.catalogue(g["tr_flux"], g["tr_var"], g["tr_bands"], g["tr_nb"], g["tr_z"])
rows = gpm.fit(cat)
ref = g["gp_rows"]
m = ref != 0
print("fit rows: max rel err %.2e" % np.abs(rows[m] / ref[m] - 1).max())
grid, tz = gpm.predict(rows)
nz, nT, nB = g["model_mean"].shape
G = grid.grid.cpu().numpy()
print("model grid from device rows: mean %.2e var %.2e" % (
    np.abs(G[:, :nB, :nz].transpose(2, 0, 1) / g["model_mean"] - 1).max(),
    np.abs(G[:, nB:, :nz].transpose(2, 0, 1) / g["model_covar"] - 1).max()))
prior = grid.gaussian_prior(g["zgrid"], tz, float(g["zPriorSigma"]))
t = apply.Targets(g["ta_flux"], g["ta_var"], g["ta_bands"], g["ta_nb"], g["ta_z"])
K = g["top_idx"].shape[1]
res = apply.apply_fused(grid, prior, g["zgrid"], t, ell_hat=1.0,
                        ell_var=float(g["ellPriorSigma"]) ** 2, K=K, want_evid=True,
                        want_metrics=True, confidence_levels=g["confidenceLevels"])
big = g["pdfs"] > 1e-200 * g["pdfs"].max(axis=1, keepdims=True)
print("end to end: pdf %.2e  evid %.2e  metrics %.2e  top-K equal %s" % (
    np.abs(res["pdf"][big] / g["pdfs"][big] - 1).max(),
    np.abs(res["top_ev"] / g["top_ev"] - 1).max(),
    np.abs(res["metrics"][:, :g["metrics"].shape[1]] / g["metrics"] - 1).max(),
    np.array_equal(res["top_idx"], g["top_idx"])))

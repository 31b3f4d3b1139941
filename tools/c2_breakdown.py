import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from delight_b200 import apply, workloads
data = workloads.load_lsst()
zgrid, tmpl = data["zgrid"], data["tmpl"][:, :5]
nt, nB, nz = tmpl.shape
f_mod = np.ascontiguousarray(tmpl.transpose(2, 0, 1))
grid = apply.ModelGrid.from_legacy(f_mod, np.zeros_like(f_mod))
prior = grid.template_prior(zgrid, data["p_t"], data["p_z_t"])
rng = np.random.default_rng(3)
nO = 100000
ztrue = zgrid[rng.integers(0, nz, nO)]
fl = f_mod[np.searchsorted(zgrid, ztrue), rng.integers(0, nt, nO)] * rng.uniform(.5, 2, (nO, 1))
fl = fl * (1 + 0.05 * rng.standard_normal(fl.shape)); fv = (0.05 * fl) ** 2
ws = apply.FusedWorkspace()
tg = apply.Targets.full(fl, fv, ztrue).to_device()
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
CL = [0.1, 0.5, 0.68, 0.95]
print("pdf only            %.1f ms" % t(lambda: apply.apply_fused(grid, prior, zgrid, tg, ell_var=1e24, K=0, workspace=ws, out_device=True)))
print("pdf + metrics       %.1f ms" % t(lambda: apply.apply_fused(grid, prior, zgrid, tg, ell_var=1e24, K=0, want_metrics=True, workspace=ws, out_device=True)))
print("pdf + metrics + CI  %.1f ms" % t(lambda: apply.apply_fused(grid, prior, zgrid, tg, ell_var=1e24, K=0, want_metrics=True, confidence_levels=CL, workspace=ws, out_device=True)))

"""G10-style TARGET catalogue (28 bands, every target a random subset of 2..21 of them): time of
the fused apply with per-target band lists (one launch per band COUNT) against the same targets
with the first nf bands each (homogeneous band set per count -> the fast kernels).
python tools/het_scan.py [nO nT]   (needs a GPU)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from delight_b200 import apply

nO, nT = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 1000)
nz, nB = 1500, 28
zgrid = np.linspace(0.01, 3.0, nz)
g = torch.Generator(device="cuda").manual_seed(1)
mean = 1e5 * (0.5 + torch.rand((nz, nT, nB), generator=g, device="cuda", dtype=torch.float64))
var = (mean * 0.01) ** 2
grid = apply.ModelGrid.from_legacy(mean, var)
del mean, var
rng = np.random.default_rng(2)
prior = grid.gaussian_prior(zgrid, rng.uniform(0.01, 3.0, nT), 0.5)
nfs = rng.integers(2, 22, nO)
fl = np.zeros((nO, nB)); fv = np.zeros((nO, nB))
het = np.zeros((nO, nB), dtype=np.int64); hom = np.zeros((nO, nB), dtype=np.int64)
for o in range(nO):
    het[o, :nfs[o]] = np.sort(rng.choice(nB, nfs[o], replace=False))
    hom[o, :nfs[o]] = np.arange(nfs[o])
    fl[o, :nfs[o]] = 4e-5 * rng.uniform(0.5, 1.5, nfs[o])
    fv[o, :nfs[o]] = (0.05 * fl[o, :nfs[o]]) ** 2
ws = apply.FusedWorkspace()
for name, bands in (("per-target band lists", het), ("homogeneous per count", hom)):
    t = apply.Targets(fl, fv, bands, nfs).to_device()
    r = apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50, workspace=ws, out_device=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50, workspace=ws, out_device=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cells = float((nfs * 0 + 1).sum()) * nT * nz
    flop = float((27 * nfs + 30).sum()) * nT * nz
    print("%-24s %d band sets, %2d launches: %8.1f ms  %.3e cells/s  %.1f TFLOP/s algorithmic" % (
        name, len(list(t.groups())), r["fused_launches"], dt * 1e3, cells / dt, flop / dt * 1e-12))

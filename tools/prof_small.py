"""One invocation of every non-fused kernel of the library at bench-like sizes, for the ncu pass of
tools/prof_r2.sh (subsystems 2 and 3, the drop-in operator, the epilogue kernels).  Prints CUDA
event times per section (needs a GPU)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from delight_b200 import _lib, apply, gp as gpmod, workloads  # noqa: E402

L = _lib.lib()
data = workloads.load_lsst()
H = workloads.HYPER
zgrid = data["zgrid"]
nz, nB = zgrid.size, data["tmpl"].shape[1]


def timed(name, fn, reps=1):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %9.3f ms" % (name, e0.elapsed_time(e1) / reps))
    return out


# (2) kernel grids: 6 LSST bands and 28 synthetic bands, nzGP = 150
mk6 = lambda: gpmod.PhotozGPBatched(data["tmpl"], zgrid, data["amp"], data["mu"], data["sig"],
                                    H["lines_pos"], H["V_C"], H["V_L"], H["alpha_C"],
                                    H["alpha_L"], data["zgridGP"])
gpm = timed("kernel grids, 6 bands x 150^2 (incl. uploads)", mk6)
# (3) learn / predict for 10k galaxies
nT = 10000
tr = workloads.mock_catalogue(nT, data, seed=1, ell=1e6)
tcat = gpmod.catalogue(tr["fluxes"], tr["fluxes_var"], np.tile(np.arange(nB), (nT, 1)),
                       np.full(nT, nB), tr["z"])
rows = timed("gp_fit 10k galaxies x 6 bands (incl. H2D)", lambda: gpm.fit(tcat, out_device=True))
grid, train_z = timed("gp_predict 10k galaxies -> 1500 z", lambda: gpm.predict(rows))
prior = grid.gaussian_prior(zgrid, train_z, 0.5)
# drop-in operator on device-resident inputs: one target against the whole grid
G = grid.grid[:, :, :nz]
mm = G[:, :nB, :].permute(2, 0, 1).contiguous()
mc = G[:, nB:, :].permute(2, 0, 1).contiguous()
cat = workloads.mock_catalogue(64, data, seed=2)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
like = torch.empty((nz, nT), dtype=torch.float64, device="cuda")
scr = torch.zeros(1, dtype=torch.int64, device="cuda")
fo, fv, eh, ev = d(cat["fluxes"][0]), d(cat["fluxes_var"][0]), d(np.ones(nz)), d(np.full(nz, .25))
timed("non-fused likelihood 1500 x 10k x 6 (device)", lambda: _lib.check(
    L.dl_approx_flux_likelihood_dev(_lib.ptr(like), nz, nT, nB, _lib.ptr(fo), _lib.ptr(fv),
                                    _lib.ptr(mm), _lib.ptr(mc), _lib.ptr(eh), _lib.ptr(ev),
                                    _lib.ptr(scr), _lib.stream_ptr())), reps=5)
cells = nz * nT
# epilogue kernels at the bench shape: 2048 targets (top-K over 10k evidences, metrics, CI levels)
t = apply.Targets.full(cat["fluxes"].repeat(32, axis=0), cat["fluxes_var"].repeat(32, axis=0),
                       cat["z"].repeat(32))
res = timed("fused apply 2048 targets + top-50 + metrics + CI", lambda: apply.apply_fused(
    grid, prior, zgrid, t, ell_var=0.25, K=50, want_metrics=True,
    confidence_levels=[0.1, 0.5, 0.68, 0.95], out_device=True))
timed("compressed second pass, 2048 targets x 50", lambda: apply.apply_compressed(
    grid, prior, zgrid, t, res["top_idx"].cpu().numpy(), ell_var=0.25, out_device=True))
print("non-fused likelihood: %d cells, algorithmic bytes %d per call (2 nf 8 read + 8 written + "
      "16 second pass)" % (cells, cells * (2 * nB * 8 + 24)))

"""Synthetic catalogues of the BASELINE.json shapes (no network, no survey data).

Mocks follow scripts/simulateWithSEDs.py:27-62: z ~ U(zgrid[0], zgrid[-1]), type ~ randint(nt),
flux_b = ell * T[type, b](z) (linear interpolation of the <sed>_fluxredshiftmod tables),
Gaussian noise sigma = noiseLevel * flux, variance column = sigma^2; ell = 1e6 for the training
set and 1 for targets.  The tables (8 CWW templates x 6 LSST bands x 1500 redshifts) and the
filters' Gaussian-mixture coefficients in data/lsst_cww.npz are outputs of the reference's own
processFilters.py / processSEDs.py (tests/golden/make_golden.py).
"""
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lsst_cww.npz")


def load_lsst():
    d = np.load(DATA)
    return {k: d[k] for k in d.files}


def interp_table(tmpl, zgrid, types, z):
    """T[type, :, z] by linear interpolation on the uniform fine grid -> [n, nB]."""
    x = (z - zgrid[0]) / (zgrid[1] - zgrid[0])
    i = np.clip(np.floor(x).astype(np.int64), 0, zgrid.size - 2)
    w = (x - i)[:, None]
    return tmpl[types, :, i] * (1 - w) + tmpl[types, :, i + 1] * w


# hyper-parameters of tests/parametersTest.cfg:55-64 (the reference's shipped defaults)
HYPER = dict(V_C=0.5, V_L=0.5, alpha_C=1e3, alpha_L=1e2, zPriorSigma=0.5, ellPriorSigma=0.5,
             lines_pos=[6500.0, 5002.26, 3732.22], lines_width=[20.0, 20.0, 20.0],
             extraFracFluxError=1e-5)


def mock_catalogue(n, data, seed, ell=1.0, noise=0.05, nB=None, extra_frac=1e-5):
    """simulateWithSEDs-style draw as the catalogue reader hands it on (io.py:349-351 adds
    (extraFracFluxError * flux)^2 to every variance).
    Returns dict(z, types, fluxes [n,nB], fluxes_var [n,nB])."""
    rng = np.random.default_rng(seed)
    zgrid, tmpl = data["zgrid"], data["tmpl"]
    nB = nB or tmpl.shape[1]
    z = rng.uniform(zgrid[0], zgrid[-1], n)
    types = rng.integers(0, tmpl.shape[0], n)
    true = ell * interp_table(tmpl[:, :nB], zgrid, types, z)
    sig = true * noise
    fluxes = true + sig * rng.standard_normal(true.shape)
    return dict(z=z, types=types, fluxes=fluxes, fluxes_var=sig ** 2 + (extra_frac * fluxes) ** 2)


def standin_model_grid(nT, data, seed, nB=None, device="cuda"):
    """Stand-in for learn + predict when only the likelihood is being measured: per training
    galaxy its best-fit template scaled to its fluxes, with a smooth z-dependent perturbation
    (what the GP residual adds) and a variance that grows away from the galaxy's own redshift.
    Built with torch on `device` (input generation, not the hot path).
    Returns (mean [nz, nT, nB], var [nz, nT, nB], train_z [nT]) in the legacy layout."""
    import torch
    cat = mock_catalogue(nT, data, seed, ell=1e6, nB=nB)
    zgrid = torch.from_numpy(data["zgrid"]).to(device)
    tmpl = torch.from_numpy(np.ascontiguousarray(data["tmpl"][:, :nB or data["tmpl"].shape[1]]))
    tmpl = tmpl.to(device)
    g = torch.Generator(device=device).manual_seed(seed + 17)
    types = torch.from_numpy(cat["types"]).to(device)
    zt = torch.from_numpy(cat["z"]).to(device)
    nB_ = tmpl.shape[1]
    base = tmpl[types].permute(2, 0, 1) * 1e6                       # [nz, nT, nB]
    amp = 0.05 * torch.rand((1, nT, nB_), generator=g, device=device, dtype=torch.float64)
    freq = 2 + 6 * torch.rand((1, nT, nB_), generator=g, device=device, dtype=torch.float64)
    dz = zgrid[:, None, None] - zt[None, :, None]
    mean = base * (1 + amp * torch.sin(freq * dz))
    var = (mean * (0.01 + 0.05 * (1 - torch.exp(-0.5 * (dz / 0.3) ** 2)))) ** 2
    return mean.contiguous(), var.contiguous(), cat["z"]

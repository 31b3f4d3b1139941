"""Fused-kernel time vs number of bands (synthetic grid, kernel events via dl_timing_*):
python tools/nf_scan.py [nO nT]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from delight_b200 import _lib, apply

nO, nT = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1500)
nz = 1500
L = _lib.lib()
zgrid = np.linspace(0.01, 3.0, nz)
for nf in [int(x) for x in os.environ.get('NF_LIST', '4,6,8,9,10,13,16,21,28').split(',')]:
    g = torch.Generator(device="cuda").manual_seed(nf)
    mean = 1e5 * (0.5 + torch.rand((nz, nT, nf), generator=g, device="cuda", dtype=torch.float64))
    var = (mean * 0.01) ** 2
    grid = apply.ModelGrid.from_legacy(mean, var)
    tz = np.random.default_rng(nf).uniform(0.01, 3.0, nT)
    prior = grid.gaussian_prior(zgrid, tz, 0.5)
    rng = np.random.default_rng(1)
    fl = 4e-5 * rng.uniform(0.5, 1.5, (nO, nf))
    t = apply.Targets(fl, (0.05 * fl) ** 2, np.tile(np.arange(nf), (nO, 1)), np.full(nO, nf), None)
    apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50, out_device=True)
    L.dl_timing_enable(1)
    apply.apply_fused(grid, prior, zgrid, t, ell_var=0.25, K=50, out_device=True)
    torch.cuda.synchronize()
    ms, n = ctypes.c_double(), ctypes.c_int()
    L.dl_timing_collect(ctypes.byref(ms), ctypes.byref(n))
    L.dl_timing_enable(0)
    cells = nO * nT * nz
    print("nf=%2d  fused kernel %8.2f ms  %.3e cells/s  %.1f TFLOP/s algorithmic" %
          (nf, ms.value, cells / ms.value * 1e3, cells * (27 * nf + 30) / ms.value * 1e-9))
    del grid, prior, mean, var

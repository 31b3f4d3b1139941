"""Where does the model-variance difference GPU vs reference come from?  (needs a GPU)
Compares, on the golden SDSS mock: device kernel grids vs the reference's; device model_var vs
the reference's and vs the long-double truth evaluated on (a) the reference's kernel grids and
(b) the device's own kernel grids."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as orc
from delight_b200 import gp

g = np.load("tests/golden/sdss_pipeline.npz")
gpm = gp.PhotozGPBatched(g["tmpl"], g["zgrid"], g["amp"], g["mu"], g["sig"], g["lines_pos"],
                         float(g["V_C"]), float(g["V_L"]), float(g["alpha_C"]),
                         float(g["alpha_L"]), g["zgridGP"])
rel = lambda a, b: np.abs(a / b - 1).max()
dKC, dKL = gpm.KC.cpu().numpy(), gpm.KL.cpu().numpy()
dKCd, dKLd = gpm.KCd.cpu().numpy(), gpm.KLd.cpu().numpy()
print("kernel grids device vs reference: KC %.2e KL %.2e KCd %.2e KLd %.2e" % (
    rel(dKC, g["KC_grid"]), rel(dKL, g["KL_grid"]), rel(dKCd, g["KCd"]), rel(dKLd, g["KLd"])))
grid, tz = gpm.predict(g["gp_rows"])
G = grid.grid.cpu().numpy()
nz = g["zgrid"].size
nB = g["amp"].shape[0]
dmean = G[:, :nB, :nz].transpose(2, 0, 1)
dvar = G[:, nB:, :nz].transpose(2, 0, 1)
print("device vs reference: mean %.2e var %.2e" % (rel(dmean, g["model_mean"]), rel(dvar, g["model_covar"])))
rows = g["gp_rows"]
nt = g["tmpl"].shape[0]
for name, grids in (("reference grids", dict(KC=g["KC_grid"], KL=g["KL_grid"], KCd=g["KCd"], KLd=g["KLd"], fzGrid=1 + g["zgridGP"])),
                    ("device grids", dict(KC=dKC, KL=dKL, KCd=dKCd, KLd=dKLd, fzGrid=1 + g["zgridGP"]))):
    worst = [0, 0, 0]
    for loc in range(rows.shape[0]):
        B = int(rows[loc, 0]); z, ell = rows[loc, 1], rows[loc, 2]
        bands = rows[loc, 3:3 + B].astype(int)
        flat = rows[loc, 3 + B:]
        best = int(np.argmax(flat[:nt]))
        Lm = np.zeros((B, B)); Lm[np.tril_indices(B)] = flat[nt:nt + B * (B + 1) // 2]
        D = flat[nt + B * (B + 1) // 2:nt + B * (B + 1) // 2 + B]
        hm, hv = orc.gp_predict_interp_hp(grids, float(g["V_C"]), float(g["V_L"]), g["tmpl"],
                                          g["zgrid"], g["zgridGP"], bands, z, ell, best, Lm, D,
                                          g["zgrid"])
        hv = hv.astype(float)
        worst[0] = max(worst[0], rel(dvar[:, loc, :], hv))
        worst[1] = max(worst[1], rel(g["model_covar"][:, loc, :], hv))
        worst[2] = max(worst[2], rel(dmean[:, loc, :], hm.astype(float)))
    print("truth on %s: device var %.2e  reference var %.2e  device mean %.2e" % (name, *worst))

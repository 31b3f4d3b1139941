import numpy as np, sys
sys.path.insert(0,'.')
from delight_b200 import apply
from oracle import oracle as orc
rng = np.random.default_rng(5)
nO, nT, nz, nB = 96, 512, 1500, 6
zgrid = 0.01 + 0.002 * np.arange(nz)
base = rng.uniform(0.5, 2.0, (1, nT, nB))
mm = 1e5 * base * (1 + 0.6 * np.sin(zgrid[:, None, None] * rng.uniform(2, 6, (1, nT, nB))))
mc = (mm * 0.01) ** 2
zt = rng.uniform(0.01, 3.0, nT)
src, zi = rng.integers(0, nT, nO), rng.integers(0, nz, nO)
fl = mm[zi, src] * 4e-10
fv = (0.02 * fl) ** 2
grid = apply.ModelGrid.from_legacy(mm, mc)
prior = grid.gaussian_prior(zgrid, zt, 0.5)
ref = orc.apply_fused(fl, fv, np.full(nO,nB), np.tile(np.arange(nB),(nO,1)), mm, mc, orc.z_prior(zgrid, zt, 0.5), zgrid, np.ones(nz), np.full(nz,.25), 50, want_evid=True)
def cmp(name, res, perm=None):
    p = res["pdf"]; r = ref["pdf"] if perm is None else ref["pdf"][perm]
    scale = r.max(axis=1, keepdims=True)
    rel = np.abs(p-r)/np.maximum(r, 1e-280*scale)
    big = r > 1e-100*scale
    print(name, "max rel (r>1e-100 max):", rel[big].max(), " max rel all:", rel.max(), "argmax", np.unravel_index(np.argmax(np.where(big, rel, 0)), rel.shape), "value there", r[np.unravel_index(np.argmax(np.where(big, rel, 0)), rel.shape)]/scale[np.unravel_index(np.argmax(np.where(big, rel, 0)), rel.shape)[0]])
    if perm is None:
        e = res["evid"]; er = ref["evid"]; es = er.max(axis=1, keepdims=True)
        rel = np.abs(e-er)/np.maximum(er, 1e-280*es); big = er > 1e-100*es
        print("   evid max rel:", rel[big].max())
for ns in (1, 2, 4, 9, None):
    cmp("nsplit=%s"%ns, apply.apply_fused(grid, prior, zgrid, apply.Targets.full(fl, fv), ell_var=0.25, K=50, nsplit=ns, want_evid=True))
perm = rng.permutation(nO)
for ns in (1, 9):
    cmp("perm nsplit=%s"%ns, apply.apply_fused(grid, prior, zgrid, apply.Targets.full(fl[perm], fv[perm]), ell_var=0.25, K=50, nsplit=ns), perm)

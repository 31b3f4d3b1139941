"""Batched per-training-galaxy Gaussian process on the device.

Host-side mirror of delight.photoz_gp.PhotozGP for the two uses the pipeline makes of it:

    gp = PhotozGPBatched(tmpl, zgrid, amp, mu, sig, lines_pos, V_C, V_L, alpha_C, alpha_L,
                         redshiftGridGP)                       # = PhotozGP(...) constructor
    rows = gp.fit(catalogue)          # delight-learn.py loop  -> gpparams rows [nT, numCol]
    grid = gp.predict(rows)           # delight-apply.py:89-98 -> apply.ModelGrid (device)

`tmpl` is the [nt, nB, nz] stack of the <sed>_fluxredshiftmod.txt tables the reference wraps in
interp1d objects (io.py:263-277); everything else has the reference constructor's meaning
(photoz_gp.py:158-184).
"""
import numpy as np

from . import _lib
from .apply import ModelGrid, Targets


class PhotozGPBatched:
    def __init__(self, tmpl, zgrid, bandCoefAmplitudes, bandCoefPositions, bandCoefWidths,
                 lines_pos, var_C, var_L, alpha_C, alpha_L, redshiftGridGP):
        torch = _lib.require_cuda()
        L = _lib.lib()
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        amp = np.asarray(bandCoefAmplitudes, dtype=np.float64)
        self.nB, self.NC = amp.shape
        self.nt = int(np.shape(tmpl)[0])
        self.nz = int(np.size(zgrid))
        if np.shape(tmpl) != (self.nt, self.nB, self.nz):
            raise ValueError("tmpl must be [nt, nB, nz]")
        self.V_C, self.V_L = float(var_C), float(var_L)
        self.alpha_C, self.alpha_L = float(alpha_C), float(alpha_L)
        zgp = np.ascontiguousarray(redshiftGridGP, dtype=np.float64)
        self.nzg = zgp.size
        lines = np.ascontiguousarray(lines_pos, dtype=np.float64)
        self.NL = lines.size
        # Photoz_kernel.norms (photoz_kernels.py:186-187)
        norms = np.sqrt(2 * np.pi) * np.sum(amp * np.asarray(bandCoefWidths), axis=1)
        self.d_amp, self.d_mu, self.d_sig = dev(amp), dev(bandCoefPositions), dev(bandCoefWidths)
        self.d_lines = dev(lines if self.NL else np.zeros(1))
        self.d_norms = dev(norms)
        self.d_zgp, self.d_fz = dev(zgp), dev(1.0 + zgp)
        self.d_tmpl, self.d_zgrid = dev(tmpl), dev(zgrid)
        nB, nzg = self.nB, self.nzg
        # construct_interpolators (photoz_kernels.py:342-416) in one launch
        self.KC = torch.empty((nB, nB, nzg, nzg), dtype=torch.float64, device="cuda")
        self.KL = torch.empty_like(self.KC)
        self.KCd = torch.empty((nB, nzg), dtype=torch.float64, device="cuda")
        self.KLd = torch.empty_like(self.KCd)
        _lib.check(L.dl_kernel_grids_dev(
            nB, nzg, self.NC, self.NL, self.alpha_C, self.alpha_L, _lib.ptr(self.d_amp),
            _lib.ptr(self.d_mu), _lib.ptr(self.d_sig), _lib.ptr(self.d_lines),
            _lib.ptr(self.d_norms), _lib.ptr(self.d_fz), _lib.ptr(self.KC), _lib.ptr(self.KL),
            _lib.ptr(self.KCd), _lib.ptr(self.KLd), _lib.stream_ptr()))
        self.numCol = 3 + nB + nB * (nB + 1) // 2 + nB + self.nt     # delight-learn.py:46
        # galaxy-independent lookup tables of the predict kernels, once per model
        self.tables = torch.empty(int(L.dl_gp_tables_bytes(nzg, self.nz)), dtype=torch.uint8,
                                  device="cuda")
        _lib.check(L.dl_gp_tables_dev(nzg, self.nz, _lib.ptr(self.d_zgp), _lib.ptr(self.d_fz),
                                      _lib.ptr(self.d_zgrid), _lib.ptr(self.tables),
                                      _lib.stream_ptr()))

    def fit(self, cat, out_device=False):
        """cat: apply.Targets-like packed catalogue (fluxes, fluxes_var, bands, nf, ztrue=z).
        Returns the gpparams rows [nT, numCol]."""
        torch = _lib.require_cuda()
        if cat.ztrue is None:
            raise ValueError("training catalogue needs redshifts")
        nT = len(cat)
        if cat.fluxes.shape[1] != self.nB:
            raise ValueError("catalogue arrays must be [nT, nB] left-packed")
        if nT and (cat.bands[np.arange(cat.bands.shape[1])[None, :] < cat.nf[:, None]].max()
                   >= self.nB):
            raise ValueError("band index outside the model")
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        rows = torch.zeros((nT, self.numCol), dtype=torch.float64, device="cuda")
        if nT:
            args = [dev(cat.nf), dev(np.where(cat.bands < 0, 0, cat.bands)), dev(cat.ztrue),
                    dev(cat.fluxes), dev(cat.fluxes_var)]
            _lib.check(_lib.lib().dl_gp_fit_dev(
                nT, self.nB, self.nt, self.nzg, self.nz, _lib.ptr(self.KC), _lib.ptr(self.KL),
                _lib.ptr(self.d_zgp), _lib.ptr(self.d_fz), _lib.ptr(self.d_tmpl),
                _lib.ptr(self.d_zgrid), self.V_C, self.V_L, *[_lib.ptr(a) for a in args],
                _lib.ptr(rows), self.numCol, _lib.stream_ptr()))
        bad = ~torch.isfinite(rows).all(dim=1)
        if bool(bad.any().item()):
            # scipy.linalg.cholesky raises LinAlgError in the reference (photoz_gp.py:209)
            idx = torch.nonzero(bad).flatten()[:8].cpu().tolist()
            raise _lib.DelightB200Error(
                "GP fit failed (K + diag(Yvar) not positive definite or non-finite input) for "
                "%d training galaxies, first rows: %s" % (int(bad.sum().item()), idx))
        return rows if out_device else rows.cpu().numpy()

    def predict(self, rows):
        """gpparams rows (NumPy or CUDA tensor) -> (ModelGrid in the fused layout, train_z)."""
        torch = _lib.require_cuda()
        if not isinstance(rows, torch.Tensor):
            rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64))
        rows = rows.cuda().contiguous()
        nT, numCol = rows.shape
        if nT:
            # range-check what the kernel will use as indices (a corrupt gpparams file must not
            # become an out-of-bounds read on the device)
            B = rows[:, 0]
            col = torch.arange(self.nB, device="cuda")[None, :]
            bands = rows[:, 3:3 + self.nB]
            used = col < B[:, None]
            if bool(((B < 0) | (B > self.nB) | (B != B.round())).any().item()) or \
                    bool((used & ((bands < 0) | (bands >= self.nB) | (bands != bands.round())))
                         .any().item()):
                raise ValueError("gpparams rows: band count / band indices out of range")
        nzp = int(_lib.lib().dl_fused_padded_nz(self.nz))
        grid = torch.empty((nT, 2 * self.nB, nzp), dtype=torch.float64, device="cuda")
        train_z = torch.empty(nT, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().dl_gp_predict_dev(
            nT, self.nB, self.nt, self.nzg, self.nz, _lib.ptr(self.KC), _lib.ptr(self.KL),
            _lib.ptr(self.KCd), _lib.ptr(self.KLd), _lib.ptr(self.d_zgp), _lib.ptr(self.d_fz),
            _lib.ptr(self.d_tmpl), _lib.ptr(self.d_zgrid), self.V_C, self.V_L, _lib.ptr(rows),
            numCol, _lib.ptr(grid), _lib.ptr(train_z), _lib.ptr(self.tables), _lib.stream_ptr()))
        return ModelGrid(grid, nT, self.nB, self.nz), train_z


    def predict_at(self, rows, z_eval):
        """The GP of every row evaluated at ONE redshift each (photoz_gp.py:293-347 with
        redshiftGrid = [z], as the cross-validation of delight-learn.py:92-94 calls it):
        returns (mean [nT, nB], var [nT, nB]) as NumPy arrays."""
        torch = _lib.require_cuda()
        if not isinstance(rows, torch.Tensor):
            rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64))
        rows = rows.cuda().contiguous()
        nT, numCol = rows.shape
        ze = torch.from_numpy(np.ascontiguousarray(z_eval, dtype=np.float64)).cuda()
        if ze.numel() != nT:
            raise ValueError("z_eval must have one redshift per row")
        out = torch.empty((nT, 2 * self.nB), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().dl_gp_predict_at_dev(
            nT, self.nB, self.nt, self.nzg, self.nz, _lib.ptr(self.KC), _lib.ptr(self.KL),
            _lib.ptr(self.KCd), _lib.ptr(self.KLd), _lib.ptr(self.d_zgp), _lib.ptr(self.d_fz),
            _lib.ptr(self.d_tmpl), _lib.ptr(self.d_zgrid), self.V_C, self.V_L, _lib.ptr(rows),
            numCol, _lib.ptr(ze), _lib.ptr(out), _lib.ptr(self.tables), _lib.stream_ptr()))
        out = out.cpu().numpy()
        return out[:, :self.nB], out[:, self.nB:]


def catalogue(fluxes, fluxes_var, bands, nf, z):
    """Packed training catalogue (same container as the targets)."""
    return Targets(fluxes, fluxes_var, bands, nf, z)

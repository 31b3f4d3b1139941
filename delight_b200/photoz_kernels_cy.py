"""Drop-in for the reference's Cython module `delight.photoz_kernels_cy`.

Same names, signatures (output order KL, KC, ...), accumulate-then-normalise contract of
`kernelparts`, overwrite contract of `kernelparts_diag` / `kernel_parts_interp`, and strided
in-place outputs (e.g. KL_grid[b1, b2, :, :] as passed by delight/photoz_kernels.py:366).
"""
import numpy as np

from . import _lib
from .utils_cy import _view, _dev, _dev_out

__all__ = ["kernelparts", "kernelparts_diag", "kernel_parts_interp"]


def _coefs(fcoefs_amp, fcoefs_mu, fcoefs_sig, lines_mu, norms, NC, NL):
    amp = _view(fcoefs_amp, np.float64, 2, "fcoefs_amp")
    mu = _view(fcoefs_mu, np.float64, 2, "fcoefs_mu")
    sig = _view(fcoefs_sig, np.float64, 2, "fcoefs_sig")
    lm = _view(lines_mu, np.float64, 1, "lines_mu")
    nr = _view(norms, np.float64, 1, "norms")
    if amp.shape != mu.shape or amp.shape != sig.shape or amp.shape[1] < NC or lm.size < NL \
            or nr.size < amp.shape[0]:
        raise ValueError("filter coefficient arrays have inconsistent shapes")
    return (_dev(amp[:, :NC]), _dev(mu[:, :NC]), _dev(sig[:, :NC]),
            _dev(lm[:NL]) if NL > 0 else _dev(np.zeros(1)), _dev(nr), amp.shape[0])


def _check_bands(b, nB, name):
    if b.size and (b.min() < 0 or b.max() >= nB):
        raise ValueError("%s contains band indices outside [0, %d)" % (name, nB))


def kernelparts(NO1, NO2, NC, NL, alpha_C, alpha_L, fcoefs_amp, fcoefs_mu, fcoefs_sig,
                lines_mu, lines_sig, norms, b1, fz1, b2, fz2, grad_needed,
                KL, KC, D_alpha_C, D_alpha_L, D_alpha_z):
    """delight/photoz_kernels_cy.pyx:103-177."""
    NO1, NO2, NC, NL = int(NO1), int(NO2), int(NC), int(NL)
    b1 = _view(b1, np.int64, 1, "b1")
    b2 = _view(b2, np.int64, 1, "b2")
    fz1 = _view(fz1, np.float64, 1, "fz1")
    fz2 = _view(fz2, np.float64, 1, "fz2")
    _view(lines_sig, np.float64, 1, "lines_sig")       # accepted, never read (as the reference)
    outs = [_view(a, np.float64, 2, n) for a, n in
            ((KL, "KL"), (KC, "KC"), (D_alpha_C, "D_alpha_C"), (D_alpha_L, "D_alpha_L"),
             (D_alpha_z, "D_alpha_z"))]
    if NO1 * NO2 == 0:
        return
    if b1.size < NO1 or fz1.size < NO1 or b2.size < NO2 or fz2.size < NO2 or \
            any(o.shape[0] < NO1 or o.shape[1] < NO2 for o in outs):
        raise ValueError("array shapes are inconsistent with NO1=%d NO2=%d" % (NO1, NO2))
    d_amp, d_mu, d_sig, d_lm, d_nr, nB = _coefs(fcoefs_amp, fcoefs_mu, fcoefs_sig, lines_mu,
                                                norms, NC, NL)
    _check_bands(b1[:NO1], nB, "b1")
    _check_bands(b2[:NO2], nB, "b2")
    d_b1, d_f1, d_b2, d_f2 = _dev(b1[:NO1]), _dev(fz1[:NO1]), _dev(b2[:NO2]), _dev(fz2[:NO2])
    d_out = [_dev_out(o[:NO1, :NO2]) for o in outs]
    _lib.check(_lib.lib().dl_kernelparts_dev(
        NO1, NO2, NC, NL, float(alpha_C), float(alpha_L), _lib.ptr(d_amp), _lib.ptr(d_mu),
        _lib.ptr(d_sig), _lib.ptr(d_lm), _lib.ptr(d_nr), _lib.ptr(d_b1), _lib.ptr(d_f1),
        _lib.ptr(d_b2), _lib.ptr(d_f2), int(bool(grad_needed)), *[_lib.ptr(o) for o in d_out],
        NO2, _lib.stream_ptr()))
    for o, d in zip(outs[:5 if grad_needed else 2], d_out):
        o[:NO1, :NO2] = d.cpu().numpy()


def kernelparts_diag(NO1, NC, NL, alpha_C, alpha_L, fcoefs_amp, fcoefs_mu, fcoefs_sig,
                     lines_mu, lines_sig, norms, b1, fz1, grad_needed,
                     KL, KC, D_alpha_C, D_alpha_L):
    """delight/photoz_kernels_cy.pyx:38-100."""
    NO1, NC, NL = int(NO1), int(NC), int(NL)
    b1 = _view(b1, np.int64, 1, "b1")
    fz1 = _view(fz1, np.float64, 1, "fz1")
    _view(lines_sig, np.float64, 1, "lines_sig")
    outs = [_view(a, np.float64, 1, n) for a, n in
            ((KL, "KL"), (KC, "KC"), (D_alpha_C, "D_alpha_C"), (D_alpha_L, "D_alpha_L"))]
    if NO1 == 0:
        return
    if b1.size < NO1 or fz1.size < NO1 or any(o.size < NO1 for o in outs):
        raise ValueError("array shapes are inconsistent with NO1=%d" % NO1)
    d_amp, d_mu, d_sig, d_lm, d_nr, nB = _coefs(fcoefs_amp, fcoefs_mu, fcoefs_sig, lines_mu,
                                                norms, NC, NL)
    _check_bands(b1[:NO1], nB, "b1")
    d_b1, d_f1 = _dev(b1[:NO1]), _dev(fz1[:NO1])
    d_out = [_dev_out(o[:NO1]) for o in outs]
    _lib.check(_lib.lib().dl_kernelparts_diag_dev(
        NO1, NC, NL, float(alpha_C), float(alpha_L), _lib.ptr(d_amp), _lib.ptr(d_mu),
        _lib.ptr(d_sig), _lib.ptr(d_lm), _lib.ptr(d_nr), _lib.ptr(d_b1), _lib.ptr(d_f1),
        int(bool(grad_needed)), *[_lib.ptr(o) for o in d_out], _lib.stream_ptr()))
    for o, d in zip(outs[:4 if grad_needed else 2], d_out):
        o[:NO1] = d.cpu().numpy()


def kernel_parts_interp(NO1, NO2, Kinterp, b1, fz1, p1s, b2, fz2, p2s, fzGrid, Kgrid):
    """delight/photoz_kernels_cy.pyx:9-35."""
    NO1, NO2 = int(NO1), int(NO2)
    Kinterp = _view(Kinterp, np.float64, 2, "Kinterp")
    b1 = _view(b1, np.int64, 1, "b1")
    b2 = _view(b2, np.int64, 1, "b2")
    p1s = _view(p1s, np.int64, 1, "p1s")
    p2s = _view(p2s, np.int64, 1, "p2s")
    fz1 = _view(fz1, np.float64, 1, "fz1")
    fz2 = _view(fz2, np.float64, 1, "fz2")
    fzGrid = _view(fzGrid, np.float64, 1, "fzGrid")
    Kgrid = _view(Kgrid, np.float64, 4, "Kgrid")
    if NO1 * NO2 == 0:
        return
    nB, nB2, nzg, nzg2 = Kgrid.shape
    if nB != nB2 or nzg != nzg2 or fzGrid.size < nzg or Kinterp.shape[0] < NO1 or \
            Kinterp.shape[1] < NO2 or min(b1.size, fz1.size, p1s.size) < NO1 or \
            min(b2.size, fz2.size, p2s.size) < NO2:
        raise ValueError("array shapes are inconsistent")
    _check_bands(b1[:NO1], nB, "b1")
    _check_bands(b2[:NO2], nB, "b2")
    for p in (p1s[:NO1], p2s[:NO2]):
        if p.size and (p.min() < 0 or p.max() > nzg - 2):
            raise ValueError("bin positions outside [0, nz-2]")
    torch = _lib.require_cuda()
    d_k = torch.empty((NO1, NO2), dtype=torch.float64, device="cuda")
    args = [_dev(a) for a in (b1[:NO1], fz1[:NO1], p1s[:NO1], b2[:NO2], fz2[:NO2], p2s[:NO2],
                              fzGrid[:nzg], Kgrid)]
    _lib.check(_lib.lib().dl_kernel_parts_interp_dev(
        NO1, NO2, _lib.ptr(d_k), NO2, *[_lib.ptr(a) for a in args], nB, nzg,
        _lib.stream_ptr()))
    Kinterp[:NO1, :NO2] = d_k.cpu().numpy()

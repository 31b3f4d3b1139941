"""Batched, fused apply path: the B200 replacement of the per-target host loop of
scripts/delight-apply.py:114-170.

    grid  = ModelGrid.from_legacy(model_mean, model_covar)      # [nz, nT, nB] arrays of the script
    prior = grid.gaussian_prior(redshiftGrid, train_z, sigma)   # delight-apply.py:104
    out   = apply_fused(grid, prior, redshiftGrid, targets, ell_hat=1.0, ell_var=s**2, K=50)

Everything numeric happens in libdelight_b200.so; torch is used for device memory, streams and
index plumbing (gather/scatter of target rows) only.
"""
import ctypes

import numpy as np

from . import _lib

FUSED_MAX_BANDS = 32


def partition(n, rank, size):
    """Contiguous rank range with the reference's formula (scripts/delight-apply.py:50-53)."""
    first = int(rank * n / float(size))
    last = int(min(n, (rank + 1) * n / float(size)))
    return first, last


class Targets:
    """Target catalogue in packed form: per target the list of valid bands (io.py:331-337) and
    the fluxes / variances in those bands, left-packed into [nO, nB] arrays."""

    def __init__(self, fluxes, fluxes_var, bands, nf, ztrue=None):
        self.fluxes = np.ascontiguousarray(fluxes, dtype=np.float64)
        self.fluxes_var = np.ascontiguousarray(fluxes_var, dtype=np.float64)
        self.bands = np.ascontiguousarray(bands, dtype=np.int64)
        self.nf = np.ascontiguousarray(nf, dtype=np.int64)
        self.ztrue = None if ztrue is None else np.ascontiguousarray(ztrue, dtype=np.float64)
        if self.fluxes.ndim != 2 or self.fluxes.shape != self.fluxes_var.shape or \
                self.bands.shape != self.fluxes.shape or self.nf.shape != (self.fluxes.shape[0],):
            raise ValueError("Targets: fluxes, fluxes_var, bands must be [nO, nB]; nf [nO]")
        if self.nf.size and (self.nf.min() < 1 or self.nf.max() > self.fluxes.shape[1]):
            raise ValueError("Targets: nf must be in [1, nB]")
        valid = np.arange(self.fluxes.shape[1])[None, :] < self.nf[:, None]
        if not np.all(np.isfinite(self.fluxes[valid])) or \
                not np.all(self.fluxes_var[valid] > 0) or \
                not np.all(np.isfinite(self.fluxes_var[valid])):
            # the catalogue reader drops such entries (io.py:331-337); the kernels assume it
            raise ValueError("Targets: fluxes must be finite and variances finite and > 0 "
                             "in the valid bands")
        if np.any(self.bands[valid] < 0):
            raise ValueError("Targets: negative band index in a valid slot")

    @classmethod
    def full(cls, fluxes, fluxes_var, ztrue=None):
        """All targets observed in all nB bands."""
        nO, nB = np.shape(fluxes)
        return cls(fluxes, fluxes_var, np.tile(np.arange(nB), (nO, 1)), np.full(nO, nB), ztrue)

    def __len__(self):
        return self.fluxes.shape[0]

    def slice(self, first, last):
        return Targets(self.fluxes[first:last], self.fluxes_var[first:last],
                       self.bands[first:last], self.nf[first:last],
                       None if self.ztrue is None else self.ztrue[first:last])

    def to_device(self):
        """Upload fluxes / variances once (per launch group); later apply_fused calls on this
        object start from HBM-resident inputs."""
        torch = _lib.require_cuda()
        self._dev = {}
        for key, nf, union, rows, idx in self.launch_groups():
            self._dev[key] = (
                torch.from_numpy(np.ascontiguousarray(self.fluxes[idx, :nf])).cuda(),
                torch.from_numpy(np.ascontiguousarray(self.fluxes_var[idx, :nf])).cuda(),
                torch.from_numpy(idx).cuda(),
                None if rows is None else torch.from_numpy(rows).cuda())
        return self

    def groups(self):
        """Yield (band tuple, indices) for targets sharing the same band list."""
        if getattr(self, "_groups", None) is None:
            self._groups = list(self._compute_groups())
        for g in self._groups:
            yield g

    def _compute_groups(self):
        nO, nB = self.fluxes.shape
        key = np.where(np.arange(nB)[None, :] < self.nf[:, None], self.bands, -1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        for g in range(uniq.shape[0]):
            bands = tuple(int(b) for b in uniq[g] if b >= 0)
            yield bands, np.nonzero(inv == g)[0]

    def launch_groups(self, per_band_set=False):
        """Launch groups of the fused kernel: one per band COUNT.  Yields
        (key, nf, union_bands, rows, indices): `rows` is None when every target of the group is
        observed in exactly `union_bands`, else int32 [n, nf] with the position of each of the
        target's bands (io.py:331-337 valid-band list) in `union_bands`.
        per_band_set=True (or more than 32 distinct bands in a group): one group per band SET."""
        cache = "_lgroups_set" if per_band_set else "_lgroups"
        if getattr(self, cache, None) is None:
            setattr(self, cache, list(self._compute_launch_groups(per_band_set)))
        for g in getattr(self, cache):
            yield g

    def _compute_launch_groups(self, per_band_set):
        if per_band_set:
            for bands, idx in self.groups():
                yield ("set",) + bands, len(bands), bands, None, idx
            return
        for nf in np.unique(self.nf):
            nf = int(nf)
            idx = np.nonzero(self.nf == nf)[0]
            b = self.bands[idx, :nf]
            if np.all(b == b[0]):
                yield ("nf", nf), nf, tuple(int(x) for x in b[0]), None, idx
                continue
            union = np.unique(b)
            if union.size > FUSED_MAX_BANDS:
                for bands, sub in self.groups():
                    if len(bands) == nf:
                        yield ("set",) + bands, nf, bands, None, sub
                continue
            rows = np.searchsorted(union, b).astype(np.int32)
            yield ("nf", nf), nf, tuple(int(x) for x in union), np.ascontiguousarray(rows), idx


class ModelGrid:
    """Per-training-galaxy flux predictions over the fine redshift grid, resident in HBM in the
    fused kernel's layout [nT][2*nB][nzp] (rows 0..nB-1 means, nB..2nB-1 variances, z contiguous,
    nzp = nz padded to the 128-redshift chunk)."""

    def __init__(self, grid, nT, nB, nz):
        self.grid, self.nT, self.nB, self.nz = grid, int(nT), int(nB), int(nz)
        self.nzp = int(_lib.lib().dl_fused_padded_nz(self.nz))

    @classmethod
    def from_legacy(cls, model_mean, model_covar):
        """From the script's arrays model_mean/model_covar [nz, nT, nB] (delight-apply.py:84-85),
        NumPy (uploaded) or CUDA tensors."""
        torch = _lib.require_cuda()
        mm = _as_dev(model_mean)
        mc = _as_dev(model_covar)
        if mm.ndim != 3 or mm.shape != mc.shape:
            raise ValueError("model_mean / model_covar must both be [nz, nT, nB]")
        nz, nT, nB = mm.shape
        nzp = int(_lib.lib().dl_fused_padded_nz(nz))
        grid = torch.empty((nT, 2 * nB, nzp), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().dl_relayout_grid_dev(nz, nT, nB, _lib.ptr(mm), _lib.ptr(mc),
                                                  _lib.ptr(grid), _lib.stream_ptr()))
        return cls(grid, nT, nB, nz)

    def gaussian_prior(self, zgrid, train_z, sigma):
        """prior[t][z] = exp(-((z - z_t)/sigma)^2 / 2)  (delight-apply.py:104) -> [nT][nzp]."""
        torch = _lib.require_cuda()
        zg, zt = _as_dev(zgrid), _as_dev(train_z)
        if zg.numel() != self.nz or zt.numel() != self.nT:
            raise ValueError("gaussian_prior: zgrid [nz], train_z [nT]")
        prior = torch.empty((self.nT, self.nzp), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().dl_zprior_dev(self.nT, self.nz, _lib.ptr(zg), _lib.ptr(zt),
                                           float(sigma), _lib.ptr(prior), _lib.stream_ptr()))
        return prior

    def template_prior(self, zgrid, p_t, p_z_t):
        """p(z, t) = p_t z / p_z_t^2 exp(-z^2 / (2 p_z_t^2))  (scripts/templateFitting.py:79-83)
        for a grid whose nT "training galaxies" are the nt templates -> [nt][nzp]."""
        torch = _lib.require_cuda()
        zg, pt, pzt = _as_dev(zgrid), _as_dev(p_t), _as_dev(p_z_t)
        if zg.numel() != self.nz or pt.numel() != self.nT or pzt.numel() != self.nT:
            raise ValueError("template_prior: zgrid [nz], p_t [nt], p_z_t [nt]")
        prior = torch.empty((self.nT, self.nzp), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().dl_template_prior_dev(self.nT, self.nz, _lib.ptr(zg), _lib.ptr(pt),
                                                   _lib.ptr(pzt), _lib.ptr(prior),
                                                   _lib.stream_ptr()))
        return prior

    def explicit_prior(self, prior_zt):
        """From an explicit prior[nz, nT] (e.g. scripts/templateFitting.py:79-82)."""
        torch = _lib.require_cuda()
        p = _as_dev(prior_zt)
        if tuple(p.shape) != (self.nz, self.nT):
            raise ValueError("explicit_prior: expected [nz, nT]")
        prior = torch.empty((self.nT, self.nzp), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().dl_relayout_prior_dev(self.nz, self.nT, _lib.ptr(p),
                                                   _lib.ptr(prior), _lib.stream_ptr()))
        return prior


def _as_dev(a):
    torch = _lib.require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != torch.float64:
            raise ValueError("expected float64")
        return t.cuda().contiguous()
    a = np.ascontiguousarray(a, dtype=np.float64)
    if not a.flags.writeable:
        a = a.copy()
    return torch.from_numpy(a).cuda()


class FusedWorkspace:
    """Re-usable device scratch for apply_fused (the raw evidences dominate:
    8 B x tile targets x nT)."""

    def __init__(self, max_bytes=None):
        torch = _lib.require_cuda()
        if max_bytes is None:
            free, _ = torch.cuda.mem_get_info()
            max_bytes = min(int(free * 0.5), 48 << 30)
        self.max_bytes = int(max_bytes)
        self.buf = None
        self._sm = None

    def sm_count(self):
        if self._sm is None:
            sm = ctypes.c_int(0)
            _lib.check(_lib.lib().dl_device_info(ctypes.byref(sm), None, None, None))
            self._sm = sm.value
        return self._sm

    def get(self, nbytes):
        torch = _lib.require_cuda()
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
        return self.buf

    def tile_size(self, nO, nT, nz, evid_external):
        """Targets per call such that the scratch (8 B x (nzp + nT) + ~1.5 KB per target) fits."""
        L = _lib.lib()
        unit = 64
        per = L.dl_fused_workspace_bytes(unit, nT, nz, int(evid_external))
        tile = max(unit, (self.max_bytes // max(per, 1)) * unit)
        return int(min(tile, (nO + unit - 1) // unit * unit))


def _metrics(L, torch, pdf, zg, targets, confidence_levels):
    """computeMetrics on device (utils.py:213-227): 7 + nCL columns followed by the argmax index."""
    nO, nz = pdf.shape
    zt = _as_dev(targets.ztrue) if targets.ztrue is not None else None
    metrics = torch.empty((nO, 8), dtype=torch.float64, device="cuda")
    _lib.check(L.dl_metrics_dev(nO, nz, _lib.ptr(pdf), nz, _lib.ptr(zg), _lib.ptr(zt),
                                _lib.ptr(metrics), _lib.stream_ptr()))
    if confidence_levels is not None and len(confidence_levels):
        # computeMetrics: CIlevel(grid, PDF, 1.0 - confI) per level  (utils.py:223-225)
        fr = _as_dev(1.0 - np.asarray(confidence_levels, dtype=np.float64))
        ci = torch.empty((nO, fr.numel()), dtype=torch.float64, device="cuda")
        _lib.check(L.dl_ci_levels_dev(nO, nz, _lib.ptr(pdf), nz, _lib.ptr(zg), fr.numel(),
                                      _lib.ptr(fr), 200, _lib.ptr(ci), _lib.stream_ptr()))
        metrics = torch.cat([metrics[:, :7], ci, metrics[:, 7:]], dim=1)
    return metrics


def apply_fused_to_host(grid, prior, zgrid, targets, host_out, chunk=None, **kw):
    """apply_fused for outputs that are large next to the compute (config 2: 12 KB of PDF per
    target against 0.2 us of kernel time): the targets are processed in chunks and the
    device-to-host copies of one chunk (into the caller's PINNED tensors `host_out[name]`, rows
    in target order; names as returned by apply_fused) overlap the kernels of the next on a
    second stream.  Returns host_out."""
    torch = _lib.require_cuda()
    nO = len(targets)
    chunk = int(chunk or max(4096, -(-nO // 8)))
    kw = dict(kw, out_device=True)
    kw.setdefault("workspace", FusedWorkspace())
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    keep = []
    for first in range(0, nO, chunk):
        last = min(nO, first + chunk)
        res = apply_fused(grid, prior, zgrid, targets.slice(first, last), **kw)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            for name, dst in host_out.items():
                src = res.get(name)
                if src is not None:
                    dst[first:last].copy_(src, non_blocking=True)
                    src.record_stream(side)
        keep.append(res)
    side.synchronize()
    return host_out


def apply_fused(grid, prior, zgrid, targets, ell_hat=1.0, ell_var=1.0, K=0, want_evid=False,
                want_metrics=False, workspace=None, out_device=False,
                confidence_levels=None, cov_scale=1.0):
    """Likelihood + PDF + evidences + top-K for all `targets` against the whole model grid.

    ell_hat / ell_var: scalars (what delight-apply.py:119-144 passes: 1 and ellPriorSigma^2) or
    arrays [nz] (the general operator signature, utils_cy.pyx:92-93).
    cov_scale: factor on the model variance (delight-optimize.py:122 passes V_C * model_covar).
    confidence_levels (with want_metrics): the cfg's confidenceLevels; the metrics then carry the
    reference's 7 + nCL columns of computeMetrics (utils.py:213-227) followed by the argmax index.
    Targets with different valid-band lists (io.py:331-337) share a launch as long as they have
    the same NUMBER of bands: one kernel launch per band count, not per band set.
    Returns dict(pdf [nO, nz], top_idx [nO, K] int64, top_ev [nO, K], evid [nO, nT] | None,
    metrics [nO, 8 (+ nCL)] | None) as NumPy arrays (or CUDA tensors with out_device=True).
    """
    torch = _lib.require_cuda()
    L = _lib.lib()
    nO = len(targets)
    nT, nB, nz = grid.nT, grid.nB, grid.nz
    valid = np.arange(targets.bands.shape[1])[None, :] < targets.nf[:, None]
    if nO and targets.bands[valid].max() >= nB:
        raise ValueError("target band index outside the model grid")
    if K > nT:
        raise ValueError("Ncompress (K=%d) exceeds the number of training galaxies (%d)" % (K, nT))
    if nO == 0:                     # empty catalogue slice (e.g. more ranks than targets)
        empty = lambda *shape, dt=torch.float64: torch.empty(shape, dtype=dt, device="cuda")
        out = dict(pdf=empty(0, nz), top_idx=empty(0, K, dt=torch.int64) if K else None,
                   top_ev=empty(0, K) if K else None, evid=empty(0, nT) if want_evid else None,
                   metrics=empty(0, 8 + (len(confidence_levels) if confidence_levels is not None
                                         else 0)) if want_metrics else None, fused_launches=0)
        if not out_device:
            out = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v)
                   for k, v in out.items()}
        return out
    ws = workspace or FusedWorkspace()
    zg = _as_dev(zgrid)
    zdep = np.ndim(ell_hat) > 0 or np.ndim(ell_var) > 0
    if zdep:
        eh = _as_dev(np.broadcast_to(np.asarray(ell_hat, dtype=np.float64), (nz,)))
        ev = _as_dev(np.broadcast_to(np.asarray(ell_var, dtype=np.float64), (nz,)))
        ehs = evs = 1.0
    else:
        eh = ev = None
        ehs, evs = float(ell_hat), float(ell_var)

    pdf = torch.empty((nO, nz), dtype=torch.float64, device="cuda")
    top_idx = torch.empty((nO, K), dtype=torch.int64, device="cuda") if K > 0 else None
    top_ev = torch.empty((nO, K), dtype=torch.float64, device="cuda") if K > 0 else None
    evid = torch.empty((nO, nT), dtype=torch.float64, device="cuda") if want_evid else None
    launches = 0

    # per-target band lists are not combined with z-dependent ell priors in the kernel
    for key, nf, union, rows, idx in targets.launch_groups(per_band_set=zdep):
        if nf > FUSED_MAX_BANDS:
            raise _lib.DelightB200Error("fused likelihood supports up to %d bands per target, "
                                        "got %d" % (FUSED_MAX_BANDS, nf))
        single = idx.size == nO
        cached = getattr(targets, "_dev", None)
        if cached is not None and key in cached:
            obs, var, didx, drows = cached[key]
        else:
            obs_h = targets.fluxes[idx, :nf] if not single else targets.fluxes[:, :nf]
            var_h = targets.fluxes_var[idx, :nf] if not single else targets.fluxes_var[:, :nf]
            obs = torch.from_numpy(np.ascontiguousarray(obs_h)).cuda(non_blocking=True)
            var = torch.from_numpy(np.ascontiguousarray(var_h)).cuda(non_blocking=True)
            didx = None
            drows = None if rows is None else torch.from_numpy(rows).cuda(non_blocking=True)
        union_c = (ctypes.c_int * len(union))(*union)
        n_g = idx.size
        if single:
            g_pdf, g_idx, g_ev, g_evid = pdf, top_idx, top_ev, evid
        else:
            g_pdf = torch.empty((n_g, nz), dtype=torch.float64, device="cuda")
            g_idx = torch.empty((n_g, K), dtype=torch.int64, device="cuda") if K > 0 else None
            g_ev = torch.empty((n_g, K), dtype=torch.float64, device="cuda") if K > 0 else None
            g_evid = torch.empty((n_g, nT), dtype=torch.float64, device="cuda") \
                if want_evid else None
        ext = int(want_evid or K == 0)
        tile = ws.tile_size(n_g, nT, nz, ext)
        for t0 in range(0, n_g, tile):
            n = min(tile, n_g - t0)
            nbytes = L.dl_fused_workspace_bytes(n, nT, nz, ext)
            buf = ws.get(nbytes)
            _lib.check(L.dl_like_fused_dev(
                n, nT, nz, nB, nf, len(union), union_c,
                _lib.ptr(drows[t0:t0 + n]) if drows is not None else None,
                _lib.ptr(obs[t0:t0 + n]), _lib.ptr(var[t0:t0 + n]),
                _lib.ptr(grid.grid), _lib.ptr(prior), _lib.ptr(zg), ehs, evs, _lib.ptr(eh),
                _lib.ptr(ev), float(cov_scale), K, _lib.ptr(g_pdf[t0:t0 + n]), nz,
                _lib.ptr(g_evid[t0:t0 + n]) if g_evid is not None else None,
                _lib.ptr(g_idx[t0:t0 + n]) if K > 0 else None,
                _lib.ptr(g_ev[t0:t0 + n]) if K > 0 else None,
                _lib.ptr(buf), buf.numel(), _lib.stream_ptr()))
            launches += 1
        if not single:
            if didx is None:
                didx = torch.from_numpy(idx).cuda()
            pdf.index_copy_(0, didx, g_pdf)
            if K > 0:
                top_idx.index_copy_(0, didx, g_idx)
                top_ev.index_copy_(0, didx, g_ev)
            if want_evid:
                evid.index_copy_(0, didx, g_evid)

    metrics = _metrics(L, torch, pdf, zg, targets, confidence_levels) if want_metrics else None
    out = dict(pdf=pdf, top_idx=top_idx, top_ev=top_ev, evid=evid, metrics=metrics,
               fused_launches=launches)
    if not out_device:
        for k, v in list(out.items()):
            if isinstance(v, torch.Tensor):
                out[k] = v.cpu().numpy()
    return out


def apply_compressed(grid, prior, zgrid, targets, indices, ell_hat=1.0, ell_var=1.0,
                     normalized=False, want_metrics=False, confidence_levels=None,
                     cov_scale=1.0, out_device=False):
    """Second ("compressed") pass of delight-apply.py:109-135: every target is compared only with
    the training galaxies whose indices the first pass stored (`indices` [nO, K], the rows of
    compressIndicesFile).  Likelihood = the reference's NumPy twin utils.py:115-176 with
    marginalizeEll / renormalize and `normalized` as given (the script passes False), times the
    prior; pdf = sum over the selected galaxies (:145), evidences = trapz_z (:146).

    Like `np.in1d(targetIndices, indices)` (:125) the selection is a set: duplicates count once,
    entries that are no training index are ignored, columns are taken in ascending index order.
    Returns dict(pdf [nO, nz], sel [nO, K] int64 (ascending, -1 padded), evid [nO, K] (aligned
    with sel), metrics | None).
    """
    torch = _lib.require_cuda()
    L = _lib.lib()
    nO = len(targets)
    nT, nB, nz = grid.nT, grid.nB, grid.nz
    indices = np.asarray(indices)
    if indices.ndim != 2 or indices.shape[0] != nO:
        raise ValueError("indices must be [n_targets, K]")
    if not np.issubdtype(indices.dtype, np.integer):
        raise ValueError("indices must be integers")
    K = indices.shape[1]
    if K < 1:
        raise ValueError("need at least one stored index per target")
    valid = np.arange(targets.bands.shape[1])[None, :] < targets.nf[:, None]
    if nO and targets.bands[valid].max() >= nB:
        raise ValueError("target band index outside the model grid")
    # the in1d mask as a sorted, de-duplicated, -1 padded index list per target
    big = np.iinfo(np.int64).max
    sel = np.where((indices >= 0) & (indices < nT), indices, big).astype(np.int64)
    sel.sort(axis=1)
    dup = np.zeros_like(sel, dtype=bool)
    dup[:, 1:] = sel[:, 1:] == sel[:, :-1]
    sel[dup] = big
    sel.sort(axis=1)
    sel[sel == big] = -1
    zg = _as_dev(zgrid)
    zdep = np.ndim(ell_hat) > 0 or np.ndim(ell_var) > 0
    if zdep:
        eh = _as_dev(np.broadcast_to(np.asarray(ell_hat, dtype=np.float64), (nz,)))
        ev = _as_dev(np.broadcast_to(np.asarray(ell_var, dtype=np.float64), (nz,)))
        ehs = evs = 1.0
    else:
        eh = ev = None
        ehs, evs = float(ell_hat), float(ell_var)
    pdf = torch.empty((nO, nz), dtype=torch.float64, device="cuda")
    evid = torch.zeros((nO, K), dtype=torch.float64, device="cuda")
    sel_d = torch.from_numpy(sel).cuda()
    for bands, idx in (targets.groups() if nO else []):
        nf = len(bands)
        single = idx.size == nO
        obs = torch.from_numpy(np.ascontiguousarray(targets.fluxes[idx, :nf])).cuda()
        var = torch.from_numpy(np.ascontiguousarray(targets.fluxes_var[idx, :nf])).cuda()
        didx = None if single else torch.from_numpy(idx).cuda()
        g_sel = sel_d if single else sel_d.index_select(0, didx).contiguous()
        g_pdf = pdf if single else torch.empty((idx.size, nz), dtype=torch.float64, device="cuda")
        g_ev = evid if single else torch.empty((idx.size, K), dtype=torch.float64, device="cuda")
        bands_c = (ctypes.c_int * nf)(*bands)
        _lib.check(L.dl_like_compressed_dev(
            idx.size, nT, nz, nB, nf, bands_c, _lib.ptr(obs), _lib.ptr(var), _lib.ptr(grid.grid),
            _lib.ptr(prior), _lib.ptr(zg), ehs, evs, _lib.ptr(eh), _lib.ptr(ev), float(cov_scale),
            K, _lib.ptr(g_sel), int(bool(normalized)), _lib.ptr(g_pdf), nz, _lib.ptr(g_ev),
            _lib.stream_ptr()))
        if not single:
            pdf.index_copy_(0, didx, g_pdf)
            evid.index_copy_(0, didx, g_ev)
    metrics = _metrics(L, torch, pdf, zg, targets, confidence_levels) \
        if want_metrics and nO else None
    out = dict(pdf=pdf, sel=sel_d, evid=evid, metrics=metrics)
    if not out_device:
        for k, v in list(out.items()):
            if isinstance(v, torch.Tensor):
                out[k] = v.cpu().numpy()
    return out

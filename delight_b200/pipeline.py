"""File-level drivers of the batched path: the same parameter file, catalogues and output files as
scripts/delight-learn.py and scripts/delight-apply.py, with the per-galaxy Python loops replaced
by the batched device calls of `gp` / `apply` and targets sharded over ranks (torchrun).

    python -m delight_b200.pipeline learn        params.cfg
    python -m delight_b200.pipeline apply        params.cfg
    python -m delight_b200.pipeline templatefit  params.cfg       (scripts/templateFitting.py)
    torchrun --nproc-per-node 8 -m delight_b200.pipeline apply params.cfg

Host-side format code only (the numeric work is in libdelight_b200.so).  The readers restate the
semantics of delight/io.py for the keys and columns this path uses:
  * parameter file: the [Bands]/[Templates]/[Training]/[Target]/[Other] keys of io.py:12-190;
  * catalogue rows: io.py:305-351 -- per object the valid bands are those with finite, positive
    flux AND variance; objects with a bad reference flux, z < 0 (also: no redshift column) or
    fewer than two valid bands are skipped; variance += (training_extraFracFluxError * flux)^2
    for targets too (the reference uses the training key for both prefixes, io.py:350-351);
  * gpparams rows: `%i` + `%.12e` columns (delight-learn.py:48,130);
  * outputs: `%.2e` PDFs / metrics / evidences, `%i` indices (delight-apply.py:222-233);
  * any catalogue / gpparams / output path ending in `.npy` is read / written as a binary NumPy
    array with the same rows and columns (text parsing dominates at 10^6 objects).
Rows of skipped objects stay zero at the END of the output arrays, as in the reference (its
`loc` counter only advances on yielded objects).
"""
import configparser
import os
import sys

import numpy as np

from . import apply as _apply
from . import distributed as _dist
from . import gp as _gp


# ------------------------------------------------------------------------------------------------
# parameter file (subset of delight/io.py:12-190 used by learn / apply)
# ------------------------------------------------------------------------------------------------
def parse_params(path):
    cfg = configparser.ConfigParser()
    if not os.path.isfile(path):
        raise Exception(path + " : file not found")
    cfg.read(path)
    for sec in ("Bands", "Templates", "Training", "Target", "Other"):
        if not cfg.has_section(sec):
            raise Exception(sec + " not found in parameter file")
    floats = lambda s: [float(x) for x in s.split(" ") if x]
    p = {}
    p["bands_directory"] = cfg.get("Bands", "directory")
    p["bandNames"] = cfg.get("Bands", "Names").split(" ")
    p["templates_directory"] = cfg.get("Templates", "directory")
    p["templates_names"] = cfg.get("Templates", "names").split(" ")
    p["p_t"] = np.array(floats(cfg.get("Templates", "p_t")))
    p["p_z_t"] = np.array(floats(cfg.get("Templates", "p_z_t")))
    for pre, sec in (("training_", "Training"), ("target_", "Target")):
        p[pre + "catFile"] = cfg.get(sec, "catFile")
        p[pre + "referenceBand"] = cfg.get(sec, "referenceBand")
        p[pre + "bandOrder"] = cfg.get(sec, "bandOrder").split(" ")
        p[pre + "extraFracFluxError"] = cfg.getfloat(sec, "extraFracFluxError")
    p["training_paramFile"] = cfg.get("Training", "paramFile")
    p["training_numChunks"] = cfg.getint("Training", "numChunks")
    # cross-validation of delight-learn (io.py:104-108)
    p["training_crossValidate"] = cfg.getboolean("Training", "crossValidate", fallback=False)
    p["training_CV_bandOrder"] = cfg.get("Training", "crossValidationBandOrder",
                                         fallback="").split(" ")
    p["training_CVfile"] = cfg.get("Training", "CVfile", fallback="")
    for k in ("compressIndicesFile", "compressMargLikFile", "redshiftpdfFile", "metricsFile"):
        p[k] = cfg.get("Target", k)
    for k in ("redshiftpdfFileTemp", "metricsFileTemp"):        # templateFitting.py:128-131
        p[k] = cfg.get("Target", k, fallback="")
    p["Ncompress"] = cfg.getint("Target", "Ncompress")
    p["useCompression"] = cfg.getboolean("Target", "useCompression")
    p["compressionFilesFound"] = os.path.isfile(p["compressIndicesFile"]) and \
        os.path.isfile(p["compressMargLikFile"])
    for k in ("zPriorSigma", "ellPriorSigma", "alpha_C", "alpha_L", "V_C", "V_L", "redshiftMin",
              "redshiftMax", "redshiftBinSize"):
        p[k] = cfg.getfloat("Other", k)
    p["redshiftNumBinsGPpred"] = cfg.getint("Other", "redshiftNumBinsGPpred")
    p["lines_pos"] = floats(cfg.get("Other", "lines_pos"))
    p["confidenceLevels"] = floats(cfg.get("Other", "confidenceLevels"))
    return p


def read_band_coefficients(p):
    """io.readBandCoefficients (io.py:219-238)."""
    amp, mu, sig = [], [], []
    for band in p["bandNames"]:
        d = np.loadtxt(os.path.join(p["bands_directory"], band + "_gaussian_coefficients.txt"))
        amp.append(d[:, 0]); mu.append(d[:, 1]); sig.append(d[:, 2])
    return np.vstack(amp), np.vstack(mu), np.vstack(sig)


def create_grids(p):
    """io.createGrids (io.py:241-260): fine linear grid, log-spaced GP grid."""
    zgrid = np.arange(p["redshiftMin"], p["redshiftMax"], p["redshiftBinSize"])
    zgp = np.logspace(np.log10(p["redshiftMin"]), np.log10(p["redshiftMax"] * 1.01),
                      p["redshiftNumBinsGPpred"])
    return zgrid, zgp


def read_template_tables(p):
    """The <sed>_fluxredshiftmod.txt tables behind io.readSEDs (io.py:263-277) -> [nt, nB, nz]."""
    return np.stack([np.loadtxt(os.path.join(p["templates_directory"],
                                             s + "_fluxredshiftmod.txt")).T
                     for s in p["templates_names"]])


def column_positions(p, prefix):
    """io.readColumnPositions (io.py:193-216)."""
    order = p[prefix + "bandOrder"]
    band_idx = np.array([ib for ib, b in enumerate(p["bandNames"]) if b in order])
    names = [p["bandNames"][i] for i in band_idx]
    cols = np.array([order.index(b) for b in names])
    vcols = np.array([order.index(b + "_var") for b in names])
    zcol = order.index("redshift") if "redshift" in order else -1
    refcol = order.index(p[prefix + "referenceBand"])
    return band_idx, cols, vcols, zcol, refcol


def read_catalogue(p, prefix, first=0, last=None, cv=False):
    """Catalogue rows [first, last) -> packed apply.Targets (io.py:305-351 semantics).
    Returns (targets, n_lines) where n_lines counts file rows in the range; with cv=True also
    the cross-validation columns of the kept rows (io.py:353-362): a dict with band_idx [nCV]
    (indices into bandNames), mask / flux / var [n, nCV]."""
    band_idx, cols, vcols, zcol, refcol = column_positions(p, prefix)
    data = _load(p[prefix + "catFile"])[first:last]
    n_lines = data.shape[0]
    flux, var = data[:, cols], data[:, vcols]
    mask = np.isfinite(flux) & np.isfinite(var) & (flux > 0.0) & (var > 0.0)
    z = data[:, zcol] if zcol >= 0 else np.full(n_lines, -1.0)
    ref = data[:, refcol]
    keep = ~((ref <= 0) | ~np.isfinite(ref) | (z < 0) | (mask.sum(axis=1) <= 1))
    flux, var, mask, z = flux[keep], var[keep], mask[keep], z[keep]
    n, nb = flux.shape
    nf = mask.sum(axis=1)
    order = np.argsort(~mask, axis=1, kind="stable")          # valid bands first, in band order
    take = lambda a: np.take_along_axis(a, order, axis=1)
    valid = np.arange(nb)[None, :] < nf[:, None]
    fl = np.where(valid, take(flux), 0.0)
    fv = np.where(valid, take(var), 0.0)
    fv = np.where(valid, fv + (p["training_extraFracFluxError"] * fl) ** 2, 0.0)
    bands = np.where(valid, take(np.broadcast_to(band_idx, flux.shape)), 0)
    # left-pack into nB-wide arrays (nB = bands of the model)
    nB = len(p["bandNames"])
    pad = lambda a, dt: np.concatenate([a, np.zeros((n, nB - nb), dtype=dt)], axis=1)
    cat = _apply.Targets(pad(fl, float), pad(fv, float), pad(bands, np.int64), nf, z)
    if not cv:
        return cat, n_lines
    order = p[prefix + "CV_bandOrder"]                     # readColumnPositions(refFlux=False)
    cv_idx = np.array([ib for ib, b in enumerate(p["bandNames"]) if b in order], dtype=np.int64)
    names = [p["bandNames"][i] for i in cv_idx]
    ccols = np.array([order.index(b) for b in names], dtype=np.int64)
    cvcols = np.array([order.index(b + "_var") for b in names], dtype=np.int64)
    cdata = data[keep]
    cfl = cdata[:, ccols] if ccols.size else np.zeros((n, 0))
    cfv = cdata[:, cvcols] if ccols.size else np.zeros((n, 0))
    cmask = np.isfinite(cfl) & np.isfinite(cfv) & (cfl > 0.0) & (cfv > 0.0)
    with np.errstate(invalid="ignore"):
        cfv = cfv + (p["training_extraFracFluxError"] * cfl) ** 2
    return cat, n_lines, dict(band_idx=cv_idx, mask=cmask, flux=cfl, var=cfv)


def make_gp(p):
    amp, mu, sig = read_band_coefficients(p)
    zgrid, zgp = create_grids(p)
    tmpl = read_template_tables(p)
    gpm = _gp.PhotozGPBatched(tmpl, zgrid, amp, mu, sig, p["lines_pos"], p["V_C"], p["V_L"],
                              p["alpha_C"], p["alpha_L"], zgp)
    return gpm, zgrid


# ------------------------------------------------------------------------------------------------
# learn  (scripts/delight-learn.py)
# ------------------------------------------------------------------------------------------------
def learn(cfg_path, write=True):
    """scripts/delight-learn.py.  With `crossValidate: True` the GP of every training galaxy is
    also evaluated at its own redshift and the per-band terms
    -0.5 (model - flux)^2 / (model_var + flux_var) of the held-out bands are written to
    training_CVfile (delight-learn.py:92-100,131-132), row = catalogue line of the galaxy's
    position among the valid ones (`firstLine + loc`), one column per cross-validation band."""
    p = parse_params(cfg_path)
    rank, size = _dist.init()
    gpm, _ = make_gp(p)
    n_total = _count_rows(p["training_catFile"])
    first, last = _dist.partition(n_total, rank, size)            # delight-learn.py:31-33
    cv = p["training_crossValidate"]
    got = read_catalogue(p, "training_", first, last, cv=cv)
    cat, n_lines = got[0], got[1]
    import torch
    rows = torch.zeros((n_lines, gpm.numCol), dtype=torch.float64, device="cuda")
    if len(cat):
        rows[:len(cat)] = gpm.fit(cat, out_device=True)
    chi2 = None
    if cv:
        cvd = got[2]
        ncv = cvd["band_idx"].size
        chi2_local = np.zeros((n_lines, ncv))
        if len(cat) and ncv:
            mean, var = gpm.predict_at(rows[:len(cat)], cat.ztrue)       # :92-94
            m, v = mean[:, cvd["band_idx"]], var[:, cvd["band_idx"]]
            with np.errstate(invalid="ignore", divide="ignore"):
                term = -0.5 * (m - cvd["flux"]) ** 2 / (v + cvd["var"])  # :98-100
            chi2_local[:len(cat)] = np.where(cvd["mask"], term, 0.0)
        # the reference sums zero-padded per-rank arrays (Allreduce SUM, :112): same as a gather
        chi2 = _dist.gather_rows(torch.from_numpy(chi2_local).cuda(), n_total)
    rows = _dist.gather_rows(rows, n_total)                       # delight-learn.py:125
    if rank == 0:
        rows = rows.cpu().numpy()
        if write:
            fmt = "%i " + "%.12e " * (rows.shape[1] - 1)          # delight-learn.py:48
            _save(p["training_paramFile"], rows, fmt)
            if cv:
                _save(p["training_CVfile"], chi2.cpu().numpy(), "%.18e")   # np.savetxt default
    if cv:
        return (rows, chi2.cpu().numpy()) if rank == 0 else None
    return rows if rank == 0 else None


# ------------------------------------------------------------------------------------------------
# apply  (scripts/delight-apply.py)
# ------------------------------------------------------------------------------------------------
def apply(cfg_path, write=True, cap_training_like_reference=True):
    """cap_training_like_reference: delight-apply.py:79-80 bounds the training rows it loads by
    the number of TARGET objects (`TR_lastLine` uses numObjectsTarget); keep that for parity."""
    p = parse_params(cfg_path)
    if p["training_numChunks"] != 1:
        raise NotImplementedError("numChunks > 1 normalises each chunk separately in the "
                                  "reference (SURVEY appendix A); only numChunks = 1")
    compressed = p["useCompression"] and p["compressionFilesFound"]     # delight-apply.py:109
    rank, size = _dist.init()
    import torch
    gpm, zgrid = make_gp(p)
    n_train = _count_rows(p["training_paramFile"])
    n_target = _count_rows(p["target_catFile"])
    tr_last = int(min(n_train, n_target)) if cap_training_like_reference else n_train
    if tr_last < n_train and rank == 0:
        print("delight_b200.pipeline.apply: only the first %d of %d training rows are used, like "
              "the reference (delight-apply.py:79-80 bounds the training rows by the number of "
              "TARGET objects); pass cap_training_like_reference=False to use all of them"
              % (tr_last, n_train), file=sys.stderr)
    rows = _load(p["training_paramFile"])[:tr_last]
    # rows of skipped training galaxies are all zero (they sit at the end of every learn rank's
    # block); they are left out of the model grid, and `line_of` maps the remaining ones back to
    # line numbers of the gpparams file, which is what the compress files hold
    # (targetIndices[sortind], delight-apply.py:150-153)
    line_of = np.nonzero(rows[:, 0] >= 1)[0].astype(np.int64)
    rows = rows[line_of]
    ncol = gpm.numCol
    if rows.shape[1] < ncol:
        raise ValueError("gpparams file has too few columns for this band / template set")
    grid, train_z = gpm.predict(np.ascontiguousarray(rows[:, :ncol]))
    prior = grid.gaussian_prior(zgrid, train_z, p["zPriorSigma"])
    first, last = _dist.partition(n_target, rank, size)           # delight-apply.py:50-53
    targets, n_lines = read_catalogue(p, "target_", first, last)
    K = p["Ncompress"]
    ncl = len(p["confidenceLevels"])
    with_z = "redshift" in p["target_bandOrder"]
    if compressed:
        # second pass: only the training galaxies stored by the first one (:123-135)
        stored = np.atleast_2d(_load(p["compressIndicesFile"]))[first:last].astype(np.int64)
        to_grid = np.full(tr_last + 1, -1, dtype=np.int64)         # file line -> grid row
        to_grid[line_of] = np.arange(line_of.size)
        stored = to_grid[np.where((stored >= 0) & (stored < tr_last), stored, tr_last)]
        res = _apply.apply_compressed(grid, prior, zgrid, targets, stored[:len(targets)],
                                      ell_hat=1.0, ell_var=p["ellPriorSigma"] ** 2,
                                      normalized=False, want_metrics=with_z,
                                      confidence_levels=p["confidenceLevels"], out_device=True)
        res["top_idx"], res["top_ev"] = res["sel"], res["evid"]
    else:
        res = _apply.apply_fused(grid, prior, zgrid, targets, ell_hat=1.0,
                                 ell_var=p["ellPriorSigma"] ** 2, K=K, want_metrics=with_z,
                                 confidence_levels=p["confidenceLevels"], out_device=True)

    if res["top_idx"] is not None and res["top_idx"].numel():
        # grid rows -> gpparams line numbers (-1 padding of the compressed pass stays)
        lo = torch.from_numpy(line_of).cuda()
        ti = res["top_idx"]
        res["top_idx"] = torch.where(ti >= 0, lo[ti.clamp(min=0)], ti)

    def full(a, width, dt=torch.float64):
        out = torch.zeros((n_lines, width), dtype=dt, device="cuda")
        if a is not None and a.shape[0]:
            out[:a.shape[0]] = a[:, :width]
        return out
    pdfs = _dist.gather_rows(full(res["pdf"], zgrid.size), n_target)          # :201
    idx = _dist.gather_rows(full(res["top_idx"], K, torch.int64), n_target)   # :206-210
    ev = _dist.gather_rows(full(res["top_ev"], K), n_target)
    if with_z and res["metrics"] is not None and res["metrics"].shape[0]:
        # the reference only fills the metrics of objects whose PDF has mass (:164-166)
        res["metrics"][res["pdf"].sum(dim=1) <= 0] = 0.0
    met = _dist.gather_rows(full(res["metrics"], 7 + ncl), n_target) if with_z else None  # :216
    if rank != 0:
        return None
    out = dict(pdfs=pdfs.cpu().numpy(), top_idx=idx.cpu().numpy(), top_ev=ev.cpu().numpy(),
               metrics=met.cpu().numpy() if with_z else None)
    if write:                                                      # delight-apply.py:222-233
        _save(p["redshiftpdfFile"], out["pdfs"], "%.2e")
        if with_z:
            _save(p["metricsFile"], out["metrics"], "%.2e")
        if p["useCompression"] and not p["compressionFilesFound"]:
            _save(p["compressMargLikFile"], out["top_ev"], "%.2e")
            _save(p["compressIndicesFile"], out["top_idx"], "%i")
    return out


# ------------------------------------------------------------------------------------------------
# hyper-parameter sweep  (scripts/delight-optimize.py:52-128)
# ------------------------------------------------------------------------------------------------
def sweep(cfg_path, ellPriorSigmas=(1.0, 10.0), V_Cs=(0.1, 1.0), zPriorSigmas=(0.1, 1.0)):
    """The likelihood part of delight-optimize.py as a multi-configuration run on one resident
    model grid: for every (ellPriorSigma, V_C, zPriorSigma) one fused pass over all targets with
    `V_C * model_covar` (:122), prior width zPriorSigma (:124-126) and the full computeMetrics
    row per target (:131-134).  As in the reference, V_C only rescales the model variance (the
    script sets `gp.var_C`, which the kernel object never reads) and the extraFracFluxError loop
    has no effect on the catalogue reader, so the grid is fitted and predicted once.
    Returns {(ellPriorSigma, V_C, zPriorSigma): metrics [n_valid_targets, 7 + nCL]} on rank 0
    (targets sharded over ranks)."""
    p = parse_params(cfg_path)
    rank, size = _dist.init()
    gpm, zgrid = make_gp(p)
    n_train = _count_rows(p["training_catFile"])
    cat, _ = read_catalogue(p, "training_", 0, n_train)
    grid, train_z = gpm.predict(gpm.fit(cat, out_device=True))
    n_target = _count_rows(p["target_catFile"])
    first, last = _dist.partition(n_target, rank, size)
    targets, n_lines = read_catalogue(p, "target_", first, last)
    targets.to_device()
    ws = _apply.FusedWorkspace()
    import torch
    out = {}
    for zs in zPriorSigmas:
        prior = grid.gaussian_prior(zgrid, train_z, zs)
        for es in ellPriorSigmas:
            for vc in V_Cs:
                res = _apply.apply_fused(grid, prior, zgrid, targets, ell_hat=1.0,
                                         ell_var=es ** 2, K=0, want_metrics=True,
                                         confidence_levels=p["confidenceLevels"], cov_scale=vc,
                                         workspace=ws, out_device=True)
                m = res["metrics"][:, :7 + len(p["confidenceLevels"])]
                # a validity column travels with the rows: zero padding of skipped catalogue
                # lines is told apart by it, not by a metric value that may itself be 0
                pad = torch.zeros((n_lines, m.shape[1] + 1), dtype=torch.float64, device="cuda")
                pad[:m.shape[0], :-1] = m
                pad[:m.shape[0], -1] = 1.0
                g = _dist.gather_rows(pad, n_target)
                if rank == 0:
                    g = g.cpu().numpy()
                    out[(es, vc, zs)] = g[g[:, -1] == 1.0, :-1]
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------
# template fitting  (scripts/templateFitting.py)
# ------------------------------------------------------------------------------------------------
def template_fit(cfg_path, write=True):
    """scripts/templateFitting.py on the fused kernel: the nt templates play the role of training
    galaxies with zero model variance (utils.py:141-146 then runs a single iteration, which is
    what the kernel's second pass reduces to), ell prior (1, 1e24) (:71-78), the analytic p(z,t)
    prior of :79-83 computed on the device, PDFs = sum over templates (:84), computeMetrics for
    every object with mass (:85-89).  Writes redshiftpdfFileTemp and, when the catalogue has
    redshifts, metricsFileTemp (:128-131).  Targets are sharded over ranks."""
    p = parse_params(cfg_path)
    rank, size = _dist.init()
    import torch
    zgrid, _ = create_grids(p)
    tmpl = read_template_tables(p)                                 # [nt, nB, nz]
    f_mod = np.ascontiguousarray(tmpl.transpose(2, 0, 1))          # templateFitting.py:38-43
    grid = _apply.ModelGrid.from_legacy(f_mod, np.zeros_like(f_mod))
    prior = grid.template_prior(zgrid, p["p_t"], p["p_z_t"])
    n_target = _count_rows(p["target_catFile"])
    first, last = _dist.partition(n_target, rank, size)           # templateFitting.py:46-49
    targets, n_lines = read_catalogue(p, "target_", first, last)
    ncl = len(p["confidenceLevels"])
    res = _apply.apply_fused(grid, prior, zgrid, targets, ell_hat=1.0, ell_var=(1e12) ** 2, K=0,
                             want_metrics=True, confidence_levels=p["confidenceLevels"],
                             out_device=True)
    if res["metrics"] is not None and res["metrics"].shape[0]:
        res["metrics"][res["pdf"].sum(dim=1) <= 0] = 0.0          # :85

    def full(a, width):
        out = torch.zeros((n_lines, width), dtype=torch.float64, device="cuda")
        if a is not None and a.shape[0]:
            out[:a.shape[0]] = a[:, :width]
        return out
    pdfs = _dist.gather_rows(full(res["pdf"], zgrid.size), n_target)          # :115
    met = _dist.gather_rows(full(res["metrics"], 7 + ncl), n_target)          # :122
    if rank != 0:
        return None
    out = dict(pdfs=pdfs.cpu().numpy(), metrics=met.cpu().numpy())
    if write:
        _save(p["redshiftpdfFileTemp"], out["pdfs"], "%.2e")
        if "redshift" in p["target_bandOrder"]:
            _save(p["metricsFileTemp"], out["metrics"], "%.2e")
    return out


def _save(path, arr, fmt):
    """Text in the reference's format, or binary when the cfg names a .npy file."""
    if path.endswith(".npy"):
        np.save(path, arr)
    else:
        np.savetxt(path, arr, fmt=fmt)


def _count_rows(path):
    if path.endswith(".npy"):
        return int(np.load(path, mmap_mode="r").shape[0])
    return sum(1 for _ in open(path))


def _load(path):
    return np.load(path) if path.endswith(".npy") else np.loadtxt(path, ndmin=2)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    cmds = {"learn": learn, "apply": apply, "templatefit": template_fit}
    if len(argv) != 2 or argv[0] not in cmds:
        raise SystemExit("usage: python -m delight_b200.pipeline {learn|apply|templatefit} "
                         "params.cfg")
    print("--- DELIGHT-%s (b200) ---" % argv[0].upper())
    cmds[argv[0]](argv[1])


if __name__ == "__main__":
    main()

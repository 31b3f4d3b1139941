#!/usr/bin/env python
"""Headline benchmark: galaxy-PDFs/s through the fused photo-z likelihood (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[2] -- 10k targets x 10k training GPs x 1500
redshifts x 6 bands (LSST ugrizy), fused PDF + top-50, float64 -- the largest configuration that
is quoted per single B200; per GPU when N > 1 (weak scaling: every rank owns 10k targets, the
model grid is replicated, no data-path collective; results are gathered to rank 0 over NCCL).
A "step" is one pass of the hot path over the rank's target batch.

One JSON line is printed by rank 0 (see the keys below).  `value` = targets/s with inputs
resident in HBM; `e2e` = the same through the public API with host buffers (H2D of the target
fluxes from pinned memory and D2H of PDFs / top-K inside the timed region).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_TARGETS = 10000
N_TRAIN = 10000
N_COMPRESS = 50
ELL_VAR = 0.25            # ellPriorSigma 0.5 (tests/parametersTest.cfg:57)
Z_PRIOR_SIGMA = 0.5
FLOP_PER_CELL = lambda nf: 27 * nf + 30     # SURVEY.md 8(d), fused path
KERNEL_SOURCES = ["delight_b200/csrc/like_fused_kernel.cuh", "delight_b200/csrc/like_fused.cu",
                  "delight_b200/csrc/common.cuh"]
TRAFFIC_RECORD = os.path.join(REPO, "profiles", "r2_fused_traffic.json")


def kernel_source_hash():
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(REPO, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(nO, nT, nz, nB):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE main-pass launch of the fused kernel,
    from the committed ncu capture profiles/r2_fused_traffic.json (tools/traffic_record.py).
    The record carries the launch shape and a hash of the kernel sources: it is only reported
    for the same shape, and flagged when the sources changed since the capture."""
    try:
        rec = json.load(open(TRAFFIC_RECORD))
    except Exception:
        return None, "no capture (profiles/r2_fused_traffic.json missing)"
    if list(rec.get("shape", [])) != [nO, nT, nz, nB]:
        return None, "capture is for shape %s" % rec.get("shape")
    note = "ncu capture %s: read %.2f GB + written %.2f GB" % (
        rec.get("when", "?"), rec["dram_bytes_read"] / 1e9, rec["dram_bytes_write"] / 1e9)
    if rec.get("kernel_source_hash") != kernel_source_hash():
        note += " (kernel sources changed since the capture)"
    return rec["dram_bytes_read"] + rec["dram_bytes_write"], note


def workload_name(nO, nT, nz, nB):
    return ("C3 approx_flux_likelihood fused PDF+top-%d: %d targets x %d training GPs x %d z "
            "x %d bands (LSST ugrizy), f64" % (N_COMPRESS, nO, nT, nz, nB))


def shared_config(nO, nT, nz, nB):
    """The `config` of the JSON line: identical (keys and values) for both arms."""
    return {"workload": workload_name(nO, nT, nz, nB), "targets_per_gpu": nO, "n_train": nT,
            "nz": nz, "bands": nB, "Ncompress": N_COMPRESS,
            "l2": "inputs larger than L2: model grid 2 x nz x n_train x bands x 8 B = %.2f GB, "
                  "re-streamed every step" % (2.0 * nz * nT * nB * 8 / 1e9),
            "parallelism": "targets sharded by rank, no data-path collective"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path, on the host cores
# --------------------------------------------------------------------------------------------
def reference_step(cy_like, model_mean, model_covar, prior, zgrid, cat, first, count):
    """Loop body of scripts/delight-apply.py:114-162 for `count` targets (all bands valid)."""
    nz, nT, nB = model_mean.shape
    bands = np.arange(nB)
    out = []
    for o in range(first, first + count):
        ell_hat_z = np.ones(nz)
        like_grid = np.zeros((nz, nT))
        cy_like(like_grid, nz, nT, bands.size, cat["fluxes"][o], cat["fluxes_var"][o],
                model_mean[:, :, bands], model_covar[:, :, bands],
                ell_hat=ell_hat_z, ell_var=(ell_hat_z * np.sqrt(ELL_VAR)) ** 2)
        like_grid *= prior
        pdf = like_grid.sum(axis=1)
        evidences = np.trapz(like_grid, x=zgrid, axis=0)
        sortind = np.argsort(evidences)[::-1][0:N_COMPRESS]
        out.append((pdf, sortind, evidences[sortind]))
    return out


def set_omp_threads(n):
    """Thread count of the OpenMP runtime the reference's Cython modules use (libgomp)."""
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def reference_model_grid(nT, data):
    """Mock training catalogue -> the reference's own learn and predictAndInterpolate
    (scripts/delight-learn.py:61-90, scripts/delight-apply.py:84-98) on the host."""
    from delight_b200 import workloads
    from oracle import ref_pipeline
    nB = data["tmpl"].shape[1]
    cat = workloads.mock_catalogue(nT, data, seed=1, ell=1e6)
    set_omp_threads(1)      # the GP object's tiny prange loops are 30x slower with many threads
    gp, f_mod = ref_pipeline.make_gp(data, workloads.HYPER)
    bands = np.tile(np.arange(nB), (nT, 1))
    rows = ref_pipeline.learn(gp, f_mod, cat["z"], bands, np.full(nT, nB), cat["fluxes"],
                              cat["fluxes_var"])
    mm, mc, zt = ref_pipeline.model_grid(gp, rows, data["zgrid"], data["tmpl"].shape[0])
    set_omp_threads(os.cpu_count() or 1)
    return mm, mc, zt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from delight_b200 import workloads
    from oracle import ref, oracle as orc
    data = workloads.load_lsst()
    nz, nB = data["zgrid"].size, data["tmpl"].shape[1]
    cores = os.cpu_count() or 1
    if ref.available():
        ref.load()
        from delight.utils_cy import approx_flux_likelihood_cy as cy_like
        kind = "reference"
    else:
        cy_like, kind = orc.approx_flux_likelihood_cy, "port"
    nT = args.train
    t_setup = time.perf_counter()
    model_mean, model_covar, zt = reference_model_grid(nT, data)
    t_setup = time.perf_counter() - t_setup
    zgrid = data["zgrid"]
    prior = orc.z_prior(zgrid, zt, Z_PRIOR_SIGMA)
    per_step = args.ref_targets
    cat = workloads.mock_catalogue(per_step * (args.steps + args.warmup), data, seed=2)
    for w in range(args.warmup):
        reference_step(cy_like, model_mean, model_covar, prior, zgrid, cat, w * per_step, 1)
    t0 = time.perf_counter()
    for s in range(args.steps):
        reference_step(cy_like, model_mean, model_covar, prior, zgrid, cat,
                       (args.warmup + s) * per_step, per_step)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "galaxy-PDFs/sec", "value": value, "unit": "galaxies/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.targets, nT, nz, nB),
        "arm_notes": {"model_grid": "reference learn + predictAndInterpolate on %d mock training "
                                    "galaxies, %.0f s on the host (untimed setup)" % (nT, t_setup),
                      "sample": "%d targets per step" % per_step},
        "cpu_baseline": {"value": value, "unit": "galaxies/s", "cores": cores, "kind": kind,
                         "sample": "%d targets/step against the full %d x %d x %d grid, loop "
                                   "body of delight-apply.py:114-162 incl. its per-target "
                                   "fancy-index copies of the grid; OpenMP over %d threads"
                                   % (per_step, nz, nT, nB, cores)},
        "e2e": {"value": value, "unit": "galaxies/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def cpu_baseline_sample(data, model_mean, model_covar, zt, budget_s=20.0):
    """Reference loop body timed on the host cores of this box on a bounded sample, against the
    same model grid the GPU arm uses (downloaded in the script's [nz, nT, nB] layout)."""
    from delight_b200 import workloads
    from oracle import ref, oracle as orc
    if ref.available():
        ref.load()
        from delight.utils_cy import approx_flux_likelihood_cy as cy_like
        kind = "reference"
    else:
        cy_like, kind = orc.approx_flux_likelihood_cy, "port"
    zgrid = data["zgrid"]
    nz, nT, nB = model_mean.shape
    prior = orc.z_prior(zgrid, zt, Z_PRIOR_SIGMA)
    cat = workloads.mock_catalogue(64, data, seed=2)
    set_omp_threads(os.cpu_count() or 1)
    reference_step(cy_like, model_mean, model_covar, prior, zgrid, cat, 0, 1)     # warm
    n, t0, outs = 0, time.perf_counter(), []
    while n < 64 and (time.perf_counter() - t0 < budget_s or n < 3):
        outs += reference_step(cy_like, model_mean, model_covar, prior, zgrid, cat, n, 1)
        n += 1
    dt = time.perf_counter() - t0
    base = {"value": n / dt, "unit": "galaxies/s", "cores": os.cpu_count() or 1, "kind": kind,
            "sample": "%d targets against the full %d x %d x %d grid (loop body of "
                      "delight-apply.py:114-162 incl. its per-target fancy-index grid copies), "
                      "%.1f s" % (n, nz, nT, nB, dt)}
    return base, cat, outs


def parity_check(gpu, outs):
    """The GPU path vs what the CPU reference just produced for the same targets at the full
    bench shape: PDFs (relative, above 1e-290 x max) and top-K indices where the evidence is
    representable (the reference's argsort is unstable on exact ties)."""
    pdf_ref = np.stack([o[0] for o in outs])
    idx_ref = np.stack([o[1] for o in outs])
    ev_ref = np.stack([o[2] for o in outs])
    n = pdf_ref.shape[0]
    pdf = gpu["pdf"][:n]
    scale = pdf_ref.max(axis=1, keepdims=True)
    big = pdf_ref > 1e-290 * scale
    rel = np.abs(pdf - pdf_ref)[big] / pdf_ref[big]
    solid = ev_ref > 1e-250 * ev_ref[:, :1]
    return {"targets": int(n), "pdf_max_rel_err": float(rel.max()),
            "evid_max_rel_err": float((np.abs(gpu["top_ev"][:n] - ev_ref)[solid]
                                       / ev_ref[solid]).max()),
            "topk_indices_equal": bool(np.array_equal(gpu["top_idx"][:n][solid],
                                                      idx_ref[solid])),
            "tolerance": 1e-9}


def _ev_ms(fn, reps=1):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def config2_templatefit(data, budget_s=6.0):
    """BASELINE.json configs[1]: templateFitting.py-equivalent, 100k targets x 8 CWW templates x
    1500 z x 5 bands (ugriz).  Resident and end-to-end (pinned-host catalogue in, 1.2 GB of PDFs
    + metrics out) timings, the reference's own loop body (NumPy twin, p(z,t) prior, sum,
    computeMetrics; templateFitting.py:60-89) timed on a sample of the same targets, and the
    GPU-vs-reference error on that sample."""
    import torch
    from delight_b200 import apply
    zgrid, tmpl = data["zgrid"], data["tmpl"][:, :5]
    nt, nB, nz = tmpl.shape
    f_mod = np.ascontiguousarray(tmpl.transpose(2, 0, 1))
    grid = apply.ModelGrid.from_legacy(f_mod, np.zeros_like(f_mod))
    prior = grid.template_prior(zgrid, data["p_t"], data["p_z_t"])      # templateFitting.py:79-83
    rng = np.random.default_rng(3)
    nO = 100000
    ztrue = zgrid[rng.integers(0, nz, nO)]
    fl = f_mod[np.searchsorted(zgrid, ztrue), rng.integers(0, nt, nO)] * rng.uniform(.5, 2, (nO, 1))
    fl = fl * (1 + 0.05 * rng.standard_normal(fl.shape))
    fv = (0.05 * fl) ** 2
    ws = apply.FusedWorkspace()
    CL = [0.1, 0.5, 0.68, 0.95]
    tg = apply.Targets.full(fl, fv, ztrue).to_device()
    run = lambda t=tg: apply.apply_fused(grid, prior, zgrid, t, ell_hat=1.0, ell_var=1e24, K=0,
                                         want_metrics=True, confidence_levels=CL, workspace=ws,
                                         out_device=True)
    run()
    ms, res = _ev_ms(run, 2)
    # end to end: host catalogue in, PDFs + metrics out
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_fl, h_fv = pin(fl), pin(fv)
    h_pdf = torch.empty((nO, nz), dtype=torch.float64).pin_memory()
    h_met = torch.empty((nO, 8 + len(CL)), dtype=torch.float64).pin_memory()

    def e2e():
        # host catalogue in, PDFs + metrics into pinned host arrays; the copies of one chunk of
        # targets overlap the kernels of the next (apply.apply_fused_to_host)
        t = apply.Targets.full(h_fl.numpy(), h_fv.numpy(), ztrue)
        apply.apply_fused_to_host(grid, prior, zgrid, t, {"pdf": h_pdf, "metrics": h_met},
                                  ell_hat=1.0, ell_var=1e24, K=0, want_metrics=True,
                                  confidence_levels=CL, workspace=ws)
    e2e()
    t0 = time.perf_counter()
    e2e()
    e2e_s = time.perf_counter() - t0
    out = {"workload": "C2 templateFitting-equivalent: 100k targets x 8 CWW templates x 1500 z x "
                       "5 bands, PDF + computeMetrics (4 confidence levels)",
           "galaxies_per_s": nO / (ms * 1e-3), "ms": ms,
           "e2e_galaxies_per_s": nO / e2e_s, "e2e_ms": e2e_s * 1e3,
           "h2d_bytes": h_fl.numel() * 16, "d2h_bytes": h_pdf.numel() * 8 + h_met.numel() * 8}
    # the reference's loop body on the host, same targets
    try:
        from oracle import ref
        if ref.available():
            ref.load()
            from delight.utils import approx_flux_likelihood, computeMetrics
            b2 = data["p_z_t"] ** 2
            p_z = data["p_t"][None, :] * zgrid[:, None] / b2[None, :] * np.exp(
                -0.5 * zgrid[:, None] ** 2 / b2[None, :])
            pdfs, mets, n, t0 = [], [], 0, time.perf_counter()
            while n < 256 and (time.perf_counter() - t0 < budget_s or n < 4):
                like = approx_flux_likelihood(fl[n], fv[n], f_mod, normalized=True,
                                              marginalizeEll=True, ell_hat=1,
                                              ell_var=(1 * 1e12) ** 2)
                like *= p_z
                pdfs.append(like.sum(axis=1))
                mets.append(computeMetrics(ztrue[n], zgrid, pdfs[-1], CL))
                n += 1
            dt = time.perf_counter() - t0
            pr, mr = np.stack(pdfs), np.stack(mets)
            g_pdf, g_met = res["pdf"][:n].cpu().numpy(), res["metrics"][:n, :mr.shape[1]].cpu().numpy()
            big = pr > 1e-290 * pr.max(axis=1, keepdims=True)
            out["cpu_baseline"] = {"value": n / dt, "unit": "galaxies/s", "cores": 1,
                                   "kind": "reference",
                                   "sample": "%d targets, loop body of templateFitting.py:60-89 "
                                             "(NumPy, single thread), %.1f s" % (n, dt)}
            out["parity_check"] = {"targets": n,
                                   "pdf_max_rel_err": float((np.abs(g_pdf - pr)[big] / pr[big]).max()),
                                   "metrics_max_rel_err": float(np.abs(g_met / mr - 1).max()),
                                   "tolerance": 1e-9}
    except Exception as e:                                   # the checker must not sink the bench
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def config4_learn(data, full=False):
    """BASELINE.json configs[3]: delight-learn on 100k synthetic training galaxies, 28 bands with
    2..21 valid per galaxy (G10-style): kernel grids for 28 bands + batched template fit, K(X),
    Cholesky, gpparams row.  full=True also times the reference (PhotozGP constructor = its
    kernel-grid build, and its learn loop on a sample) and compares the rows."""
    import torch
    from delight_b200 import gp as gpmod
    rng = np.random.default_rng(4)
    zgrid = data["zgrid"]
    nz = zgrid.size
    nB, NC, nT = 28, 7, 100000
    mu0 = np.linspace(3500, 22000, nB)
    amp = rng.uniform(0.2, 1.0, (nB, NC))
    mu = mu0[:, None] + np.linspace(-300, 300, NC)[None, :]
    sig = rng.uniform(80, 200, (nB, NC))
    zgp = data["zgridGP"]
    tm = 1e-6 * rng.uniform(0.5, 2.0, (8, nB, 1)) * (1 + 0.5 * np.sin(
        zgrid[None, None, :] * rng.uniform(1, 5, (8, nB, 1))))
    lines = [6500.0, 5002.26, 3732.22]
    mk = lambda: gpmod.PhotozGPBatched(tm, zgrid, amp, mu, sig, lines, 0.5, 0.5, 1e3, 1e2, zgp)
    mk()
    grid_ms, gpm = _ev_ms(mk)
    z = rng.uniform(0.05, 2.9, nT)
    nb = rng.integers(2, 22, nT)
    order = np.argsort(rng.random((nT, nB)), axis=1)
    keep = np.arange(nB)[None, :] < nb[:, None]
    bands = np.sort(np.where(keep, order, nB + 1), axis=1)
    bands = np.where(bands > nB, 0, bands)
    ty = rng.integers(0, 8, nT)
    iz = np.clip(((z - zgrid[0]) / (zgrid[1] - zgrid[0])).astype(int), 0, nz - 1)
    true = 1e6 * tm[ty[:, None], bands, iz[:, None]]
    fl = np.where(keep, true * (1 + 0.05 * rng.standard_normal(true.shape)), 0.0)
    fv = np.where(keep, (0.05 * true) ** 2, 0.0)
    cat = gpmod.catalogue(fl, fv, bands, nb, z)
    gpm.fit(cat, out_device=True)
    t0 = time.perf_counter(); rows = gpm.fit(cat, out_device=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"workload": "C4 delight-learn: 100k training galaxies, 28 bands (2..21 valid each), "
                       "template fit + K(X) + Cholesky + gpparams row; incl. H2D of the catalogue",
           "galaxies_per_s": nT / dt, "ms": dt * 1e3, "kernel_grid_28_bands_ms": grid_ms,
           "kernel_grid_pairs": nB * nB * zgp.size * zgp.size,
           "finite": bool(torch.isfinite(rows).all().item())}
    if full:
        from oracle import ref, ref_pipeline
        if ref.available():
            d28 = dict(tmpl=tm, zgrid=zgrid, amp=amp, mu=mu, sig=sig, zgridGP=zgp)
            hyper = dict(lines_pos=lines, lines_width=[20.0] * 3, V_C=0.5, V_L=0.5, alpha_C=1e3,
                         alpha_L=1e2)
            set_omp_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            gp_ref, f_mod = ref_pipeline.make_gp(d28, hyper)
            ctor_s = time.perf_counter() - t0
            set_omp_threads(1)
            n = 200
            t0 = time.perf_counter()
            rr = ref_pipeline.learn(gp_ref, f_mod, z[:n], bands[:n], nb[:n], fl[:n], fv[:n])
            ls = time.perf_counter() - t0
            got = rows[:n].cpu().numpy()
            m = rr != 0
            out["cpu_baseline"] = {"value": n / ls, "unit": "galaxies/s", "cores": 1,
                                   "kind": "reference",
                                   "sample": "%d galaxies through the reference's learn loop "
                                             "(delight-learn.py:61-90), %.1f s; its PhotozGP "
                                             "constructor (28-band kernel grids, OpenMP over %d "
                                             "threads): %.1f s" % (n, ls, os.cpu_count() or 1,
                                                                   ctor_s),
                                   "kernel_grid_s": ctor_s}
            out["parity_check"] = {"galaxies": n,
                                   "rows_max_rel_err": float(np.abs(got[m] / rr[m] - 1).max()),
                                   "band_lists_equal": bool(np.array_equal(got[:, :3 + nB] != 0,
                                                                           rr[:, :3 + nB] != 0)),
                                   "tolerance": 1e-9}
    return out


def config5_strong(data, grid_cache, world, rank, total_targets, nT=100000):
    """BASELINE.json configs[4] scaled to a bench-sized slice: `total_targets` targets (STRONG
    scaling: split over the ranks with the reference's formula) against 100k training GPs x 1500
    z x 6 bands, model grid (14.7 GB) replicated per GPU.  Max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    from delight_b200 import apply, gp as gpmod, workloads
    H = workloads.HYPER
    zgrid = data["zgrid"]
    nB = data["tmpl"].shape[1]
    gpm = grid_cache["gpm"]
    tr = workloads.mock_catalogue(nT, data, seed=11, ell=1e6)
    tcat = gpmod.catalogue(tr["fluxes"], tr["fluxes_var"], np.tile(np.arange(nB), (nT, 1)),
                           np.full(nT, nB), tr["z"])
    t0 = time.perf_counter()
    grid, train_z = gpm.predict(gpm.fit(tcat, out_device=True))
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    prior = grid.gaussian_prior(zgrid, train_z, Z_PRIOR_SIGMA)
    first, last = apply.partition(total_targets, rank, world)
    cat = workloads.mock_catalogue(total_targets, data, seed=12)
    tg = apply.Targets.full(cat["fluxes"][first:last], cat["fluxes_var"][first:last]).to_device()
    ws = apply.FusedWorkspace()
    run = lambda: apply.apply_fused(grid, prior, zgrid, tg, ell_hat=1.0, ell_var=ELL_VAR,
                                    K=N_COMPRESS, workspace=ws, out_device=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, res = _ev_ms(run)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    del grid, prior
    return {"workload": "C5-shaped slice, strong scaling: %d targets in all x %d training GPs x "
                        "%d z x %d bands; model grid %.1f GB replicated per GPU"
                        % (total_targets, nT, zgrid.size, nB,
                           nT * (2 * nB + 1) * 1536 * 8 / 1e9),
            "galaxies_per_s": total_targets / (ms * 1e-3), "ms": ms, "n_gpus": world,
            "scaling": "strong", "fused_launches_per_rank": res["fused_launches"],
            "model_grid_setup_s": setup_s}


def config1_stock_scripts(n=300):
    """BASELINE.json configs[0]: the reference's own delight-learn.py / delight-apply.py
    (byte-compiled, unmodified) on an LSST-band mock of n + n objects -- once on the reference's
    Cython operators, once with delight_b200's CUDA operators registered under the reference's
    module names (drop-in mode).  Wall-clock per script; output files compared."""
    import tempfile
    from delight_b200 import workloads
    from oracle import pipeline as refpipe, ref
    if not ref.available():
        return {"unavailable": "oracle/_ref not built"}
    data = workloads.load_lsst()
    out = {"workload": "C1 stock scripts: %d training + %d target galaxies, 6 LSST bands, "
                       "1500 z, nzGP 150" % (n, n)}
    files = {}
    for mode in ("reference", "dropin"):
        with tempfile.TemporaryDirectory(prefix="dl_c1_") as wd:
            cfg = refpipe.make_workdir(wd, bands=refpipe.LSST, refband="lsst_i", zmin=0.01,
                                       zmax=3.01, dz=0.002, nzGP=150, numObjects=n,
                                       Ncompress=10, copy_data=False)
            refpipe.write_preprocessed(wd, refpipe.LSST, data["amp"], data["mu"], data["sig"],
                                       data["tmpl"])
            for name, seed, ell in (("galaxies-fluxredshifts.txt", 21, 1e6),
                                    ("galaxies-fluxredshifts2.txt", 22, 1.0)):
                c = workloads.mock_catalogue(n, data, seed=seed, ell=ell, extra_frac=0.0)
                np.savetxt(os.path.join(wd, name),
                           np.column_stack([c["fluxes"], c["fluxes_var"], c["z"], c["types"]]))
            t0 = time.perf_counter()
            refpipe.run_script("delight-learn", cfg, omp_threads=1, dropin=mode == "dropin")
            t1 = time.perf_counter()
            refpipe.run_script("delight-apply", cfg, dropin=mode == "dropin")
            t2 = time.perf_counter()
            out[mode] = {"learn_s": t1 - t0, "apply_s": t2 - t1}
            files[mode] = (np.loadtxt(os.path.join(wd, "galaxies-gpparams.txt")),
                           np.loadtxt(os.path.join(wd, "redshiftpdfs.txt")))
    a, b = files["reference"], files["dropin"]
    m = a[0] != 0
    out["parity_check"] = {"gpparams_max_rel_err": float(np.abs(b[0][m] / a[0][m] - 1).max()),
                           "pdf_files_max_rel_err (%.2e text)": float(
                               (np.abs(b[1] - a[1]) / np.maximum(np.abs(a[1]), 1e-300))[a[1] > 0].max())}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from delight_b200 import _lib, apply, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()

    data = workloads.load_lsst()
    zgrid = data["zgrid"]
    nz, nB = zgrid.size, data["tmpl"].shape[1]
    nT, nO = args.train, args.targets
    H = workloads.HYPER
    # model grid, replicated on every GPU: mock training catalogue -> batched GP fit (learn)
    # -> batched GP predict, all on the device
    from delight_b200 import gp as gpmod
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    gpm = gpmod.PhotozGPBatched(data["tmpl"], zgrid, data["amp"], data["mu"], data["sig"],
                                H["lines_pos"], H["V_C"], H["V_L"], H["alpha_C"], H["alpha_L"],
                                data["zgridGP"])
    ev[1].record()
    tr = workloads.mock_catalogue(nT, data, seed=1, ell=1e6)
    tcat = gpmod.catalogue(tr["fluxes"], tr["fluxes_var"], np.tile(np.arange(nB), (nT, 1)),
                           np.full(nT, nB), tr["z"])
    rows = gpm.fit(tcat, out_device=True)
    ev[2].record()
    grid, train_z = gpm.predict(rows)
    ev[3].record()
    torch.cuda.synchronize()
    setup_ms = {"kernel_grids_ms": ev[0].elapsed_time(ev[1]),
                "learn_ms_incl_h2d": ev[1].elapsed_time(ev[2]),
                "predict_ms": ev[2].elapsed_time(ev[3])}
    prior = grid.gaussian_prior(zgrid, train_z, Z_PRIOR_SIGMA)
    cat = workloads.mock_catalogue(nO, data, seed=100 + rank)                # this rank's shard
    targets = apply.Targets.full(cat["fluxes"], cat["fluxes_var"], cat["z"])
    ws = apply.FusedWorkspace()
    zg_dev = torch.from_numpy(zgrid).cuda()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(tg):
        return apply.apply_fused(grid, prior, zg_dev, tg, ell_hat=1.0, ell_var=ELL_VAR,
                                 K=N_COMPRESS, workspace=ws, out_device=True)

    # ---- kernel-resident throughput ------------------------------------------------------
    resident = apply.Targets.full(cat["fluxes"], cat["fluxes_var"], cat["z"]).to_device()
    for _ in range(args.warmup):
        step_resident(resident)
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    L.dl_timing_enable(1)
    launches0 = L.dl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident(resident)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    kms, kn = ctypes.c_double(0), ctypes.c_int(0)
    L.dl_timing_collect(ctypes.byref(kms), ctypes.byref(kn))
    L.dl_timing_enable(0)
    launches = L.dl_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * nO * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers --------------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_flux, h_var = pin(cat["fluxes"]), pin(cat["fluxes_var"])
    h_pdf = torch.empty((nO, nz), dtype=torch.float64).pin_memory()
    h_idx = torch.empty((nO, N_COMPRESS), dtype=torch.int64).pin_memory()
    h_ev = torch.empty((nO, N_COMPRESS), dtype=torch.float64).pin_memory()

    def step_e2e():
        tg = apply.Targets.full(h_flux.numpy(), h_var.numpy())
        tg._dev = {("nf", nB): (h_flux.cuda(non_blocking=True),
                                h_var.cuda(non_blocking=True), None, None)}
        out = step_resident(tg)
        if world > 1:          # gather to rank 0 like the (commented) Gatherv calls, :201-217
            bufs = [torch.empty_like(out["pdf"]) for _ in range(world)] if rank == 0 else None
            dist.gather(out["pdf"], bufs, dst=0)
            ib = [torch.empty_like(out["top_idx"]) for _ in range(world)] if rank == 0 else None
            dist.gather(out["top_idx"], ib, dst=0)
            eb = [torch.empty_like(out["top_ev"]) for _ in range(world)] if rank == 0 else None
            dist.gather(out["top_ev"], eb, dst=0)
        h_pdf.copy_(out["pdf"], non_blocking=True)
        h_idx.copy_(out["top_idx"], non_blocking=True)
        h_ev.copy_(out["top_ev"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * nO * args.steps / (float(ms2.item()) * 1e-3)
    h2d = h_flux.numel() * 8 + h_var.numel() * 8
    d2h = h_pdf.numel() * 8 + h_idx.numel() * 8 + h_ev.numel() * 8

    # ---- secondary configurations (BASELINE.json configs[0], [1], [3], [4]) ----------------------
    extras = None
    if not args.no_extras:
        extras = {}
        if args.c5_targets > 0:          # every rank takes part (strong scaling over the ranks)
            del resident
            ws.buf = None
            torch.cuda.empty_cache()
            c5 = config5_strong(data, {"gpm": gpm}, world, rank, args.c5_targets)
            if rank == 0:
                extras["c5_strong"] = c5
        if rank == 0 and world == 1:
            extras["c2_templatefit"] = config2_templatefit(data)
            extras["c4_learn"] = config4_learn(data, full=args.extras == "full")
            if args.extras == "full":
                extras["c1_stock_scripts"] = config1_stock_scripts()
    if rank == 0:
        peaks = measured_peaks()
        fp64 = ctypes.c_double(0)
        L.dl_fp64_peak(5, ctypes.byref(fp64))
        cells_per_launch = float(nO) * nT * nz * args.steps / max(kn.value, 1)
        k_ms = kms.value / max(kn.value, 1)
        achieved = cells_per_launch * FLOP_PER_CELL(nB) / (k_ms * 1e-3) * 1e-12
        grid_bytes = grid.grid.numel() * 8 + prior.numel() * 8
        traffic, traffic_note = measured_traffic(nO, nT, nz, nB)
        roofline = {
            "bound": "fp64", "kernel": "fused_like_kernel<%d,false,false,MAIN>" % nB,
            "bound_note": "FP64 vector (FMA) pipe: no stage is a dense contraction (tensor cores "
                          "not applicable) and HBM traffic is 3 orders of magnitude below its "
                          "roofline (BASELINE.json metric: % of FP64/HBM roofline)",
            "achieved": achieved, "peak": fp64.value, "unit": "TFLOP/s",
            "frac": achieved / fp64.value if fp64.value else None,
            "peak_source": "measured live: dl_fp64_peak (dependent FMA chains, all SMs); "
                           "MEASURED_PEAKS.json has no FP64 entry; datasheet 37.2 TFLOP/s "
                           "(profiles/r2_fp64_peak.json)",
            "flop_per_cell": FLOP_PER_CELL(nB), "cells_per_launch": cells_per_launch,
            "kernel_ms": k_ms, "kernel_share_of_step": kms.value / ms_total,
            "traffic": traffic, "traffic_source": traffic_note,
            "traffic_unit": "bytes per launch (ncu dram read + write of the main pass)",
            "hbm_algorithmic_bytes_per_launch": grid_bytes + nO * (2 * nB * 8 + nz * 8
                                                                  + N_COMPRESS * 16 + 64),
            "hbm_peak_GBs": peaks.get("hbm_gbs"),
        }
        cpu, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            G = grid.grid[:, :, :nz].cpu().numpy()                 # [nT, 2nB, nz]
            mm = np.ascontiguousarray(G[:, :nB, :].transpose(2, 0, 1))
            mc = np.ascontiguousarray(G[:, nB:, :].transpose(2, 0, 1))
            del G
            cpu, pcat, pouts = cpu_baseline_sample(data, mm, mc, train_z.cpu().numpy())
            n = len(pouts)
            g = apply.apply_fused(grid, prior, zg_dev,
                                  apply.Targets.full(pcat["fluxes"][:n], pcat["fluxes_var"][:n]),
                                  ell_hat=1.0, ell_var=ELL_VAR, K=N_COMPRESS, workspace=ws)
            parity = parity_check(g, pouts)
        line = {
            "metric": "galaxy-PDFs/sec", "value": value, "unit": "galaxies/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": shared_config(nO, nT, nz, nB),
            "arm_notes": {"model_grid": "device learn + predict of %d simulateWithSEDs-style "
                                        "mock training galaxies, replicated per GPU, %.2f GB in "
                                        "the kernel's layout" % (nT, grid_bytes / 1e9),
                          "setup": setup_ms},
            "e2e": {"value": e2e_value, "unit": "galaxies/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_check": parity,
            "other_configs": extras,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--targets", type=int, default=N_TARGETS, help="targets per GPU per step")
    ap.add_argument("--train", type=int, default=N_TRAIN)
    ap.add_argument("--ref-targets", type=int, default=8,
                    help="reference arm: targets per step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--extras", default="default", choices=["default", "full"],
                    help="full: also the reference baselines of config 4 and the stock scripts "
                         "of config 1 (minutes of host time)")
    ap.add_argument("--c5-targets", type=int, default=16384,
                    help="targets IN ALL of the config-5-shaped strong-scaling extra (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

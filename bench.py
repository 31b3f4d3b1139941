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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_like_kernel launch at the default
# configuration, from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` on a B200
# (profiles/r1_fused_like_dram_traffic.csv, grid 3756 x 4): 4.24 GB read (1.6 GB grid + prior,
# re-read a few times over the launch, L2 hit rate 94 %) + 19.78 GB written (per-z-chunk evidence
# partials, 12 x 16 B per (target, training), + the PDF partials).
NCU_TRAFFIC_BYTES_DEFAULT_CONFIG = 4237682176 + 19782158336


def workload_name(nO, nT, nz, nB):
    return ("C3 approx_flux_likelihood fused PDF+top-%d: %d targets x %d training GPs x %d z "
            "x %d bands (LSST ugrizy), f64" % (N_COMPRESS, nO, nT, nz, nB))


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
        "config": {"workload": workload_name(N_TARGETS, nT, nz, nB),
                   "inputs": "larger than L2 (model grid 2 x %.0f MB)" % (model_mean.nbytes / 1e6),
                   "model_grid": "reference learn + predictAndInterpolate on %d mock training "
                                 "galaxies, %.0f s on the host (untimed setup)" % (nT, t_setup)},
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


def other_configs(data):
    """Secondary BASELINE.json configurations, timed once each (CUDA events, inputs resident):
    config 2 (templateFitting-equivalent, 100k targets x 8 templates x 1500 z x 5 bands) and
    config 4 (delight-learn on 100k training galaxies, 28 bands with 2..21 valid per galaxy)."""
    import torch
    from delight_b200 import apply, gp as gpmod
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- config 2
    zgrid, tmpl = data["zgrid"], data["tmpl"][:, :5]
    nt, nB, nz = tmpl.shape
    f_mod = np.ascontiguousarray(tmpl.transpose(2, 0, 1))
    b2 = data["p_z_t"] ** 2
    p_z = data["p_t"][None, :] * zgrid[:, None] / b2[None, :] * np.exp(
        -0.5 * zgrid[:, None] ** 2 / b2[None, :])                    # templateFitting.py:79-82
    grid = apply.ModelGrid.from_legacy(f_mod, np.zeros_like(f_mod))
    prior = grid.explicit_prior(p_z)
    rng = np.random.default_rng(3)
    nO = 100000
    fl = f_mod[rng.integers(0, nz, nO), rng.integers(0, nt, nO)] * rng.uniform(.5, 2, (nO, 1))
    fl = fl * (1 + 0.05 * rng.standard_normal(fl.shape))
    tg = apply.Targets.full(fl, (0.05 * fl) ** 2).to_device()
    ws = apply.FusedWorkspace()
    run = lambda: apply.apply_fused(grid, prior, zgrid, tg, ell_hat=1.0, ell_var=1e24, K=0,
                                    want_metrics=True, workspace=ws, out_device=True)
    run()
    ev0.record(); run(); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out["c2_templatefit"] = {"workload": "100k targets x 8 CWW templates x 1500 z x 5 bands, "
                             "PDF + metrics", "galaxies_per_s": nO / (ms * 1e-3), "ms": ms}
    del grid, prior, tg
    # ---- config 4
    rng = np.random.default_rng(4)
    nB, NC, nzg, nT = 28, 7, 150, 100000
    mu0 = np.linspace(3500, 22000, nB)
    amp = rng.uniform(0.2, 1.0, (nB, NC))
    mu = mu0[:, None] + np.linspace(-300, 300, NC)[None, :]
    sig = rng.uniform(80, 200, (nB, NC))
    zgp = data["zgridGP"]
    tm = 1e-6 * rng.uniform(0.5, 2.0, (8, nB, 1)) * (1 + 0.5 * np.sin(
        zgrid[None, None, :] * rng.uniform(1, 5, (8, nB, 1))))
    ev0.record()
    gpm = gpmod.PhotozGPBatched(tm, zgrid, amp, mu, sig, [6500.0, 5002.26, 3732.22], 0.5, 0.5,
                                1e3, 1e2, zgp)
    ev1.record(); torch.cuda.synchronize()
    grid_ms = ev0.elapsed_time(ev1)
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
    ok = bool(torch.isfinite(rows).all().item())
    out["c4_learn"] = {"workload": "100k training galaxies, 28 bands (2..21 valid each), "
                       "template fit + K(X) + Cholesky + gpparams row; incl. H2D of the "
                       "catalogue", "galaxies_per_s": nT / dt, "ms": dt * 1e3,
                       "kernel_grid_28_bands_ms": grid_ms, "kernel_grid_pairs": nB * nB * nzg * nzg,
                       "finite": ok}
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

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = other_configs(data)
    if rank == 0:
        peaks = measured_peaks()
        fp64 = ctypes.c_double(0)
        L.dl_fp64_peak(5, ctypes.byref(fp64))
        cells_per_launch = float(nO) * nT * nz * args.steps / max(kn.value, 1)
        k_ms = kms.value / max(kn.value, 1)
        achieved = cells_per_launch * FLOP_PER_CELL(nB) / (k_ms * 1e-3) * 1e-12
        grid_bytes = grid.grid.numel() * 8 + prior.numel() * 8
        roofline = {
            "bound": "fp64", "kernel": "fused_like_kernel<%d,false,2>" % nB,
            "bound_note": "FP64 vector (FMA) pipe: no stage is a dense contraction (tensor cores "
                          "not applicable) and HBM traffic is 3 orders of magnitude below its "
                          "roofline (BASELINE.json metric: % of FP64/HBM roofline)",
            "achieved": achieved, "peak": fp64.value, "unit": "TFLOP/s",
            "frac": achieved / fp64.value if fp64.value else None,
            "peak_source": "measured live: dl_fp64_peak (dependent FMA chains, all SMs); "
                           "MEASURED_PEAKS.json has no FP64 entry; datasheet 37.2 TFLOP/s",
            "flop_per_cell": FLOP_PER_CELL(nB), "cells_per_launch": cells_per_launch,
            "kernel_ms": k_ms, "kernel_share_of_step": kms.value / ms_total,
            "traffic": (NCU_TRAFFIC_BYTES_DEFAULT_CONFIG
                        if (nO, nT) == (N_TARGETS, N_TRAIN) else None),
            "traffic_unit": "bytes per launch (ncu dram read+write, profiles/)",
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
            "config": {"workload": workload_name(nO, nT, nz, nB),
                       "targets_per_gpu": nO, "n_train": nT, "nz": nz, "bands": nB,
                       "Ncompress": N_COMPRESS,
                       "model_grid": "device learn + predict of %d simulateWithSEDs-style mock "
                                     "training galaxies, replicated per GPU, %.2f GB"
                                     % (nT, grid_bytes / 1e9),
                       "setup": setup_ms,
                       "l2": "inputs larger than L2 (model grid %.2f GB re-streamed per step)"
                             % (grid_bytes / 1e9),
                       "parallelism": "targets sharded by rank, no data-path collective"},
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

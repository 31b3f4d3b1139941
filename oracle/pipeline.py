"""Harness that drives the UNMODIFIED reference pipeline -- TEST INFRASTRUCTURE ONLY.

Runs the byte-compiled reference scripts of oracle/_ref/scripts (processFilters -> processSEDs ->
simulateWithSEDs -> delight-learn -> delight-apply) in a scratch directory on a writable copy of
the reference's data files.  Needs /root/reference/data (filters, SEDs), i.e. it only works in
the build container -- it is used to GENERATE the committed fixtures of tests/golden/, never at
GPU-test / bench time.

The cfg written here is tests/parametersTest.cfg with the non-numeric fixes of SURVEY.md 8c:
`sed_fmt` added, every output given its own file name, crossValidate off.
"""
import os
import shutil
import subprocess
import sys

from . import ref

REPO = os.path.dirname(ref.HERE)

SDSS = ["U_SDSS", "G_SDSS", "R_SDSS", "I_SDSS", "Z_SDSS"]
LSST = ["lsst_u", "lsst_g", "lsst_r", "lsst_i", "lsst_z", "lsst_y"]
CWW = ["El_B2004a", "Sbc_B2004a", "Scd_B2004a", "SB3_B2004a", "Im_B2004a", "SB2_B2004a",
       "ssp_25Myr_z008", "ssp_5Myr_z008"]

CFG = """
[Bands]
Names: {bands}
directory: {wd}/FILTERS
numCoefs: {numCoefs}

[Templates]
directory: {wd}/CWW_SEDs
sed_fmt: dat
names: {seds}
p_t: 0.27 0.26 0.25 0.069 0.021 0.11 0.0061 0.0079
p_z_t:0.23 0.39 0.33 0.31 1.1 0.34 1.2 0.14
lambdaRef: 4.5e3

[Simulation]
numObjects: {numObjects}
noiseLevel: {noiseLevel}
trainingFile: {wd}/galaxies-fluxredshifts.txt
targetFile: {wd}/galaxies-fluxredshifts2.txt

[Training]
extraFracFluxError: 1e-5
catFile: {wd}/galaxies-fluxredshifts.txt
paramFile: {wd}/galaxies-gpparams.txt
referenceBand: {refband}
bandOrder: {train_order}
optimize: True
numChunks: 1
CVfile: {wd}/galaxies-gpCV.txt
crossValidate: False
crossValidationBandOrder: {cv_order}

[Target]
extraFracFluxError: 1e-5
catFile: {wd}/galaxies-fluxredshifts2.txt
referenceBand: {refband}
bandOrder: {target_order}
compressIndicesFile: {wd}/compressIndices.txt
compressMargLikFile: {wd}/compressLikes.txt
Ncompress: {Ncompress}
useCompression: True
redshiftpdfFile: {wd}/redshiftpdfs.txt
redshiftpdfFileComp: {wd}/redshiftpdfs-comp.txt
redshiftpdfFileTemp: {wd}/redshiftpdfs-cww.txt
metricsFile: {wd}/metrics.txt
metricsFileTemp: {wd}/metrics-cww.txt

[Other]
rootDir: {wd}
zPriorSigma: {zPriorSigma}
ellPriorSigma: {ellPriorSigma}
fluxLuminosityNorm: 2.2
alpha_C: 1e3
V_C: 0.5
alpha_L: 1.0e2
V_L: 0.5
lines_pos: 6500 5002.26 3732.22
lines_width: 20.0 20.0 20.0
redshiftMin: {zmin}
redshiftMax: {zmax}
redshiftNumBinsGPpred: {nzGP}
redshiftBinSize: {dz}
redshiftDisBinSize: 0.3
confidenceLevels: 0.1 0.50 0.68 0.95
"""


def make_workdir(wd, bands=SDSS, train_bands=None, refband=None, reference="/root/reference",
                 numObjects=100, noiseLevel=0.05, zmin=0.01, zmax=1.511, dz=0.002, nzGP=150,
                 Ncompress=50, zPriorSigma=0.5, ellPriorSigma=0.5, numCoefs=7, copy_data=True):
    """Copy the needed data files to `wd` (copy_data=False: the caller provides the
    pre-processed coefficient / flux-redshift files itself) and write params.cfg; returns the
    cfg path."""
    os.makedirs(wd, exist_ok=True)
    os.makedirs(os.path.join(wd, "FILTERS"), exist_ok=True)
    os.makedirs(os.path.join(wd, "CWW_SEDs"), exist_ok=True)
    for b in bands if copy_data else []:
        shutil.copy(os.path.join(reference, "data", "FILTERS", b + ".res"),
                    os.path.join(wd, "FILTERS"))
    for s in CWW if copy_data else []:
        shutil.copy(os.path.join(reference, "data", "CWW_SEDs", s + ".dat"),
                    os.path.join(wd, "CWW_SEDs"))
    train_bands = list(train_bands or bands)
    refband = refband or bands[len(bands) // 2 + 1 if len(bands) > 3 else 0]
    tr = [b if b in train_bands else "_" for b in bands]
    train_order = " ".join(tr + [b + "_var" if b in train_bands else "_" for b in bands]
                           + ["redshift"])
    cv = ["_" if b in train_bands else b for b in bands]
    cv_order = " ".join(cv + ["_" if b in train_bands else b + "_var" for b in bands] + ["_"])
    target_order = " ".join(list(bands) + [b + "_var" for b in bands] + ["redshift"])
    cfg = os.path.join(wd, "params.cfg")
    with open(cfg, "w") as f:
        f.write(CFG.format(wd=wd, bands=" ".join(bands), seds=" ".join(CWW),
                           numObjects=numObjects, noiseLevel=noiseLevel, refband=refband,
                           train_order=train_order, cv_order=cv_order,
                           target_order=target_order, Ncompress=Ncompress,
                           zPriorSigma=zPriorSigma, ellPriorSigma=ellPriorSigma,
                           zmin=zmin, zmax=zmax, nzGP=nzGP, dz=dz, numCoefs=numCoefs))
    return cfg


def write_preprocessed(wd, bands, amp, mu, sig, tmpl, seds=CWW):
    """Write <band>_gaussian_coefficients.txt and <sed>_fluxredshiftmod.txt (the outputs of
    processFilters.py:57-62 / processSEDs.py:51) from arrays, e.g. from a golden fixture."""
    import numpy as np
    for ib, b in enumerate(bands):
        np.savetxt(os.path.join(wd, "FILTERS", b + "_gaussian_coefficients.txt"),
                   np.stack([amp[ib], mu[ib], sig[ib]], axis=1), header=b)
    for it, s in enumerate(seds):
        np.savetxt(os.path.join(wd, "CWW_SEDs", s + "_fluxredshiftmod.txt"), tmpl[it].T)


_RUNNER = r"""
import sys, warnings
warnings.filterwarnings('ignore')
sys.path.insert(0, {repo!r})
{pre}
from oracle import ref
ref.load()
import numpy as np
np.random.seed({seed})
import runpy
sys.argv = [{script!r}, {cfg!r}]
runpy.run_path({script!r}, run_name='__main__')
"""


def run_script(name, cfg, seed=0, omp_threads=None, quiet=True, dropin=False):
    """Run one byte-compiled reference script on `cfg` with the global NumPy RNG seeded.
    dropin=True first aliases delight_b200's CUDA operator modules as delight.utils_cy /
    delight.photoz_kernels_cy (the script itself and the reference's Python layers are
    unchanged)."""
    env = dict(os.environ)
    if omp_threads:
        env["OMP_NUM_THREADS"] = str(omp_threads)
    pre = "import delight_b200.install; delight_b200.install.install()" if dropin else ""
    code = _RUNNER.format(repo=REPO, seed=seed, script=ref.script_path(name), cfg=cfg, pre=pre)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=os.path.dirname(cfg),
                         stdout=subprocess.PIPE if quiet else None,
                         stderr=subprocess.STDOUT if quiet else None, text=True)
    if out.returncode != 0:
        raise RuntimeError("reference script %s failed:\n%s" % (name, out.stdout))
    return out.stdout

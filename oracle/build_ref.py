#!/usr/bin/env python
"""Build the UNMODIFIED reference (ixkael/Delight) into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

Nothing in the product path (delight_b200/) may import anything from oracle/.

What this does (sources are read where they lie under /root/reference, nothing is copied
into the repository; only compiled artefacts are written, and only below oracle/_ref/):

  * delight/utils_cy.pyx, delight/photoz_kernels_cy.pyx  --cython-->  C in a temp dir
    --gcc (reference flags, setup.py:15-23: -ffast-math -fopenmp, -lm)-->
    oracle/_ref/delight/{utils_cy,photoz_kernels_cy}.<abi>.so
  * delight/{__init__,utils,io,photoz_kernels,photoz_gp}.py  --py_compile-->
    oracle/_ref/delight/*.bc    (sourceless byte-code, .pyc format; lets the GPU box time the
    reference's own Python layers for the CPU baseline -- /root/reference does not exist there.
    The suffix is .bc because *.pyc files are dropped from the snapshot sent to the GPU box;
    oracle/ref.py installs the matching importer)
  * scripts/{processFilters,processSEDs,simulateWithSEDs,delight-learn,delight-apply,
    templateFitting}.py  --py_compile-->  oracle/_ref/scripts/*.bc

The reference's own build system (setup.py) is NOT run: it imports sphinx (absent here).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import py_compile
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PYX = ["utils_cy", "photoz_kernels_cy"]
PY_MODULES = ["__init__", "utils", "io", "photoz_kernels", "photoz_gp"]
SCRIPTS = ["processFilters", "processSEDs", "simulateWithSEDs", "delight-learn",
           "delight-apply", "templateFitting"]

# reference flags, setup.py:15-23
CFLAGS = ["-O2", "-fPIC", "-ffast-math", "-fopenmp", "-Wno-uninitialized",
          "-Wno-maybe-uninitialized", "-Wno-unused-function", "-Wno-cpp",
          "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
LDFLAGS = ["-shared", "-fopenmp", "-lm"]


def _gcc():
    # the venv's default gcc may lack libgomp.spec; the system one has it
    for cand in ("/usr/bin/gcc", shutil.which("gcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no gcc found")


def ref_available(reference):
    return os.path.isfile(os.path.join(reference, "delight", "utils_cy.pyx"))


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.isfile(os.path.join(OUT, "delight", m + suffix)) for m in PYX) and \
        all(os.path.isfile(os.path.join(OUT, "delight", m + ".bc")) for m in PY_MODULES)


def build(reference="/root/reference", force=False, verbose=True):
    if not ref_available(reference):
        if built():
            return OUT
        raise FileNotFoundError(
            "reference tree %s not present and oracle/_ref not prebuilt" % reference)
    if built() and not force:
        return OUT
    import numpy
    from Cython.Build import cythonize  # noqa: F401  (probe)
    from Cython.Compiler.Main import compile as cy_compile
    from Cython.Compiler.Options import CompilationOptions, default_options

    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pkg_out = os.path.join(OUT, "delight")
    os.makedirs(pkg_out, exist_ok=True)
    os.makedirs(os.path.join(OUT, "scripts"), exist_ok=True)
    inc = [sysconfig.get_paths()["include"], numpy.get_include()]
    gcc = _gcc()
    with tempfile.TemporaryDirectory(prefix="delight_ref_build_") as tmp:
        for mod in PYX:
            src = os.path.join(reference, "delight", mod + ".pyx")
            c_out = os.path.join(tmp, mod + ".c")
            opts = CompilationOptions(default_options, output_file=c_out,
                                      language_level=3)
            res = cy_compile(src, opts, full_module_name="delight." + mod)
            if res.num_errors:
                raise RuntimeError("cython failed on " + src)
            so = os.path.join(pkg_out, mod + suffix)
            cmd = [gcc] + CFLAGS + ["-I" + i for i in inc] + [c_out, "-o", so] + LDFLAGS
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    for mod in PY_MODULES:
        py_compile.compile(os.path.join(reference, "delight", mod + ".py"),
                           cfile=os.path.join(pkg_out, mod + ".bc"),
                           dfile="<reference>/delight/%s.py" % mod, doraise=True)
    for s in SCRIPTS:
        py_compile.compile(os.path.join(reference, "scripts", s + ".py"),
                           cfile=os.path.join(OUT, "scripts", s + ".bc"),
                           dfile="<reference>/scripts/%s.py" % s, doraise=True)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(build(a.reference, a.force))
    sys.exit(0)

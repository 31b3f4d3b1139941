"""Loader for the compiled, unmodified reference in oracle/_ref -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The product path (delight_b200/) never does.

`load()` puts oracle/_ref first on sys.path, installs the three non-numeric shims that the
reference's Python layers need on this image and returns the imported
`delight` package of the reference:

  * scipy.misc.derivative (imported by delight/utils.py:4, removed from SciPy)  -> stub
  * matplotlib / matplotlib.pyplot (imported at the top of the preprocessing scripts) -> stub
  * coloredlogs -> stub
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def _install_shims():
    import scipy  # noqa: F401
    try:
        import scipy.misc as _m
    except Exception:  # scipy.misc removed entirely on newer SciPy
        _m = types.ModuleType("scipy.misc")
        sys.modules["scipy.misc"] = _m
        import scipy as _s
        _s.misc = _m
    if not hasattr(_m, "derivative"):
        _m.derivative = None
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")

            class _Anything:
                def __getattr__(self, name):
                    return _Anything()

                def __call__(self, *a, **k):
                    return _Anything()

                def __iter__(self):
                    return iter((_Anything(), _Anything()))

                def __getitem__(self, k):
                    return _Anything()

                def _arith(self, *a):
                    return _Anything()

                __mul__ = __rmul__ = __add__ = __radd__ = __sub__ = __rsub__ = _arith
                __truediv__ = __rtruediv__ = _arith

            def _module_getattr(name):
                if name.startswith("__"):
                    raise AttributeError(name)
                return _Anything()

            plt.subplots = lambda *a, **k: (_Anything(), _Anything())
            plt.__getattr__ = _module_getattr
            plt.__file__ = mpl.__file__ = "<delight oracle stub>"
            mpl.pyplot = plt
            mpl.use = lambda *a, **k: None
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "coloredlogs" not in sys.modules:
        try:
            import coloredlogs  # noqa: F401
        except Exception:
            cl = types.ModuleType("coloredlogs")
            cl.install = lambda *a, **k: None
            sys.modules["coloredlogs"] = cl


def available():
    d = os.path.join(REF_DIR, "delight")
    return os.path.isdir(d) and any(f.startswith("utils_cy") and f.endswith(".so")
                                    for f in os.listdir(d))


class _RefFinder(importlib.abc.MetaPathFinder):
    """Imports `delight` / `delight.<mod>` from oracle/_ref/delight: compiled extension modules
    (.so) and sourceless byte-code stored as <mod>.bc."""

    def find_spec(self, fullname, path=None, target=None):
        pkg_dir = os.path.join(REF_DIR, "delight")
        if fullname == "delight":
            init = os.path.join(pkg_dir, "__init__.bc")
            loader = importlib.machinery.SourcelessFileLoader(fullname, init)
            return importlib.util.spec_from_file_location(
                fullname, init, loader=loader, submodule_search_locations=[pkg_dir])
        if not fullname.startswith("delight.") or fullname.count(".") != 1:
            return None
        name = fullname.split(".")[1]
        for f in os.listdir(pkg_dir):
            if f.startswith(name + ".") and f.endswith(".so"):
                fp = os.path.join(pkg_dir, f)
                loader = importlib.machinery.ExtensionFileLoader(fullname, fp)
                return importlib.util.spec_from_file_location(fullname, fp, loader=loader)
        bc = os.path.join(pkg_dir, name + ".bc")
        if os.path.exists(bc):
            loader = importlib.machinery.SourcelessFileLoader(fullname, bc)
            return importlib.util.spec_from_file_location(fullname, bc, loader=loader)
        return None


def load():
    """Import and return the reference `delight` package from oracle/_ref."""
    if not available():
        raise ImportError("oracle/_ref is not built; run `python oracle/build_ref.py`")
    _install_shims()
    mod = sys.modules.get("delight")
    if mod is not None and not getattr(mod, "__file__", "").startswith(REF_DIR):
        raise ImportError("a different `delight` is already imported: %r" % (mod,))
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    delight = importlib.import_module("delight")
    for sub in ("utils_cy", "photoz_kernels_cy", "utils", "io", "photoz_kernels", "photoz_gp"):
        importlib.import_module("delight." + sub)
    return delight


def script_path(name):
    """Path of a byte-compiled reference script (run with runpy.run_path)."""
    return os.path.join(REF_DIR, "scripts", name + ".bc")

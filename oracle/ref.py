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


def load():
    """Import and return the reference `delight` package from oracle/_ref."""
    if not available():
        raise ImportError("oracle/_ref is not built; run `python oracle/build_ref.py`")
    _install_shims()
    mod = sys.modules.get("delight")
    if mod is not None and not getattr(mod, "__file__", "").startswith(REF_DIR):
        raise ImportError("a different `delight` is already imported: %r" % (mod,))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    delight = importlib.import_module("delight")
    for sub in ("utils_cy", "photoz_kernels_cy", "utils", "io", "photoz_kernels", "photoz_gp"):
        importlib.import_module("delight." + sub)
    return delight


def script_path(name):
    """Path of a byte-compiled reference script (run with `python <path> params.cfg`)."""
    return os.path.join(REF_DIR, "scripts", name + ".pyc")

"""Make the reference's scripts run unchanged on the CUDA operators.

    import delight_b200.install; delight_b200.install.install()

registers `delight_b200.utils_cy` and `delight_b200.photoz_kernels_cy` under the module names the
reference imports (`delight.utils_cy`, `delight.photoz_kernels_cy`; delight/photoz_kernels.py:9-11,
scripts/delight-apply.py:9), so that `from delight.utils_cy import approx_flux_likelihood_cy`
resolves to the B200 implementation while the rest of the `delight` package (io, photoz_gp,
photoz_kernels, utils) stays the reference's own Python.  Call it before the first
`import delight.photoz_kernels`.  `uninstall()` restores whatever was registered before.
"""
import sys

_saved = {}
_NAMES = {"delight.utils_cy": "utils_cy", "delight.photoz_kernels_cy": "photoz_kernels_cy"}


def install(package="delight"):
    import importlib
    from . import utils_cy, photoz_kernels_cy  # noqa: F401
    for full, short in _NAMES.items():
        name = full.replace("delight", package, 1)
        _saved.setdefault(name, sys.modules.get(name))
        mod = importlib.import_module("delight_b200." + short)
        sys.modules[name] = mod
        parent = sys.modules.get(package)
        if parent is not None:
            setattr(parent, short, mod)
    return True


def uninstall(package="delight"):
    for full, short in _NAMES.items():
        name = full.replace("delight", package, 1)
        old = _saved.pop(name, None)
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old
            parent = sys.modules.get(package)
            if parent is not None:
                setattr(parent, short, old)

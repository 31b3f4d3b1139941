"""ctypes binding of libdelight_b200.so (the C ABI of include/delight_b200.h).

The library is loaded lazily; a missing library or a missing CUDA device is a hard error --
there is no CPU path in this package.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdelight_b200.so")

_lib = None

c_ll = ctypes.c_longlong
c_d = ctypes.c_double
c_i = ctypes.c_int
c_p = ctypes.c_void_p

# name -> argtypes  (all return int unless listed in _RESTYPES)
_SIGS = {
    "dl_version": [],
    "dl_device_info": [c_p, c_p, c_p, c_p],
    "dl_approx_flux_likelihood_host": [c_p, c_ll, c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p],
    "dl_approx_flux_likelihood_dev": [c_p, c_ll, c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                      c_p],
    "dl_kernelparts_host": [c_ll, c_ll, c_ll, c_ll, c_d, c_d, c_ll, c_p, c_p, c_p, c_p, c_p, c_p,
                            c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p],
    "dl_kernelparts_dev": [c_ll, c_ll, c_ll, c_ll, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                           c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_ll, c_p],
    "dl_kernelparts_diag_host": [c_ll, c_ll, c_ll, c_d, c_d, c_ll, c_p, c_p, c_p, c_p, c_p, c_p,
                                 c_p, c_i, c_p, c_p, c_p, c_p],
    "dl_kernelparts_diag_dev": [c_ll, c_ll, c_ll, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                c_i, c_p, c_p, c_p, c_p, c_p],
    "dl_find_positions_host": [c_ll, c_ll, c_p, c_p, c_p],
    "dl_find_positions_dev": [c_ll, c_ll, c_p, c_p, c_p, c_p],
    "dl_kernel_parts_interp_host": [c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                    c_ll, c_ll],
    "dl_kernel_parts_interp_dev": [c_ll, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                   c_ll, c_ll, c_p],
    "dl_bilininterp_precomputedbins_host": [c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                            c_ll, c_ll],
    "dl_bilininterp_precomputedbins_dev": [c_ll, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_p,
                                           c_p, c_ll, c_ll, c_p],
    "dl_kernel_grids_dev": [c_ll, c_ll, c_ll, c_ll, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                            c_p, c_p, c_p, c_p],
    "dl_fused_chunk": [],
    "dl_fused_targets_per_cta": [c_i, c_i],
    "dl_fused_ctas_per_sm": [c_i, c_i],
    "dl_fused_padded_nz": [c_ll],
    "dl_fused_workspace_bytes": [c_ll, c_ll, c_ll, c_i],
    "dl_relayout_grid_dev": [c_ll, c_ll, c_ll, c_p, c_p, c_p, c_p],
    "dl_relayout_prior_dev": [c_ll, c_ll, c_p, c_p, c_p],
    "dl_zprior_dev": [c_ll, c_ll, c_p, c_p, c_d, c_p, c_p],
    "dl_template_prior_dev": [c_ll, c_ll, c_p, c_p, c_p, c_p, c_p],
    "dl_like_fused_dev": [c_ll, c_ll, c_ll, c_ll, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                          c_d, c_d, c_p, c_p, c_d, c_i, c_p, c_ll, c_p, c_p, c_p, c_p, c_ll, c_p],
    "dl_like_compressed_dev": [c_ll, c_ll, c_ll, c_ll, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_d,
                               c_p, c_p, c_d, c_i, c_p, c_i, c_p, c_ll, c_p, c_p],
    "dl_metrics_dev": [c_ll, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p],
    "dl_ci_levels_dev": [c_ll, c_ll, c_p, c_ll, c_p, c_i, c_p, c_i, c_p, c_p],
    "dl_gp_fit_dev": [c_ll, c_ll, c_ll, c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_d, c_p,
                      c_p, c_p, c_p, c_p, c_p, c_ll, c_p],
    "dl_gp_predict_dev": [c_ll, c_ll, c_ll, c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                          c_d, c_d, c_p, c_ll, c_p, c_p, c_p, c_p],
    "dl_gp_tables_bytes": [c_ll, c_ll],
    "dl_gp_tables_dev": [c_ll, c_ll, c_p, c_p, c_p, c_p, c_p],
    "dl_gp_predict_at_dev": [c_ll, c_ll, c_ll, c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                             c_d, c_d, c_p, c_ll, c_p, c_p, c_p, c_p],
    "dl_launch_count": [],
    "dl_timing_enable": [c_i],
    "dl_timing_collect": [c_p, c_p],
    "dl_fp64_peak": [c_i, c_p],
    "dl_selftest_math_dev": [c_ll, c_p, c_p, c_p],
}
_RESTYPES = {"dl_fused_padded_nz": c_ll, "dl_gp_tables_bytes": c_ll, "dl_launch_count": c_ll, "dl_fused_workspace_bytes": c_ll,
             "dl_last_error": ctypes.c_char_p}

EXPORTED = sorted(list(_SIGS) + ["dl_last_error"])


class DelightB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DelightB200Error(
                "%s is missing: build it with `python -m delight_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.dl_last_error.restype = ctypes.c_char_p
        L.dl_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, c_i)
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().dl_last_error()
        raise DelightB200Error("libdelight_b200 error %d: %s" % (rc, msg.decode() if msg else ""))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise DelightB200Error("no CUDA device: delight_b200 has no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else ctypes.c_void_p(t.data_ptr())

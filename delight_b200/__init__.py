"""delight_b200 -- B200-native (sm_100a) implementation of Delight's photo-z inference hot path.

Python stays the host; every numeric kernel lives in libdelight_b200.so (hand-written CUDA,
C ABI declared in include/delight_b200.h).  There is no CPU fallback: importing the operator
modules works anywhere, calling them without the built library or without a CUDA device raises.

Modules
    utils_cy, photoz_kernels_cy   drop-in operator modules (same names/signatures as the
                                  reference's Cython modules delight.utils_cy and
                                  delight.photoz_kernels_cy)
    apply                         batched, fused apply path (likelihood + PDF + evidences + top-K)
    gp                            batched GP fit / predict on the device
    install                       aliases the drop-in modules as delight.utils_cy / ... so that
                                  the reference's scripts run unchanged
"""
__version__ = "0.1.0"

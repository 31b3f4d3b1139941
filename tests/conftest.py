import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Make sure the C-ABI library exists (nvcc cross-compiles without a GPU).  On the GPU box
    the in-tree .so travels with the snapshot and is up to date, so this is a no-op there."""
    from delight_b200 import build as b
    if not os.path.exists(b.LIB) or (b._stale() and os.path.exists("/usr/local/cuda/bin/nvcc")):
        try:
            b.build()
        except Exception as e:          # tests that need the library will fail loudly
            print("delight_b200 build failed: %s" % e)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


GUARD = 512                 # bytes of sentinel on each side of a guarded device allocation
GUARD_MAX = 64 << 20        # larger allocations are left alone (the big-size property tests)
GUARD_TOTAL = 4 << 30


@pytest.fixture(autouse=True)
def guard_bands(request, monkeypatch):
    """Own bounds check for every GPU test (compute-sanitizer is not available on the GPU pool):
    torch.empty / zeros / empty_like for CUDA tensors -- which is how the host layer allocates
    every output and workspace the kernels write -- hand out the interior of a slightly larger
    buffer whose first and last 512 bytes hold a sentinel.  After the test every sentinel must
    be intact: a kernel that writes before or past the end of its output fails the test that
    ran it."""
    if request.node.get_closest_marker("gpu") is None:
        yield None
        return
    import torch
    if not torch.cuda.is_available():
        yield None
        return
    real_empty, real_zeros, real_empty_like = torch.empty, torch.zeros, torch.empty_like
    records = []
    total = [0]

    def guarded(shape, dtype, device):
        dtype = dtype or torch.get_default_dtype()
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * real_empty((), dtype=dtype).element_size()
        if nbytes == 0 or nbytes > GUARD_MAX or total[0] + nbytes > GUARD_TOTAL:
            return None
        flat = real_empty(nbytes + 2 * GUARD, dtype=torch.uint8, device=device)
        flat[:GUARD].fill_(0xA5)
        flat[GUARD + nbytes:].fill_(0xA5)
        records.append((flat, nbytes, tuple(int(s) for s in shape)))
        total[0] += nbytes
        return flat[GUARD:GUARD + nbytes].view(dtype).view(tuple(int(s) for s in shape))

    def is_cuda(device):
        return device is not None and torch.device(device).type == "cuda"

    def shape_of(size):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            return tuple(size[0])
        return tuple(size)

    def empty(*size, dtype=None, device=None, **kw):
        if is_cuda(device) and not kw:
            t = guarded(shape_of(size), dtype, device)
            if t is not None:
                return t
        return real_empty(*size, dtype=dtype, device=device, **kw)

    def zeros(*size, dtype=None, device=None, **kw):
        if is_cuda(device) and not kw:
            t = guarded(shape_of(size), dtype, device)
            if t is not None:
                return t.zero_()
        return real_zeros(*size, dtype=dtype, device=device, **kw)

    def empty_like(x, **kw):
        if x.is_cuda and not kw and x.is_contiguous():
            t = guarded(tuple(x.shape), x.dtype, x.device)
            if t is not None:
                return t
        return real_empty_like(x, **kw)

    def check():
        torch.cuda.synchronize()
        for flat, nbytes, shape in records:
            lo = flat[:GUARD].cpu()
            hi = flat[GUARD + nbytes:].cpu()
            assert bool((lo == 0xA5).all()) and bool((hi == 0xA5).all()), \
                "a kernel wrote outside a %s-byte buffer of shape %s" % (nbytes, shape)

    monkeypatch.setattr(torch, "empty", empty)
    monkeypatch.setattr(torch, "zeros", zeros)
    monkeypatch.setattr(torch, "empty_like", empty_like)
    yield type("Guards", (), dict(records=records, check=staticmethod(check), size=GUARD))
    monkeypatch.undo()
    check()
    request.config._guarded = getattr(request.config, "_guarded", 0) + len(records)


def pytest_terminal_summary(terminalreporter, config):
    n = getattr(config, "_guarded", 0)
    if n:
        terminalreporter.write_line("guard bands: %d device allocations checked, all intact" % n)

"""CPU-only checks: the C-ABI library loads and exports every symbol of include/delight_b200.h,
host-side logic (band grouping, rank partition, argument validation), no compute calls."""
import os
import re

import numpy as np
import pytest

import delight_b200
from delight_b200 import _lib, apply, utils_cy, photoz_kernels_cy

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(REPO, "include", "delight_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libdelight_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTED) == names
    assert L.dl_version() >= 100


def test_padding_and_workspace_queries():
    L = _lib.lib()
    c = L.dl_fused_chunk()
    assert c == 128
    assert L.dl_fused_targets_per_cta(6, 0) == 32 and L.dl_fused_ctas_per_sm(6, 0) == 2
    assert L.dl_fused_targets_per_cta(6, 1) == 64 and L.dl_fused_ctas_per_sm(21, 0) == 1
    assert L.dl_fused_padded_nz(1500) == 1536 and L.dl_fused_padded_nz(128) == 128
    a = L.dl_fused_workspace_bytes(32, 1000, 1500, 0)
    b = L.dl_fused_workspace_bytes(64, 1000, 1500, 0)
    assert 0 < a < b
    # the raw evidences (8 B per target x training galaxy) dominate: no per-z-chunk partials
    big = L.dl_fused_workspace_bytes(1000, 100000, 1500, 0)
    assert big < 1000 * (8 * 100000 + 8 * 1536 + 4096)
    assert L.dl_fused_workspace_bytes(1000, 100000, 1500, 1) < 1000 * (8 * 1536 + 4096)


def test_rank_partition_matches_reference_formula():
    # scripts/delight-apply.py:50-53
    for n in (0, 1, 7, 1000, 12345):
        for size in (1, 2, 3, 8):
            spans = [apply.partition(n, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            for r, (f, l) in enumerate(spans):
                assert f == int(r * n / float(size))
                assert l == int(min(n, (r + 1) * n / float(size)))


def test_target_grouping_by_band_set():
    fl = np.arange(20.0).reshape(4, 5) + 1
    bands = np.array([[0, 1, 2, 3, 4], [0, 1, 3, 4, -1], [0, 1, 2, 3, 4], [1, 2, -1, -1, -1]])
    nf = np.array([5, 4, 5, 2])
    t = apply.Targets(fl, fl, bands, nf)
    groups = {b: list(i) for b, i in t.groups()}
    assert groups == {(0, 1, 2, 3, 4): [0, 2], (0, 1, 3, 4): [1], (1, 2): [3]}
    with pytest.raises(ValueError):
        apply.Targets(fl, fl, bands, np.array([5, 4, 5, 0]))
    bad = fl.copy()
    bad[1, 2] = 0.0                      # zero variance in a valid slot
    with pytest.raises(ValueError):
        apply.Targets(fl, bad, bands, nf)
    bad[1, 2] = 1.0
    bad[3, 4] = -7.0                     # garbage in an unused slot is fine
    apply.Targets(fl, bad, bands, nf)


def test_dropin_signatures_and_dtype_errors():
    import inspect
    sig = inspect.signature(utils_cy.approx_flux_likelihood_cy)
    assert list(sig.parameters) == ["like", "nz", "nt", "nf", "f_obs", "f_obs_var", "f_mod",
                                    "f_mod_covar", "ell_hat", "ell_var"]
    sig = inspect.signature(photoz_kernels_cy.kernelparts)
    assert list(sig.parameters)[-5:] == ["KL", "KC", "D_alpha_C", "D_alpha_L", "D_alpha_z"]
    like = np.zeros((3, 2), dtype=np.float32)
    with pytest.raises(ValueError, match="dtype mismatch"):
        utils_cy.approx_flux_likelihood_cy(like, 3, 2, 1, np.ones(1), np.ones(1),
                                           np.ones((3, 2, 1)), np.ones((3, 2, 1)), np.ones(3),
                                           np.ones(3))
    with pytest.raises(ValueError, match="dimensions"):
        utils_cy.find_positions(2, 3, np.ones((2, 1)), np.zeros(2, dtype=np.int64), np.ones(3))


def test_bilininterp_shim_checks_its_arguments_before_touching_the_device():
    """utils_cy.pyx:25-50 indexes without bounds checks; the shim refuses inconsistent shapes and
    bins outside the grids (and treats empty inputs as a no-op) before any device work."""
    import inspect
    sig = inspect.signature(utils_cy.bilininterp_precomputedbins)
    assert list(sig.parameters) == ["numBands", "nobj", "Kinterp", "v1s", "v2s", "p1s", "p2s",
                                    "grid1", "grid2", "Kgrid"]
    g1, g2 = np.linspace(0., 1., 4), np.linspace(0., 1., 5)
    K = np.ones((2, 4, 5))
    out = np.zeros((2, 3))
    v = np.full(3, 0.5)
    p = np.ones(3, dtype=np.int64)
    utils_cy.bilininterp_precomputedbins(2, 0, out, v, v, p, p, g1, g2, K)        # no-op
    with pytest.raises(ValueError, match="inconsistent"):
        utils_cy.bilininterp_precomputedbins(3, 3, out, v, v, p, p, g1, g2, K)    # 3 bands > 2
    with pytest.raises(ValueError, match="bin positions"):
        utils_cy.bilininterp_precomputedbins(2, 3, out, v, v, p + 2, p, g1, g2, K)
    with pytest.raises(ValueError, match="dtype mismatch"):
        utils_cy.bilininterp_precomputedbins(2, 3, out, v, v, p.astype(np.int32), p, g1, g2, K)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.DelightB200Error):
        utils_cy.find_positions(2, 3, np.array([0.5, 1.5]), np.zeros(2, dtype=np.int64),
                                np.array([0., 1., 2.]))


def test_product_never_references_the_oracle():
    pkg = os.path.join(REPO, "delight_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt, os.path.join(root, f)


def test_catalogue_reader_matches_reference_parsing(golden, tmp_path):
    """pipeline.read_catalogue vs what the reference's io.getDataFromFile yielded for the same
    rows (valid-band lists, variance inflation, skip rule)."""
    from delight_b200 import pipeline
    g = golden("sdss_pipeline")
    wd = str(tmp_path)
    np.savetxt(wd + "/tr.txt", g["train_cat"])
    np.savetxt(wd + "/ta.txt", g["target_cat"])
    names = ["U_SDSS", "G_SDSS", "R_SDSS", "I_SDSS", "Z_SDSS"]
    p = dict(bandNames=names, training_catFile=wd + "/tr.txt", target_catFile=wd + "/ta.txt",
             training_bandOrder=str(g["train_bandOrder"]).split(" "),
             target_bandOrder=str(g["target_bandOrder"]).split(" "),
             training_referenceBand="I_SDSS", target_referenceBand="I_SDSS",
             training_extraFracFluxError=float(g["extraFrac"]))
    for prefix, key in (("training_", "tr"), ("target_", "ta")):
        cat, n_lines = pipeline.read_catalogue(p, prefix)
        nb = g[key + "_nb"]
        assert len(cat) == nb.size and n_lines >= nb.size
        np.testing.assert_array_equal(cat.nf, nb)
        np.testing.assert_array_equal(cat.ztrue, g[key + "_z"])
        for i in range(nb.size):
            B = nb[i]
            np.testing.assert_array_equal(cat.bands[i, :B], g[key + "_bands"][i, :B])
            np.testing.assert_array_equal(cat.fluxes[i, :B], g[key + "_flux"][i, :B])
            np.testing.assert_allclose(cat.fluxes_var[i, :B], g[key + "_var"][i, :B], rtol=1e-15)


def test_fused_kernel_binary_properties():
    """Static checks of the built sm_100a code (cuobjdump, no GPU): every instantiation of the
    fused kernel fits 128 registers (16 warps per SM in both configurations), and the 6-band main
    pass is what DESIGN 5.1 says it is -- TMA bulk copies into an mbarrier ring, MUFU-seeded
    FP64 reciprocals / rsqrt, an FP64 FMA stream with next to no local memory traffic."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    lib = _lib.lib_path() if hasattr(_lib, "lib_path") else os.path.join(
        os.path.dirname(delight_b200.__file__), "libdelight_b200.so")
    usage = subprocess.run([exe, "--dump-resource-usage", lib], stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, text=True).stdout
    regs = [int(m) for m in re.findall(r"fused_like_kernel\S*:\s*\n\s*REG:(\d+)", usage)]
    assert len(regs) == 32 * 3 * 2          # 1..32 bands x (plain, z-dependent, band lists) x 2 modes
    assert max(regs) <= 128
    assert "sm_100a" in usage or "sm_100" in usage
    name = "_ZN2dl17fused_like_kernelILi6ELb0ELb0ELi0EEEvNS_11FusedParamsE"
    sass = subprocess.run([exe, "-sass", "-fun", name, lib], stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", sass, re.M)
    count = lambda prefix: sum(1 for o in ops if o.startswith(prefix))
    assert count("UBLKCP") >= 14            # 6 means + 6 variances + prior + weights per stage
    assert count("SYNCS") > 0               # mbarrier arrive / try_wait
    assert count("MUFU.RCP64H") > 0 and count("MUFU.RSQ64H") > 0
    fp64 = count("DFMA") + count("DMUL") + count("DADD")
    assert fp64 > 600
    assert count("LDL") + count("STL") < 0.05 * fp64
    assert count("DMMA") == count("HMMA") == 0      # nothing here is a dense contraction


def test_c_abi_rejects_bad_arguments_before_any_device_work():
    """Error convention of include/delight_b200.h: invalid shapes come back as DL_ERR_INVALID with
    a message from dl_last_error(), decided on the host before anything is launched (so this runs
    without a GPU).  The reference operators index without checks (boundscheck=False)."""
    import ctypes
    L = _lib.lib()
    ll = ctypes.c_longlong
    one = (ctypes.c_int * 1)(0)
    buf = ctypes.create_string_buffer(64)
    P = ctypes.cast(buf, ctypes.c_void_p)

    def fused(nO=4, nT=5, nz=10, nB=3, nf=3, n_union=3, K=2, ld=10, cov=1.0, ws=P, wsb=64):
        return L.dl_like_fused_dev(nO, nT, nz, nB, nf, n_union, one, None, P, P, P, P, P, 1.0, 1.0,
                                   None, None, cov, K, P, ld, None, P, P, ws, wsb, None)
    cases = [(dict(nf=0), "bands per target"), (dict(nf=33, n_union=33), "bands per target"),
             (dict(K=6), "K <= min"), (dict(ld=9), "ld_pdf"), (dict(cov=-1.0), "cov_scale"),
             (dict(n_union=2), "band union"), (dict(nO=0), "bad shape"),
             (dict(), "workspace too small")]
    for kw, msg in cases:
        assert fused(**kw) == 1, kw
        assert msg in L.dl_last_error().decode(), (kw, L.dl_last_error())
    assert L.dl_bilininterp_precomputedbins_host(ll(1), ll(1), P, P, P, P, P, P, P, P, ll(1),
                                                 ll(5)) == 1
    assert "bad sizes" in L.dl_last_error().decode()
    assert L.dl_metrics_dev(ll(0), ll(10), P, ll(10), P, None, P, None) == 1
    assert L.dl_fused_workspace_bytes(ll(64), ll(1000), ll(300), 1) < \
        L.dl_fused_workspace_bytes(ll(64), ll(1000), ll(300), 0)


def test_header_is_plain_c_and_links(tmp_path):
    """The drop-in boundary is a C ABI: include/delight_b200.h compiles as pedantic C99 (no C++ or
    torch types in any signature) and a C program links against the library and calls it."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "delight_b200.h"\n'
                   'int main(void) { return dl_version() > 0 && dl_fused_padded_nz(1500) == 1536 '
                   '? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    libdir = os.path.dirname(delight_b200.__file__)
    exe = str(tmp_path / "abi")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc,
                        str(src), "-o", exe, "-L", libdir, "-ldelight_b200",
                        "-Wl,-rpath," + libdir], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    assert r.returncode == 0, r.stdout
    assert subprocess.run([exe]).returncode == 0


def test_binary_catalogues_read_like_the_text_files(golden, tmp_path):
    """a catalogue named *.npy in the cfg is the same table in binary.  The reader
    must return exactly what it returns for the text file -- whole file and a rank's slice --
    and the row counters / savers must agree on both formats."""
    from delight_b200 import pipeline
    g = golden("sdss_pipeline")
    wd = str(tmp_path)
    np.savetxt(wd + "/ta.txt", g["target_cat"], fmt="%.18e")
    np.save(wd + "/ta.npy", np.loadtxt(wd + "/ta.txt", ndmin=2))
    names = ["U_SDSS", "G_SDSS", "R_SDSS", "I_SDSS", "Z_SDSS"]
    base = dict(bandNames=names, target_bandOrder=str(g["target_bandOrder"]).split(" "),
                target_referenceBand="I_SDSS",
                training_extraFracFluxError=float(g["extraFrac"]))   # (the reader's training-key quirk)
    n = g["target_cat"].shape[0]
    assert pipeline._count_rows(wd + "/ta.txt") == pipeline._count_rows(wd + "/ta.npy") == n
    for first, last in ((None, None), (3, n - 2)):
        got = {}
        for ext in ("txt", "npy"):
            p = dict(base, target_catFile=wd + "/ta." + ext)
            got[ext] = pipeline.read_catalogue(p, "target_", first, last) \
                if first is not None else pipeline.read_catalogue(p, "target_")
        (a, na), (b, nb) = got["txt"], got["npy"]
        assert na == nb and len(a) == len(b) > 0
        for f in ("fluxes", "fluxes_var", "bands", "nf", "ztrue"):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    arr = np.arange(12.).reshape(4, 3) / 7
    pipeline._save(wd + "/o.npy", arr, "%.2e")
    pipeline._save(wd + "/o.txt", arr, "%.2e")
    np.testing.assert_array_equal(pipeline._load(wd + "/o.npy"), arr)               # lossless
    np.testing.assert_allclose(pipeline._load(wd + "/o.txt"), arr, rtol=5e-3)       # %.2e text

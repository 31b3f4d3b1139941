"""Pins the CPU oracle (oracle/) against the reference's own golden vectors and against the
outputs of the unmodified reference captured in tests/golden/ (made by make_golden.py)."""
import numpy as np
import pytest

from oracle import oracle as orc

RTOL = 1e-11   # oracle (no fast-math) vs reference (-ffast-math): re-association noise only


def test_like_matches_reference_and_twin(golden):
    g = golden("like_small")
    for c in "abcde":
        f_mod = g[c + "_f_mod"]
        nz, nt, nf = f_mod.shape
        like = np.zeros((nz, nt))
        orc.approx_flux_likelihood_cy(like, nz, nt, nf, g[c + "_f_obs"], g[c + "_f_var"], f_mod,
                                      g[c + "_f_cov"], g[c + "_ell_hat"], g[c + "_ell_var"])
        np.testing.assert_allclose(like, g[c + "_like"], rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(like, g[c + "_twin"], rtol=RTOL, atol=1e-300)
        assert like.max() == 1.0


def test_find_positions_reference_golden(golden):
    # tests/test_photoz_kernels_cy.py:62-68
    a = np.array([0., 1., 2., 3., 4.])
    b = np.array([0.5, 2.5, 3.0, 3.1, 4.0])
    pos = np.zeros(b.size, dtype=np.int64)
    orc.find_positions(b.size, a.size, b, pos, a)
    np.testing.assert_array_equal(pos, [0, 2, 2, 3, 3])
    g = golden("kernelparts_small")
    pos2 = np.full(5, 7, dtype=np.int64)
    orc.find_positions(5, 5, g["fp_vals2"], pos2, g["fp_grid"])
    np.testing.assert_array_equal(pos2, g["fp_pos2"])


def _kp_args(g, NL=3):
    amp = g["amp"]
    return (amp.shape[1], NL, float(g["alpha_C"]), float(g["alpha_L"]), amp, g["mu"], g["sig"],
            g["lines_mu"][:NL], g["lines_sig"][:NL], g["norms"])


def test_kernelparts_matches_reference(golden):
    g = golden("kernelparts_small")
    b1, b2, fz1, fz2 = g["b1"], g["b2"], g["fz1"], g["fz2"]
    NO1, NO2 = len(b1), len(b2)
    outs = [np.zeros((NO1, NO2)) for _ in range(5)]
    orc.kernelparts(NO1, NO2, *_kp_args(g), b1, fz1, b2, fz2, True, *outs)
    for o, k in zip(outs, ("KL", "KC", "DaC", "DaL", "Daz")):
        np.testing.assert_allclose(o, g[k], rtol=1e-10, atol=1e-300)
    orc.kernelparts(NO1, NO2, *_kp_args(g), b1, fz1, b2, fz2, True, *outs)
    for o, k in zip(outs, ("KL2", "KC2", "DaC2", "DaL2", "Daz2")):
        np.testing.assert_allclose(o, g[k], rtol=1e-10, atol=1e-300)
    outs = [np.full((NO1, NO2), 0.5) for _ in range(5)]
    orc.kernelparts(NO1, NO2, *_kp_args(g), b1, fz1, b2, fz2, False, *outs)
    np.testing.assert_allclose(outs[0], g["KL_ng"], rtol=1e-10)
    np.testing.assert_allclose(outs[1], g["KC_ng"], rtol=1e-10)
    np.testing.assert_array_equal(outs[2], g["DaC_ng"])
    outs = [np.zeros((NO1, NO2)) for _ in range(5)]
    orc.kernelparts(NO1, NO2, *_kp_args(g, NL=0), b1, fz1, b2, fz2, True, *outs)
    np.testing.assert_allclose(outs[0], g["KL_nl0"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(outs[1], g["KC_nl0"], rtol=1e-10)


def test_kernelparts_diag_relation_and_reference(golden):
    # the reference's known relation (tests/test_photoz_kernels_cy.py:17-59):
    # kernelparts_diag == diag(kernelparts) for KL, KC, D_alpha_C, D_alpha_L
    g = golden("kernelparts_small")
    b1, fz1 = g["b1"], g["fz1"]
    n = len(b1)
    d = [np.full(n, 0.25) for _ in range(4)]
    orc.kernelparts_diag(n, *_kp_args(g), b1, fz1, True, *d)
    for o, k in zip(d, ("KLd", "KCd", "DaCd", "DaLd")):
        np.testing.assert_allclose(o, g[k], rtol=1e-10)
    full = [np.zeros((n, n)) for _ in range(5)]
    orc.kernelparts(n, n, *_kp_args(g), b1, fz1, b1, fz1, True, *full)
    d0 = [np.zeros(n) for _ in range(4)]
    orc.kernelparts_diag(n, *_kp_args(g), b1, fz1, True, *d0)
    for a, b in zip(d0, full[:4]):
        np.testing.assert_allclose(a, np.diag(b), rtol=1e-12)


def test_kernel_parts_interp_matches_reference(golden):
    g = golden("kernelparts_small")
    out = np.zeros_like(g["ki_out"])
    orc.kernel_parts_interp(out.shape[0], out.shape[1], out, g["b1"], g["ki_x1"], g["ki_p1"],
                            g["b2"], g["ki_x2"], g["ki_p2"], g["ki_fzg"], g["ki_grid"])
    np.testing.assert_allclose(out, g["ki_out"], rtol=1e-12)


@pytest.fixture(scope="module")
def sdss(golden):
    return golden("sdss_pipeline")


def test_kernel_grids_match_reference(sdss):
    grids = orc.kernel_grids(sdss["amp"], sdss["mu"], sdss["sig"], sdss["lines_pos"],
                             float(sdss["alpha_C"]), float(sdss["alpha_L"]), sdss["zgridGP"])
    np.testing.assert_allclose(grids["KC"], sdss["KC_grid"], rtol=1e-10)
    np.testing.assert_allclose(grids["KL"], sdss["KL_grid"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(grids["KCd"], sdss["KCd"], rtol=1e-10)
    np.testing.assert_allclose(grids["KLd"], sdss["KLd"], rtol=1e-10, atol=1e-300)


def test_gp_fit_and_predict_match_reference(sdss):
    g = sdss
    grids = dict(KC=g["KC_grid"], KL=g["KL_grid"], KCd=g["KCd"], KLd=g["KLd"],
                 fzGrid=1 + g["zgridGP"])
    nt = g["tmpl"].shape[0]
    rows = g["gp_rows"]
    for i in range(rows.shape[0]):
        B = int(g["tr_nb"][i])
        bands = g["tr_bands"][i, :B]
        fit = orc.gp_fit(grids, float(g["V_C"]), float(g["V_L"]), g["tmpl"], g["zgrid"], bands,
                         g["tr_z"][i], g["tr_flux"][i, :B], g["tr_var"][i, :B])
        assert B == int(rows[i, 0])
        np.testing.assert_allclose(fit["ell"], rows[i, 2], rtol=1e-12)
        core = rows[i, 3 + B:3 + B + nt + B * (B + 1) // 2 + B]
        np.testing.assert_allclose(fit["core"], core, rtol=1e-9, atol=1e-12 * np.abs(core).max())
        mm, mv = orc.gp_predict_interp(grids, float(g["V_C"]), float(g["V_L"]), g["tmpl"],
                                       g["zgrid"], g["zgridGP"], bands, g["tr_z"][i], fit["ell"],
                                       fit["bestType"], fit["L"], fit["D"], g["zgrid"])
        np.testing.assert_allclose(mm, g["model_mean"][:, i, :], rtol=1e-9)
        np.testing.assert_allclose(mv, g["model_covar"][:, i, :], rtol=1e-11)


def _core(g, loc):
    rows = g["gp_rows"]
    nt = g["tmpl"].shape[0]
    B = int(rows[loc, 0])
    flat = rows[loc, 3 + B:]
    L = np.zeros((B, B))
    L[np.tril_indices(B)] = flat[nt:nt + B * (B + 1) // 2]
    D = flat[nt + B * (B + 1) // 2:nt + B * (B + 1) // 2 + B]
    return (rows[loc, 3:3 + B].astype(int), rows[loc, 1], rows[loc, 2], int(np.argmax(flat[:nt])),
            L, D)


def test_model_variance_against_long_double_truth(sdss):
    """The model variance Kdiag - k^T A^-1 k is a cancelling difference.  How far is the
    reference (scipy cho_solve + dot, photoz_gp.py:273-277) from the same formulas evaluated in
    long double on the same inputs?  1.7e-14 on this mock (the GPU's
    |L^-1 k|^2 form: 1.1e-14, tests/test_gpu_parity.py) -- i.e. no precision is lost that would
    justify more than 1e-12 of tolerance between implementations."""
    g = sdss
    grids = dict(KC=g["KC_grid"], KL=g["KL_grid"], KCd=g["KCd"], KLd=g["KLd"],
                 fzGrid=1.0 + g["zgridGP"])
    worst_ref = worst_orc = 0.0
    for loc in (0, 3, 7, 11):
        bands, z, ell, best, L, D = _core(g, loc)
        args = (grids, float(g["V_C"]), float(g["V_L"]), g["tmpl"], g["zgrid"], g["zgridGP"],
                bands, z, ell, best, L, D, g["zgrid"])
        hm, hv = orc.gp_predict_interp_hp(*args)
        dm, dv = orc.gp_predict_interp(*args)
        hv, hm = hv.astype(float), hm.astype(float)
        worst_ref = max(worst_ref, np.abs(g["model_covar"][:, loc, :] / hv - 1).max())
        worst_orc = max(worst_orc, np.abs(dv / hv - 1).max())
        np.testing.assert_allclose(g["model_mean"][:, loc, :], hm, rtol=1e-14)
    assert worst_ref < 1e-12 and worst_orc < 1e-12, (worst_ref, worst_orc)


def test_crossvalidation_matches_reference_script(golden, sdss):
    """delight-learn.py:92-100 restated: GP at the galaxy's own redshift, per held-out band
    -0.5 (model - flux)^2 / (model_var + flux_var), against the CV file the real script wrote
    (tests/golden/make_golden.py:gen_cv)."""
    g = sdss
    cvg = golden("cv_small")
    tr = cvg["train_cat"]
    grids = dict(KC=g["KC_grid"], KL=g["KL_grid"], KCd=g["KCd"], KLd=g["KLd"],
                 fzGrid=1.0 + g["zgridGP"])
    extra = float(g["extraFrac"])
    cv_bands, cols, vcols = [2, 4], [2, 4], [7, 9]           # R, Z of the SDSS cfg
    got = np.zeros_like(cvg["cv"])
    for loc in range(tr.shape[0]):
        bands, z, ell, best, L, D = _core(g, loc)
        mm, mv = orc.gp_predict_interp(grids, float(g["V_C"]), float(g["V_L"]), g["tmpl"],
                                       g["zgrid"], g["zgridGP"], bands, z, ell, best, L, D,
                                       np.array([z]))
        for j, (b, c, vc) in enumerate(zip(cv_bands, cols, vcols)):
            f, v = tr[loc, c], tr[loc, vc]
            if np.isfinite(f) and np.isfinite(v) and f > 0 and v > 0:
                got[loc, j] = -0.5 * (mm[0, b] - f) ** 2 / (mv[0, b] + v + (extra * f) ** 2)
    np.testing.assert_allclose(got, cvg["cv"], rtol=1e-10)
    assert cvg["cv"][5, 0] == 0.0                            # the galaxy without an R flux


def test_apply_fused_matches_reference(sdss):
    g = sdss
    nz = g["zgrid"].size
    K = g["top_idx"].shape[1]
    ell_hat = np.ones(nz)
    ell_var = np.full(nz, float(g["ellPriorSigma"]) ** 2)
    res = orc.apply_fused(g["ta_flux"], g["ta_var"], g["ta_nb"], g["ta_bands"], g["model_mean"],
                          g["model_covar"], g["prior"], g["zgrid"], ell_hat, ell_var, K,
                          want_evid=True)
    np.testing.assert_allclose(res["pdf"], g["pdfs"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(res["evid"], g["evid"], rtol=1e-10, atol=1e-300)
    np.testing.assert_array_equal(res["top_idx"], g["top_idx"])
    np.testing.assert_allclose(res["top_ev"], g["top_ev"], rtol=1e-10, atol=1e-300)
    for o in range(res["pdf"].shape[0]):
        m, _ = orc.metrics_core(g["ta_z"][o], g["zgrid"], res["pdf"][o])
        np.testing.assert_allclose(m, g["metrics"][o, :7], rtol=1e-9)
        for j, cl in enumerate(g["confidenceLevels"]):
            lev = orc.ci_level(g["zgrid"], res["pdf"][o], 1.0 - cl)
            np.testing.assert_allclose(lev, g["metrics"][o, 7 + j], rtol=1e-9)


def test_apply_fused_tie_rule():
    # two identical training galaxies -> equal evidences -> larger index first
    rng = np.random.default_rng(0)
    nz, nT, nB = 12, 4, 3
    mm = rng.uniform(1, 2, (nz, nT, nB))
    mc = rng.uniform(.1, .2, (nz, nT, nB))
    mm[:, 3] = mm[:, 1]
    mc[:, 3] = mc[:, 1]
    prior = np.ones((nz, nT))
    z = np.linspace(0.1, 1, nz)
    res = orc.apply_fused(rng.uniform(1, 2, (1, nB)), rng.uniform(.1, .2, (1, nB)),
                          np.array([nB]), np.arange(nB)[None, :], mm, mc, prior, z,
                          np.ones(nz), np.full(nz, .25), 4, want_evid=True)
    idx = list(res["top_idx"][0])
    assert idx.index(3) + 1 == idx.index(1)
    ref = np.argsort(res["evid"][0], kind="stable")[::-1]
    np.testing.assert_array_equal(res["top_idx"][0], ref)


def test_templatefit_twin(golden):
    g = golden("templatefit")
    s = golden("sdss_pipeline")
    f_mod = g["f_mod"]
    nz, nt, nB = f_mod.shape
    zero = np.zeros_like(f_mod)
    n = g["pdfs"].shape[0]
    # targets of the sdss fixture are the same catalogue rows
    res = orc.apply_fused(s["ta_flux"][:n], s["ta_var"][:n], s["ta_nb"][:n], s["ta_bands"][:n],
                          f_mod, zero, g["p_z"], s["zgrid"], np.ones(nz), np.full(nz, 1e24), 0)
    np.testing.assert_allclose(res["pdf"], g["pdfs"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(orc.template_prior(s["zgrid"], g["p_t"], g["p_z_t"]), g["p_z"],
                               rtol=1e-14)


def test_compressed_second_pass_matches_reference_twin(golden):
    """oracle.apply_compressed vs the reference's NumPy twin run on the stored columns
    (tests/golden/make_golden.py:gen_compressed_small; delight-apply.py:123-135,145-146)."""
    g = golden("compressed_small")
    nO = g["fluxes"].shape[0]
    for tag, norm, eh, ev in (("n0", False, g["ell_hat"], g["ell_var"]),
                              ("n1", True, g["ell_hat"], g["ell_var"]), ("s0", False, 1.0, 0.25)):
        r = orc.apply_compressed(g["fluxes"], g["fluxes_var"], g["nf"], g["bands"],
                                 g["model_mean"], g["model_covar"], g["prior"], g["zgrid"],
                                 eh, ev, g["indices"], normalized=norm)
        np.testing.assert_allclose(r["pdf"], g["pdf_" + tag], rtol=1e-13)
        for o in range(nO):
            n = r["sel"][o].size
            np.testing.assert_array_equal(r["sel"][o], g["sel_" + tag][o, :n])
            assert np.all(g["sel_" + tag][o, n:] == -1)
            np.testing.assert_allclose(r["evid"][o], g["evid_" + tag][o, :n], rtol=1e-13)
    # duplicate and out-of-range stored indices select fewer columns
    assert (g["sel_n0"][1] >= 0).sum() == g["indices"].shape[1] - 1
    assert (g["sel_n0"][2] >= 0).sum() == g["indices"].shape[1] - 1


def test_bilininterp_precomputedbins_matches_compiled_reference():
    """utils_cy.pyx:25-50 restated in NumPy against the compiled reference on the shapes of the
    reference's own tests/test_utils.py:89-115 (skipped where oracle/_ref is not built)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    ref.load()
    from delight.utils_cy import bilininterp_precomputedbins as cy, find_positions
    rng = np.random.default_rng(0)
    numBands, nobj, nz1, nz2 = 3, 10, 40, 50
    grid1, grid2 = np.logspace(0., 1., nz1), np.linspace(1., 10., nz2)
    v1s, v2s = rng.uniform(1, 10, nobj), rng.uniform(1, 10, nobj)
    p1s, p2s = np.zeros(nobj, dtype=np.int64), np.zeros(nobj, dtype=np.int64)
    find_positions(nobj, nz1, v1s, p1s, grid1)
    find_positions(nobj, nz2, v2s, p2s, grid2)
    Kgrid = np.stack([(grid1[:, None] * grid2[None, :]) ** (b + 1.) for b in range(numBands)])
    want, got = np.zeros((numBands, nobj)), np.zeros((numBands, nobj))
    cy(numBands, nobj, want, v1s, v2s, p1s, p2s, grid1, grid2, Kgrid)
    orc.bilininterp_precomputedbins(numBands, nobj, got, v1s, v2s, p1s, p2s, grid1, grid2, Kgrid)
    np.testing.assert_allclose(got, want, rtol=1e-13)
    # and it interpolates: exact on a bilinear function
    Kb = (grid1[:, None] + 2 * grid2[None, :])[None]
    out = np.zeros((1, nobj))
    orc.bilininterp_precomputedbins(1, nobj, out, v1s, v2s, p1s, p2s, grid1, grid2, Kb)
    np.testing.assert_allclose(out[0], v1s + 2 * v2s, rtol=1e-13)

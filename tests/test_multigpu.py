"""Multi-GPU test of the file-level drivers over NCCL (needs >= 2 GPUs; skipped otherwise):
`torchrun --nproc-per-node 2 -m delight_b200.pipeline learn|apply` must write the same files as
the single-process run -- uneven rank split, a skipped training galaxy (whose all-zero gpparams
row then sits in the MIDDLE of the gathered file) and the compress-index line numbers included.
Mirrors the rank split and the Gatherv calls of scripts/delight-apply.py:50-53,191-218 and
scripts/delight-learn.py:31-33,125."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(n, cmd, cfg):
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    if n == 1:
        argv = [sys.executable, "-m", "delight_b200.pipeline", cmd, cfg]
    else:
        argv = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                "--master-port", str(_free_port()), "-m", "delight_b200.pipeline", cmd, cfg]
    out = subprocess.run(argv, env=env, cwd=REPO, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-4000:]


def _workdir(wd, g):
    from oracle import pipeline as refpipe
    K = int(g["Ncompress"])
    cfg = refpipe.make_workdir(wd, bands=refpipe.SDSS,
                               train_bands=["U_SDSS", "G_SDSS", "I_SDSS"], refband="I_SDSS",
                               numObjects=24, dz=0.01, nzGP=40, Ncompress=K, copy_data=False)
    refpipe.write_preprocessed(wd, refpipe.SDSS, g["amp"], g["mu"], g["sig"], g["tmpl"])
    tr = g["train_cat"].copy()
    tr[4, 3] = -1.0                    # reference-band flux invalid: galaxy 4 is skipped (io.py:343)
    np.savetxt(wd + "/galaxies-fluxredshifts.txt", tr)
    np.savetxt(wd + "/galaxies-fluxredshifts2.txt", g["target_cat"][:23])      # 11 + 12 split
    return cfg


def test_pipeline_two_ranks_write_the_same_files_as_one(golden, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = golden("sdss_pipeline")
    out = {}
    for n in (1, 2):
        wd = str(tmp_path / ("ranks%d" % n))
        cfg = _workdir(wd, g)
        _run(n, "learn", cfg)
        _run(n, "apply", cfg)
        _run(n, "templatefit", cfg)
        out[n] = {k: np.loadtxt(os.path.join(wd, f), ndmin=2) for k, f in (
            ("rows", "galaxies-gpparams.txt"), ("pdf", "redshiftpdfs.txt"),
            ("metrics", "metrics.txt"), ("idx", "compressIndices.txt"),
            ("ev", "compressLikes.txt"), ("pdf_cww", "redshiftpdfs-cww.txt"),
            ("metrics_cww", "metrics-cww.txt"))}
    a, b = out[1], out[2]
    nT = a["rows"].shape[0]
    # learn: the same fitted rows; the skipped galaxy's zero row closes rank 0's block
    # (delight-learn.py: `loc` only advances on yielded objects) instead of the whole file
    valid1, valid2 = a["rows"][:, 0] >= 1, b["rows"][:, 0] >= 1
    assert valid1.sum() == valid2.sum() == nT - 1
    assert not valid1[-1] and not valid2[nT // 2 - 1] and valid2[-1]
    np.testing.assert_array_equal(a["rows"][valid1], b["rows"][valid2])
    # apply: PDFs / metrics / evidences are written with %.2e; the compress indices are line
    # numbers of each run's own gpparams file and must point at the same training galaxies
    np.testing.assert_allclose(b["pdf"], a["pdf"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(b["metrics"], a["metrics"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(b["ev"], a["ev"], rtol=1e-12, atol=0)
    ia, ib = a["idx"].astype(int), b["idx"].astype(int)
    assert valid1[ia].all() and valid2[ib].all()
    np.testing.assert_array_equal(a["rows"][ia][..., 1:3], b["rows"][ib][..., 1:3])   # (z, ell)
    np.testing.assert_allclose(b["pdf_cww"], a["pdf_cww"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(b["metrics_cww"], a["metrics_cww"], rtol=1e-12, atol=0)
    assert a["pdf"].shape[0] == 23 and np.any(a["pdf"] > 0)

"""world_size-2 gloo test (CPU) of the multi-GPU host logic: reference rank partition + Gatherv
semantics of delight_b200.distributed.gather_rows (no CUDA needed)."""
import os
import socket
import subprocess
import sys
import textwrap

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch
    from delight_b200 import distributed as D
    rank, size = D.init("gloo")
    assert size == 2
    for n in (11, 8, 1):
        first, last = D.partition(n, rank, size)
        full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
        idx = torch.arange(n, dtype=torch.int64).reshape(n, 1) * 7
        out = D.gather_rows(full[first:last].clone(), n)
        outi = D.gather_rows(idx[first:last].clone(), n)
        if rank == 0:
            assert torch.equal(out, full), (n, out)
            assert torch.equal(outi, idx) and outi.dtype == torch.int64
        else:
            assert out is None and outi is None
    assert D.counts(11, 2) == [5, 6] and sum(D.counts(12345, 8)) == 12345
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gather_rows_world2_gloo(tmp_path):
    port = _free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % rank in o


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """bench.py --impl reference launched the way the driver launches the scaling run (torchrun,
    2 ranks, no GPU needed): rank 0 alone runs the compiled reference on the host cores and prints
    ONE JSON line; the other rank exits 0 without work.  Reduced workload (300 training galaxies,
    16 targets) so that it takes seconds."""
    import json
    import pytest
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
         os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
         "--warmup", "0", "--train", "300", "--targets", "16"],
        env=env, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1
    assert d["metric"] == "galaxy-PDFs/sec" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0

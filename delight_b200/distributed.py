"""Multi-GPU plumbing: one process per GPU (torchrun), targets sharded by contiguous rank ranges
with the reference's formula, model grid replicated, results gathered to rank 0.

Mirrors the (commented-out) mpi4py calls of scripts/delight-apply.py:50-53 (partition) and
:191-218 (Gatherv of PDFs, compression indices / evidences, metrics).  The hot path has no
collective; `gather_rows` is the only communication and works on any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests).
"""
import os

from .apply import partition  # noqa: F401  (re-export)


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op when WORLD_SIZE=1)."""
    import torch
    import torch.distributed as dist
    rank, size = world()
    if size == 1 or dist.is_initialized():
        return rank, size
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, **kw)
    return rank, size


def counts(n, size):
    """Rows per rank for n objects (sendcounts of delight-apply.py:191-198, in rows)."""
    return [partition(n, r, size)[1] - partition(n, r, size)[0] for r in range(size)]


def gather_rows(local, n_total, dst=0):
    """Gatherv of row blocks: `local` [n_rank, ...] on every rank -> [n_total, ...] on `dst`
    (None elsewhere), rank blocks in rank order = catalogue order.  Variable counts are handled
    by padding to the largest block (one all_gather-free `gather` per output, as the reference
    does one Gatherv per output)."""
    import torch
    import torch.distributed as dist
    rank, size = world()
    if size == 1 or not dist.is_initialized():
        return local
    cnt = counts(n_total, size)
    if local.shape[0] != cnt[rank]:
        raise ValueError("rank %d holds %d rows, partition says %d" % (rank, local.shape[0],
                                                                      cnt[rank]))
    mx = max(cnt)
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(size)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, cnt)], dim=0)

"""Multi-GPU: one process per GPU (torchrun), independent work items sharded over ranks, ONE
all-reduce of the per-rank gradient sums per gradient.  No particle array ever crosses
GPUs: a work item's scan / resample stays on one device.

The reference has no distributed code; its only data-parallel structure is the Python loop over
subsequences (sgmcmc_sampler.py:411-418), sequences (:1264-1277) and independent chains.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(n_items, r=None, w=None):
    """Contiguous, balanced block [lo, hi) of rank r among w ranks (fixed item -> rank map, so Philox
    streams keyed by the GLOBAL item index are independent of the number of GPUs)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(int(n_items), w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def allreduce_sum(values):
    """Sum a small float64 vector over ranks (the single collective of a gradient evaluation)."""
    values = np.asarray(values, dtype=np.float64)
    if world_size() == 1:
        return values
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(values.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(value):
    if world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if world_size() > 1:
        dist.barrier()

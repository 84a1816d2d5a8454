"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards along the batch: every normalisation is per-sample, so each rank runs the
whole step on its B/N samples and the only exchange is one mean all-reduce of the flat gradient buffer.  The
per-batch supervision mask of train_semi.py:203 is a HOST random draw: every rank must take the same branch
(otherwise the Res/Cls gradient buckets would be skipped on some ranks only), so the mask stream is seeded
identically everywhere.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns
    (rank, local_rank, world_size, group or None)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return rank, local, world, None
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world, dist.group.WORLD


def shard_range(n_items: int, rank: int, world: int):
    """Consecutive, equal shards (the global-batch gradient is the mean of per-rank means only when shards are
    equal): returns [begin, end) of this rank; n_items must divide evenly."""
    if n_items % world:
        raise ValueError(f"{n_items} items do not shard evenly over {world} ranks")
    per = n_items // world
    return rank * per, (rank + 1) * per


def allreduce_mean_(flat: torch.Tensor, n_active: int, group=None):
    """In-place mean all-reduce of flat[:n_active] (the buckets whose gradients exist this step)."""
    if group is None or dist.get_world_size(group) == 1:
        return flat
    g = flat[:n_active]
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    g.mul_(1.0 / dist.get_world_size(group))
    return flat


class SupervisionMask:
    """train_semi.py:203 -- ``mask = 0 if np.random.randn(1) > rate else 1`` -- as a seeded stream that is
    identical on every rank."""

    def __init__(self, rate: float, seed: int = 1234):
        self.rate = rate
        self.rng = np.random.RandomState(seed)

    def __call__(self) -> int:
        return 0 if self.rng.randn(1)[0] > self.rate else 1


def broadcast_parameters(modules, group=None, src=0):
    """Make every rank start from rank-src's parameters (the reference initialises one process only)."""
    if group is None:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)

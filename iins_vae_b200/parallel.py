"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards along the batch: every normalisation is per-sample, so each rank runs the
whole step on its B/N samples and the only exchange is one mean all-reduce of the flat gradient buffer.  The
per-batch supervision mask of train_semi.py:203 is a HOST random draw: every rank must take the same branch
(otherwise the Res/Cls gradient buckets would be skipped on some ranks only), so the mask stream is seeded
identically everywhere.
"""
import os
import sys
import threading

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
    """In-place mean all-reduce of flat[:n_active] (the buckets whose gradients exist this step).  Host-side helper for
    callers that own their optimizer; the fused engine instead SUM-reduces its buckets on a communication stream next to
    the encoder backward and lets iins_adam_step apply the 1/world factor (engine.SemiTrainEngine._allreduce)."""
    if group is None or dist.get_world_size(group) == 1:
        return flat
    g = flat[:n_active]
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    g.mul_(1.0 / dist.get_world_size(group))
    return flat


def shutdown_distributed(engines=(), timeout_s: float = 30.0):
    """Orderly teardown of a data-parallel run.  CUDA graphs that captured NCCL kernels keep the communicator referenced:
    the engines drop their graphs first, every rank synchronises and meets at a barrier, then the process group is
    destroyed.  A watchdog ends the process if the destroy call does not return (observed in round 1 when the graphs
    were still alive) so that a benchmark or training run can never hang at exit."""
    if not dist.is_initialized():
        return
    for e in engines:
        e.close()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier()
    done = threading.Event()

    def _watchdog():
        if not done.wait(timeout_s):
            sys.stdout.flush()
            sys.stderr.flush()
            sys.stderr.write("shutdown_distributed: destroy_process_group did not return, leaving\n")
            os._exit(0)

    threading.Thread(target=_watchdog, daemon=True).start()
    dist.destroy_process_group()
    done.set()


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

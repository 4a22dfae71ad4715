"""Multi-GPU sharding of a population: one process per GPU, contiguous row slices
("islands"), no data-path collective except ONE all-gather of the fitness values (NCCL over
NVLink / NVSwitch on GPUs, gloo in the CPU tests) and a champion = argmin on every rank.

Without an initialised torch.distributed process group everything degenerates to a single shard.
"""
from __future__ import annotations

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except ImportError:   # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_bounds(n: int, rank: int | None = None, world_size: int | None = None):
    """Rows [lo, hi) of an n-row population owned by `rank`: contiguous, sizes differ by at most one,
    the first n % world_size ranks take the extra row (the split np.array_split makes)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_population_fitness(local_fitness, n: int, ctx=None, local_flag: bool = False, return_flag: bool = False):
    """All-gather of the per-rank fitness slices into the full length-n vector (identical on every rank).
    `local_flag` travels in the same message (one extra element per rank); with `return_flag` the call returns
    (fitness, any rank's flag) so that a failure seen by one rank is raised by ALL of them after the collective
    instead of leaving the others waiting in it."""
    d = _dist()
    if d is None:
        out = np.asarray(local_fitness, dtype=np.float64)
        return (out, bool(local_flag)) if return_flag else out
    import torch
    rank, w = d.get_rank(), d.get_world_size()
    nccl = d.get_backend() == "nccl"
    dev = ctx.device if (nccl and ctx is not None) else torch.device("cpu")
    cap = -(-n // w)   # equal-sized send buffers: pad the short shards
    send = torch.full((cap + 1,), float("nan"), dtype=torch.float64, device=dev)
    lo, hi = shard_bounds(n, rank, w)
    if hi > lo:
        send[:hi - lo] = torch.as_tensor(np.asarray(local_fitness, dtype=np.float64)).to(dev)
    send[cap] = 1.0 if local_flag else 0.0
    recv = torch.empty((cap + 1) * w, dtype=torch.float64, device=dev)
    d.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(w, cap + 1)
    out = np.empty(n)
    for r in range(w):
        a, b = shard_bounds(n, r, w)
        out[a:b] = recv[r, :b - a]
    return (out, bool((recv[:, cap] != 0.0).any())) if return_flag else out


def champion(fitness: np.ndarray):
    """(index, value) of the best individual; ties resolve to the lowest index on every rank."""
    i = int(np.argmin(fitness))
    return i, float(fitness[i])

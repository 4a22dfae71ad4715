"""One population over several GPUs: one process per GPU, every rank evaluates the rows DEALT to it.

The reference spreads `udp.fitness` calls over a process pool (pg.mp_bfe, toss.py:110) and its scaling script shrinks
each worker's share as workers are added (scripts/scaling_performance_benchmark.py:126-137).  Trajectory lengths
spread 4x, so contiguous row slices leave GPUs with unequal work (and each with a long tail).  Instead:

  1. every rank predicts the cost of a contiguous slice of the rows (toss_predict_cost: a short proxy integration),
  2. ONE all-gather of the int32 costs (NCCL over NVLink; 128 KB at pop 32768),
  3. every rank derives the same longest-first order (toss_cost_order: a pure function of the costs) and takes
     positions rank, rank + world, rank + 2 world, ... of it (toss_deal_rows) -- equal predicted work, and the
     shard is already in longest-first order for its own work queue,
  4. the shard is evaluated (toss_population_fitness, queue in row order),
  5. ONE all-gather of the fitness shards (+ two failure counts riding in the same message), un-dealt on the device
     into population order (toss_unpack_result).  Every rank ends with the full fitness vector; champion = argmin.

The path has no other exchange step, so there is no other collective.  With the gloo backend (CPU tests of the host
logic) steps 2, 3 and 5 run on numpy mirrors of the same index arithmetic.  Without an initialised process group
everything degenerates to a single shard.
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
    """Rows [lo, hi) of the contiguous slice whose COST rank `rank` predicts: ceil(n / world) rows each, the last
    slices shorter or empty -- so the all-gathered slices, back to back, are the cost array itself."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    cap = -(-n // world_size)
    lo = min(n, rank * cap)
    return lo, min(n, lo + cap)


def dealt_count(n: int, rank: int, world_size: int) -> int:
    """rows of an n-row population dealt to `rank`"""
    return max(0, (n - rank + world_size - 1) // world_size)


def cost_order_host(cost: np.ndarray) -> np.ndarray:
    """numpy mirror of toss_cost_order (csrc/stepper.cu cost_order_kernel): bins of cost >> 2 clamped to [0, 4095],
    longest first, ties in index order."""
    key = np.clip(np.asarray(cost, dtype=np.int64) >> 2, 0, 4095)
    return np.argsort(-key, kind="stable").astype(np.int32)


def deal_host(order: np.ndarray, rank: int, world_size: int) -> np.ndarray:
    """population row indices dealt to `rank`, in its queue order (mirror of toss_deal_rows)"""
    return np.asarray(order)[rank::world_size]


def pack_host(fitness_local, status_local, cap: int) -> np.ndarray:
    """mirror of toss_pack_result"""
    send = np.full(cap + 2, np.nan)
    fitness_local = np.asarray(fitness_local, dtype=np.float64).reshape(-1)
    send[:len(fitness_local)] = fitness_local
    st = np.asarray(status_local).reshape(-1) if status_local is not None else np.zeros(0, dtype=np.int32)
    send[cap] = float((st == 3).sum())
    send[cap + 1] = float(((st == 1) | (st == 2)).sum())
    return send


def unpack_host(recv: np.ndarray, order: np.ndarray, n: int, world_size: int, cap: int):
    """mirror of toss_unpack_result: (fitness in population order, (n_short, n_failed))"""
    recv = np.asarray(recv).reshape(world_size, cap + 2)
    pos = np.arange(n)
    fit = np.empty(n)
    fit[np.asarray(order)[:n]] = recv[pos % world_size, pos // world_size]
    return fit, (int(recv[:, cap].sum()), int(recv[:, cap + 1].sum()))


def exchange_dealt(fitness_local, status_local, order, n: int):
    """Step 5 on HOST arrays over the active process group (gloo in the CPU tests): every rank passes the fitness of
    the rows dealt to it (in dealt order); returns (full fitness vector, (n_short, n_failed)) on every rank."""
    d = _dist()
    rank, w = world()
    cap = -(-n // w)
    send = pack_host(fitness_local, status_local, cap)
    if d is None:
        return unpack_host(send, order, n, 1, cap)
    import torch
    recv = torch.empty((cap + 2) * w, dtype=torch.float64)
    d.all_gather_into_tensor(recv, torch.from_numpy(send))
    return unpack_host(recv.numpy(), order, n, w, cap)


def gather_costs_host(cost_local, n: int) -> np.ndarray:
    """Step 2 on HOST arrays: all-gather of the contiguous cost slices -> the cost array (n)"""
    d = _dist()
    rank, w = world()
    if d is None:
        return np.asarray(cost_local, dtype=np.int32)
    import torch
    cap = -(-n // w)
    send = torch.zeros(cap, dtype=torch.int32)
    send[:len(cost_local)] = torch.as_tensor(np.asarray(cost_local, dtype=np.int32))
    recv = torch.empty(cap * w, dtype=torch.int32)
    d.all_gather_into_tensor(recv, send)
    return recv.numpy()[:n].copy()


class ShardedResult(dict):
    """fitness (n,) in population order on every rank + what this rank computed for its own shard"""


def sharded_population_fitness(ctx, params, xs, fixed, n_man: int, knot_cap: int):
    """Steps 1-5 on the device over the NCCL process group.  `xs` is the WHOLE population: a CUDA tensor (n, dim)
    (device-resident inputs) or a host array (copied H2D here).  Returns a ShardedResult with CUDA tensors:
    fitness (n) in population order, flags (2) = [#knot-capacity failures, #integrator failures] over all ranks,
    order (n), local = the dict toss_population_fitness returned for this rank's shard (dealt order)."""
    import torch
    import torch.distributed as dist
    rank, w = world()
    t = ctx.torch
    xs = ctx._dev(xs)
    n, dim = xs.shape
    cap = -(-n // w)
    with torch.cuda.nvtx.range("toss:predict cost + all-gather"):
        lo, hi = shard_bounds(n, rank, w)
        cost_send = t.zeros(cap, dtype=t.int32, device=ctx.device)
        ctx.predict_cost(params, xs[lo:hi], fixed, n_man, out=cost_send)
        cost_all = t.empty(cap * w, dtype=t.int32, device=ctx.device)
        dist.all_gather_into_tensor(cost_all, cost_send)
        order = ctx.cost_order(cost_all[:n])
    x_local = ctx.deal_rows(xs, order, rank, w)
    n_local = x_local.shape[0]
    local = None
    if n_local:
        ctx.set_queue_order(True)
        try:
            local = ctx.population_fitness(params, x_local, fixed, n_man, knot_cap)
        finally:
            ctx.set_queue_order(False)
    with torch.cuda.nvtx.range("toss:fitness all-gather + un-deal"):
        send = ctx.pack_result(local["fitness"] if local else None, local["counters"] if local else None, n_local, cap)
        recv = t.empty((cap + 2) * w, dtype=t.float64, device=ctx.device)
        dist.all_gather_into_tensor(recv, send)
        fitness, flags = ctx.unpack_result(recv, order, n, w, cap)
    return ShardedResult(fitness=fitness, flags=flags, order=order, local=local, n_local=n_local)


def champion(fitness: np.ndarray):
    """(index, value) of the best individual; ties resolve to the lowest index on every rank."""
    i = int(np.argmin(fitness))
    return i, float(fitness[i])

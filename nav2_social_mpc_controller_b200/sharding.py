"""Multi-GPU sharding of a problem batch: one process per GPU, contiguous index ranges, no collective on the
solve path — only a final gather of the per-problem results. For multi-start batches the granule is
the number of starts per robot, so a robot's arg-min never crosses a rank."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_problems: int, world: int, rank: int, granule: int = 1):
    """[lo, hi) of `rank`: contiguous, multiples of `granule`, sizes differing by at most one granule."""
    if n_problems % granule:
        raise ValueError("n_problems must be a multiple of granule")
    units = n_problems // granule
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return lo_u * granule, hi_u * granule


def tiled_source_index(total: int, unique: int, world: int, rank: int) -> np.ndarray:
    """Strong-scaling batches that are larger than their set of generated scenarios (bench.py `crowd_x1M_A50`: 10^6
    problems from 16384 unique ones): problem g of the global batch is unique scenario g % unique, and rank r owns the
    contiguous range [r * (total // world), (r + 1) * (total // world)). Returns the unique-scenario index of every
    problem of `rank`, in order."""
    per_rank = total // world
    return (rank * per_rank + np.arange(per_rank, dtype=np.int64)) % unique


def gather_results(local: dict, dst: int = 0):
    """Concatenate the per-rank result dicts on `dst` in rank order (torch.distributed must be initialised;
    works with gloo on CPU and nccl on GPU since results are host numpy arrays gathered as objects)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst)
    if rank != dst:
        return None
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def solve_sharded(solve_fn, batch, granule: int = 1, dst: int = 0):
    """Strong-scaling helper: every rank solves its slice of `batch` with `solve_fn(sub_batch) -> dict` and the
    results are gathered on `dst`. `solve_fn` is Optimizer.solve_batch on a GPU rank."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(batch.n_problems, world, rank, granule)
    local = solve_fn(batch.slice(lo, hi))
    return gather_results(local, dst)

"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding + final gather reproduces the
single-process result. The per-shard solver here is the CPU oracle (tests may use it); on GPU ranks it is
Optimizer.solve_batch."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from nav2_social_mpc_controller_b200 import scenarios as sc
from nav2_social_mpc_controller_b200.sharding import shard_bounds


def test_shard_bounds_cover_and_align():
    for n, world, gran in [(4096, 8, 1), (10, 3, 1), (256 * 1024, 8, 1024), (7 * 64, 4, 64), (5, 8, 1)]:
        prev = 0
        sizes = []
        for r in range(world):
            lo, hi = shard_bounds(n, world, r, gran)
            assert lo == prev and lo % gran == 0 and hi % gran == 0
            sizes.append(hi - lo)
            prev = hi
        assert prev == n and max(sizes) - min(sizes) <= gran
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 0, 4)


def _worker(rank, world, port, q, shared=False):
    import torch.distributed as dist
    from nav2_social_mpc_controller_b200.sharding import solve_sharded
    from tests import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle_lib.load()
    batch = sc.multistart(n_robots=6, n_starts=4, shared=shared)  # shared: one scene row per robot (scenario_index)

    def solve(sub):
        r = o.solve_batch(sub, want=("u", "cost_final", "usable", "termination"))
        return {k: v[: sub.n_problems] for k, v in r.items()}
    out = solve_sharded(solve, batch, granule=4)
    if rank == 0:
        q.put({k: v.tolist() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shared", [False, True])
def test_two_rank_sharded_solve_matches_single_process(oracle, shared):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, shared)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = sc.multistart(n_robots=6, n_starts=4)
    ref = oracle.solve_batch(batch, want=("u", "cost_final", "usable", "termination"))
    for k in ref:
        assert np.array_equal(np.array(got[k]), ref[k]), k


def test_tiled_strong_scaling_partition():
    """bench.py `crowd_x1M_A50`: 10^6 problems tiled from 16384 unique scenarios, total / world per rank."""
    from nav2_social_mpc_controller_b200.sharding import tiled_source_index
    total, unique = 1_000_000, 16384
    for world in (1, 2, 4, 8):
        parts = [tiled_source_index(total, unique, world, r) for r in range(world)]
        assert all(len(p) == total // world for p in parts)
        whole = np.concatenate(parts)
        assert np.array_equal(whole, np.arange(total) % unique)          # global problem g = scenario g % unique
        assert parts[0][0] == 0 and whole.max() == unique - 1

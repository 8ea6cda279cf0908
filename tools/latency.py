#!/usr/bin/env python
"""Single-solve latency (BASELINE config 1): p50 / p99 of smpc_solve_batch (B = 1, host buffers, H2D + kernel + D2H)
and of the level-2 smpc_optimize call, for the README-example and params.yaml parameter sets."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav2_social_mpc_controller_b200 import scenarios as sc  # noqa: E402
from nav2_social_mpc_controller_b200.optimizer import Optimizer  # noqa: E402

out = {}
for name in ("readme", "params_yaml", "soc_work_obst"):
    batch = sc.single(name)
    opt = Optimizer(0)
    opt.initialize(batch.params)
    for _ in range(50):
        r = opt.solve_batch(batch)
    lat = []
    for _ in range(1000):
        t0 = time.perf_counter()
        r = opt.solve_batch(batch)
        lat.append(time.perf_counter() - t0)
    lat = np.sort(np.array(lat)) * 1e3
    out[name] = dict(S=batch.n_steps, P=2 * batch.n_blocks, p50_ms=float(lat[500]), p99_ms=float(lat[990]),
                     iterations=int(r["iterations"][0]), evaluations=int(r["n_evals"][0, 0]),
                     termination=int(r["termination"][0]), kernel_ms=opt.last_kernel_ms())
    opt.close()

# level-2: the whole Optimizer::optimize tick for ONE robot (smpc_optimize: people_to_status, format_to_optimize, SFM people
# projection, solve, post-solve, memory update — all on the GPU; warm-started second and later ticks), 3 people
import math  # noqa: E402

for name in ("readme", "soc_work_obst"):
    p = sc.make_params(name)
    rng = np.random.default_rng(5)
    pose = np.array([[2.0, 2.05, 0.1]])
    poses, cmds = sc.pure_pursuit_seed(sc._straight_path(1, pose[:, 0], np.array([2.0])), pose, p)
    people = []
    for _ in range(3):
        r_, b_ = rng.uniform(0.9, 1.8), rng.uniform(-0.7, 0.7)
        px, py = 2.0 + r_ * math.cos(b_), 2.05 + r_ * math.sin(b_)
        hd, v = math.atan2(2.05 - py, 2.0 - px), rng.uniform(0.2, 0.9)
        people.append([px, py, v * math.cos(hd), v * math.sin(hd), 0.0])
    costmap = sc.wall_costmap(80, 80, 0.05, walls_y=(0.6, 3.4))
    rows = np.arange(80)[:, None] * np.ones((1, 80), dtype=int)
    cols = np.ones((80, 1), dtype=int) * np.arange(80)[None, :]
    near = np.where(np.abs(rows - 12) <= np.abs(rows - 68), 12, 68)
    od = dict(width=80, height=80, resolution=0.05, origin_x=0.0, origin_y=0.0,
              distances=(np.abs(rows - near) * 0.05).astype(np.float32).ravel(),
              indexes=(near * 80 + cols).astype(np.uint32).ravel())
    opt = Optimizer(0)
    opt.initialize(p)
    args = (poses[0], cmds[0], np.array(people), (0.3, 0.05), p.time_step, costmap, (0.0, 0.0), 0.05, od)
    for _ in range(30):
        ok, *_rest, info = opt.optimize(*args)
    lat = []
    for _ in range(500):
        t0 = time.perf_counter()
        ok, *_rest, info = opt.optimize(*args)
        lat.append(time.perf_counter() - t0)
    lat = np.sort(np.array(lat)) * 1e3
    out["optimize_" + name] = dict(what="smpc_optimize: one robot, 3 people, the whole Optimizer::optimize tick", optimized=bool(ok),
                                   p50_ms=float(lat[250]), p99_ms=float(lat[495]), iterations=int(info["iterations"]))
    opt.close()
print(json.dumps(out))

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
print(json.dumps(out))

#!/usr/bin/env python
"""Generate tests/golden/solve_cases.npz: small fixed inputs of the hot path (level-1 batches in the include/smpc.h
layout) together with the CPU oracle's outputs for them, so that both the oracle (CPU tests) and the CUDA path (GPU
tests) are checked against vectors that do not move when either side is edited.

The reference itself ships no tests or golden vectors and cannot be built here (Ceres / Eigen / ROS 2 absent), so these
vectors pin the ORACLE'S behaviour, not Ceres': parity stays "unpinned" in the sense of DESIGN.md §4.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nav2_social_mpc_controller_b200 import scenarios as sc  # noqa: E402
from tests import oracle_lib  # noqa: E402

CASES = {
    "single_readme_A3": lambda: sc.single("readme", n_people=3),
    "single_params_yaml_A3": lambda: sc.single("params_yaml", n_people=3),
    "single_soc_work_obst_A2": lambda: sc.single("soc_work_obst", n_people=2),
    "corridor_x8": lambda: sc.corridor(B=8),
    "crowd_x8_A3": lambda: sc.crowd(B=8, A=3, config_id=6, n_maps=4),
    "crowd_x4_A20": lambda: sc.crowd(B=4, A=20, n_maps=2),
    # Ceres >= 2.1 behaviour (std::numeric_limits<Jet> specialised: ProxemicsCost has its true value and gradient in
    # differentiated evaluations; tolerance tests armed after the first successful step)
    "single_readme_A3_ceres220": lambda: sc.single("readme", n_people=3, ceres_compat=220),
    "crowd_x8_A3_ceres220": lambda: sc.crowd(B=8, A=3, config_id=6, n_maps=4, ceres_compat=220),
    "crowd_x4_A20_ceres220": lambda: sc.crowd(B=4, A=20, n_maps=2, ceres_compat=220),
    # per-problem horizons (n_steps_each): fewer steps, shorter control horizon, fewer parameter blocks per problem
    "mixed_horizon_crowd_x8_A3": lambda: sc.with_horizons(sc.crowd(B=8, A=3, config_id=6, n_maps=4),
                                                          [28, 20, 13, 7, 5, 18, 2, 1]),
    "mixed_horizon_corridor_x8": lambda: sc.with_horizons(sc.corridor(B=8), [28, 27, 12, 6, 3, 19, 11, 24]),
}
OUT_KEYS = ("u", "cmds", "path", "cost_initial", "cost_final", "iterations", "termination", "usable")


def main():
    oracle = oracle_lib.load()
    blob = {}
    for name, make in CASES.items():
        b = make()
        blob[f"{name}/params"] = np.frombuffer(bytes(b.params), dtype=np.uint8).copy()
        blob[f"{name}/meta"] = np.array([b.n_problems, b.n_steps, b.n_agents, b.n_costmaps, b.size_x, b.size_y],
                                        dtype=np.int64)
        blob[f"{name}/scalars"] = np.array([b.resolution, b.dt], dtype=np.float64)
        for k, v in b.arrays.items():
            if v is not None:
                blob[f"{name}/in/{k}"] = v
        out = oracle.solve_batch(b, want=OUT_KEYS)
        for k in OUT_KEYS:
            blob[f"{name}/out/{k}"] = out[k]
        ev = [oracle.evaluate(b, i, b.arrays["u0"][i]) for i in range(b.n_problems)]
        blob[f"{name}/eval/cost"] = np.array([e["cost"] for e in ev])
        P = 2 * b.n_blocks  # a problem with a shorter horizon has fewer parameters: its row is zero-padded
        blob[f"{name}/eval/grad"] = np.array([np.pad(e["grad"], (0, P - e["grad"].size)) for e in ev])
        print(name, "iterations", out["iterations"].tolist(), "termination", out["termination"].tolist())
    path = os.path.join(ROOT, "tests", "golden", "solve_cases.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

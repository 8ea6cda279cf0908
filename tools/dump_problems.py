#!/usr/bin/env python
"""Companion of oracle/ceres_harness.cpp (a machine with the real Ceres): dump the level-1 inputs of a golden case in
the harness's binary format, and compare the harness output (the reference's own functors under ceres::Solve) with the
CPU oracle — the step that would turn "parity unpinned" into pinned.

  python tools/dump_problems.py --case crowd_x8_A3 --out /tmp/p.bin
  oracle/_ref/ceres_harness /tmp/p.bin > /tmp/ceres.jsonl
  python tools/dump_problems.py --case crowd_x8_A3 --compare /tmp/ceres.jsonl
"""
import argparse
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_lib  # noqa: E402


def dump(batch, path):
    p = batch.params
    assert batch.arrays.get("n_steps_each") is None, "the harness takes uniform horizons"
    B, S, A = batch.n_problems, batch.n_steps, batch.n_agents
    a = batch.arrays
    with open(path, "wb") as f:
        f.write(struct.pack("<8i", B, S, A, batch.size_x, batch.size_y, p.control_horizon, p.parameter_block_length,
                            p.max_iterations))
        f.write(struct.pack("<5d", batch.resolution, batch.dt, p.fn_tol, p.gradient_tol, p.param_tol))
        f.write(struct.pack("<9d", p.distance_w, p.socialwork_w, p.velocity_w, p.angle_w, p.agent_angle_w, p.proxemics_w,
                            p.velocity_feasibility_w, p.obstacle_w, p.goal_align_w))
        for b in range(B):
            mi = int(a["costmap_index"][b]) if a.get("costmap_index") is not None else b % batch.n_costmaps
            f.write(np.ascontiguousarray(a["pose0"][b], dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(a["u0"][b], dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(a["path_xy"][b], dtype="<f8").tobytes())
            f.write(struct.pack("<d", float(a["goal_yaw"][b])))
            if A > 0:
                f.write(np.ascontiguousarray(a["agents"][b], dtype="<f8").tobytes())
            f.write(struct.pack("<B", int(a["has_people"][b])))
            f.write(np.ascontiguousarray(a["costmap_origin"][mi], dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(a["costmaps"][mi], dtype=np.uint8).tobytes())


def compare(batch, gold, jsonl):
    rows = [json.loads(line) for line in open(jsonl) if line.strip()]
    print("Ceres", rows[0]["ceres_version"], "vs oracle (ceres_compat", batch.params.ceres_compat, ")")
    ok = 0
    for r in rows:
        b = r["problem"]
        u = np.array(r["u"]).reshape(-1, 2)
        du = np.abs(u - gold["u"][b]).max()
        dc = abs(r["final_cost"] - gold["cost_final"][b]) / max(abs(gold["cost_final"][b]), 1e-300)
        same = r["iterations"] == int(gold["iterations"][b]) and bool(r["usable"]) == bool(gold["usable"][b])
        good = du <= 1e-6 and dc <= 1e-8 and same
        ok += good
        print(f"problem {b}: du {du:.3e} dcost {dc:.3e} iterations {r['iterations']}/{int(gold['iterations'][b])} "
              f"{'OK' if good else 'DIFFERENT'}")
    print(f"{ok} / {len(rows)} within the north-star tolerance with identical iteration counts")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="crowd_x8_A3")
    ap.add_argument("--out")
    ap.add_argument("--compare")
    args = ap.parse_args()
    batch, gold, _ = golden_lib.load()[args.case]
    if args.out:
        dump(batch, args.out)
        print("wrote", args.out, os.path.getsize(args.out), "bytes")
    if args.compare:
        compare(batch, gold, args.compare)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Timing of the pre-solve / fleet kernels, on one GPU, CUDA events on the launching stream:
  project_people (SFM crowd projection) at A = 3 / 20 / 50, trajectorize (seed generation), and the whole fleet tick
  smpc_optimize_batch (host buffers in / out, maps resident after the first tick).
Prints one JSON document; tools/final_run.sh stores it under profiles/."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nav2_social_mpc_controller_b200 import _lib, abi, scenarios as sc  # noqa: E402
from nav2_social_mpc_controller_b200.fleet import FleetOptimizer  # noqa: E402
from nav2_social_mpc_controller_b200.optimizer import Optimizer  # noqa: E402


def od_grid(W=80, H=80, res=0.05):
    rows = np.arange(H)[:, None] * np.ones((1, W), dtype=int)
    cols = np.ones((H, 1), dtype=int) * np.arange(W)[None, :]
    near = np.where(np.abs(rows - 12) <= np.abs(rows - 68), 12, 68)
    return (near * W + cols).astype(np.uint32).ravel()


def events(fn, stream, reps=10, warm=3):
    for _ in range(warm):
        fn()
    stream.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(stream)
        fn()
        b.record(stream)
    stream.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    p = sc.make_params("soc_work_obst")
    opt = Optimizer(0)
    opt.initialize(p)
    L, h = _lib.lib(), opt._h
    out = {"gpu": torch.cuda.get_device_name(0)}
    S = 28
    rng = np.random.default_rng(3)
    # ---- project_people
    proj = {}
    for A, B in ((3, 65536), (20, 65536), (50, 16384)):
        robot = np.zeros((B, S + 1, 6))
        robot[:, :, 0] = 0.8 + 0.03 * np.arange(S + 1)
        robot[:, :, 1] = 2.0 + rng.uniform(-0.3, 0.3, (B, 1))
        robot[:, :, 4] = 0.6
        init = np.zeros((B, A, 6))
        init[:, :, 0] = rng.uniform(0.6, 3.4, (B, A))
        init[:, :, 1] = rng.uniform(1.0, 3.0, (B, A))
        init[:, :, 2] = rng.uniform(-np.pi, np.pi, (B, A))
        init[:, :, 4] = rng.uniform(0.0, 0.6, (B, A))
        d_robot, d_init = torch.from_numpy(robot).to(dev), torch.from_numpy(init).to(dev)
        d_idx = torch.from_numpy(od_grid().view(np.int32)).to(dev)
        d_org = torch.zeros(1, 2, dtype=torch.float64, device=dev)
        agents = torch.empty(B, A, 6, S + 1, dtype=torch.float64, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        a = abi.SmpcProjectArgs()
        a.n_problems, a.n_steps, a.n_agents, a.n_grids = B, S, A, 1
        a.od_width, a.od_height, a.od_resolution = 80, 80, 0.05
        a.max_time, a.time_step = float(p.max_time), float(p.time_step)
        a.od_origin, a.od_indexes, a.od_index = d_org.data_ptr(), d_idx.data_ptr(), None
        a.robot, a.people_init, a.agents, a.status = d_robot.data_ptr(), d_init.data_ptr(), agents.data_ptr(), status.data_ptr()
        ms = events(lambda: _lib.check(L.smpc_project_people_batch_device(h, C.byref(a), stream.cuda_stream)), stream)
        pair_steps = B * S * A * (A + 1)  # every person against every other agent (incl. the robot), every step
        byts = B * (A * 6 * (S + 1) + (S + 1) * 6 + A * 6) * 8
        proj[f"A{A}"] = {"problems": B, "ms": ms, "problems_per_s": B / (ms * 1e-3), "pair_interactions_per_s":
                         pair_steps / (ms * 1e-3), "algorithmic_GB_per_s": byts / (ms * 1e-3) / 1e9}
        del d_robot, d_init, agents
    out["project_people"] = proj
    # ---- trajectorize
    B = 65536
    pose = np.stack([rng.uniform(0.5, 1.0, B), 2.0 + rng.uniform(-0.4, 0.4, B), rng.uniform(-2.5, 2.5, B)], axis=1)
    gpath = sc._straight_path(B, np.full(B, 0.6), np.full(B, 2.0))
    max_steps = int(round(float(p.max_time) / round(float(p.time_step), 6)))
    d_gp, d_pose = torch.from_numpy(gpath).to(dev), torch.from_numpy(pose).to(dev)
    poses = torch.zeros(B, max_steps + 1, 3, dtype=torch.float64, device=dev)
    cmds = torch.zeros(B, max_steps, 3, dtype=torch.float64, device=dev)
    n_steps = torch.zeros(B, dtype=torch.int32, device=dev)
    t = abi.SmpcTrajectorizeArgs()
    t.n_problems, t.n_path, t.max_steps, t.omnidirectional = B, gpath.shape[1], max_steps, 0
    t.desired_linear_vel, t.lookahead_dist = p.traj_desired_linear_vel, p.lookahead_dist
    t.max_angular_vel, t.time_step = p.max_angular_vel, round(float(p.time_step), 6)
    t.global_path, t.path_index, t.pose = d_gp.data_ptr(), None, d_pose.data_ptr()
    t.poses, t.cmds, t.n_steps = poses.data_ptr(), cmds.data_ptr(), n_steps.data_ptr()
    ms = events(lambda: _lib.check(L.smpc_trajectorize_batch_device(h, C.byref(t), stream.cuda_stream)), stream)
    out["trajectorize"] = {"robots": B, "path_points": int(gpath.shape[1]), "ms": ms, "robots_per_s": B / (ms * 1e-3)}
    opt.close()
    # ---- whole fleet tick through the C entry (host buffers)
    fleet_out = {}
    for B in (256, 4096):
        scene_pose = np.stack([rng.uniform(0.5, 1.0, B), 2.0 + rng.uniform(-0.3, 0.3, B), rng.uniform(-0.3, 0.3, B)], axis=1)
        gp = sc._straight_path(B, np.full(B, 0.6), np.full(B, 2.0))
        poses_h, cmds_h = sc.pure_pursuit_seed(gp, scene_pose, p)
        people = np.zeros((B, 3, 5))
        people[:, :, 0] = rng.uniform(1.5, 3.2, (B, 3))
        people[:, :, 1] = rng.uniform(1.2, 2.8, (B, 3))
        people[:, :, 2:4] = rng.uniform(-0.5, 0.5, (B, 3, 2))
        n_people = np.full(B, 3, dtype=np.int32)
        speed = np.tile([0.3, 0.0], (B, 1))
        costmap = sc.wall_costmap(80, 80, 0.05, walls_y=(0.6, 3.4))[None]
        od = dict(width=80, height=80, resolution=0.05, origins=[[0.0, 0.0]], indexes=od_grid())
        fl = FleetOptimizer(p, n_robots=B, n_agents=3)
        for _ in range(3):
            fl.optimize_batch(poses_h, cmds_h[:, :, :2], people, n_people, speed, costmap, np.zeros((1, 2)), 0.05, od,
                              want_people_proj=False)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            r = fl.optimize_batch(poses_h, cmds_h[:, :, :2], people, n_people, speed, costmap, np.zeros((1, 2)), 0.05, od,
                                  want_people_proj=False)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        fleet_out[f"B{B}"] = {"robots": B, "tick_ms_p50": 1e3 * ts[len(ts) // 2], "robots_per_s": B / ts[len(ts) // 2],
                              "optimized_fraction": float(r["optimized"].mean()),
                              "what": "smpc_optimize_batch: host buffers in / out, warm-started second and later ticks, "
                                      "maps resident on the device"}
        fl.close()
    out["fleet_tick"] = fleet_out
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

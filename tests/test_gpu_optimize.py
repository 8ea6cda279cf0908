"""GPU tests of the level-2 entry smpc_optimize (mirror of bool Optimizer::optimize, reference
optimizer.hpp:167-170) against tests/presolve_ref.py (pre-solve stages) + the CPU oracle (solve, post-solve)."""
import math

import numpy as np
import pytest

from nav2_social_mpc_controller_b200 import scenarios as sc
from tests import presolve_ref as ref

pytestmark = pytest.mark.gpu


def _scene(param_set="soc_work_obst", n_people=2, seed=0):
    rng = np.random.default_rng(seed)
    p = sc.make_params(param_set)
    pose = np.array([[2.0, 2.0 + rng.uniform(-0.2, 0.2), rng.uniform(-0.3, 0.3)]])
    gpath = sc._straight_path(1, pose[:, 0], np.array([2.0]))
    poses, cmds = sc.pure_pursuit_seed(gpath, pose, p)
    people = []
    for _ in range(n_people):
        r, b = rng.uniform(0.9, 1.8), rng.uniform(-0.7, 0.7)
        px, py = pose[0, 0] + r * math.cos(b), pose[0, 1] + r * math.sin(b)
        h = math.atan2(pose[0, 1] - py, pose[0, 0] - px) + rng.uniform(-0.4, 0.4)
        v = rng.uniform(0.2, 0.9)
        people.append([px, py, v * math.cos(h), v * math.sin(h), 0.0])
    costmap = sc.wall_costmap(80, 80, 0.05, walls_y=(0.6, 3.4))
    # obstacle-distance grid: nearest wall cell of each cell (walls at rows 12 and 68)
    W = H = 80
    rows = np.arange(H)[:, None] * np.ones((1, W), dtype=int)
    cols = np.ones((H, 1), dtype=int) * np.arange(W)[None, :]
    near = np.where(np.abs(rows - 12) <= np.abs(rows - 68), 12, 68)
    od = dict(width=W, height=H, resolution=0.05, origin_x=0.0, origin_y=0.0,
              distances=(np.abs(rows - near) * 0.05).astype(np.float32).ravel(),
              indexes=(near * W + cols).astype(np.uint32).ravel())
    return p, poses[0], cmds[0], np.array(people).reshape(-1, 5), (0.3, 0.05), costmap, od


@pytest.fixture()
def opt_for():
    from nav2_social_mpc_controller_b200.optimizer import Optimizer
    made = []

    def _mk(params):
        o = Optimizer(0)
        o.initialize(params)
        made.append(o)
        return o
    yield _mk
    for o in made:
        o.close()


def _reference_tick(oracle, p, poses, cmds, prev, people, speed, costmap, od):
    init = ref.people_to_status(people)
    prev_poses, prev_cmds = prev if prev is not None else (poses, cmds)
    robot, blended = ref.format_to_optimize(poses, cmds, prev_poses, prev_cmds, speed, p.current_path_w,
                                            p.current_cmds_w, p.max_time, p.time_step)
    proj = ref.project_people(init, robot, od, p.max_time, p.time_step)
    batch = ref.build_level1_batch(p, robot, proj, len(people) != 0, costmap, (0.0, 0.0), 0.05, p.time_step)
    out = oracle.solve_batch(batch)
    return robot, proj, batch, out


@pytest.mark.parametrize("param_set,n_people", [("soc_work_obst", 2), ("readme", 3), ("params_yaml", 1),
                                                ("obst_only", 0), ("soc_work_obst", 5)])
def test_optimize_matches_reference_pipeline(oracle, opt_for, param_set, n_people):
    p, poses, cmds, people, speed, costmap, od = _scene(param_set, n_people, seed=n_people)
    opt = opt_for(p)
    ok, path, new_cmds, proj, info = opt.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05,
                                                  od)
    robot, proj_ref, batch, out = _reference_tick(oracle, p, poses, cmds, None, people, speed, costmap, od)
    assert proj.shape == (len(robot), 3, 6)
    assert np.allclose(proj, np.array(proj_ref), rtol=1e-10, atol=1e-10)  # SFM projection, GPU kernel vs numpy
    assert ok == bool(out["usable"][0])
    assert info["termination"] == out["termination"][0] and info["iterations"] == out["iterations"][0]
    assert info["cost_final"] == pytest.approx(out["cost_final"][0], rel=1e-8)
    assert np.abs(new_cmds - out["cmds"][0]).max() <= 1e-6
    assert np.abs(path[:, :2] - out["path"][0][:, :2]).max() <= 1e-6
    assert np.abs(np.cos(path[:, 2] - out["path"][0][:, 2]) - 1).max() <= 1e-10
    assert new_cmds.shape[0] == batch.n_steps + 1 and path.shape[0] == batch.n_steps + 1


def test_warm_start_memory_blends_previous_solution(oracle, opt_for):
    """Second tick: previous path / cmds (the first tick's OUTPUT, shifted by one step — SURVEY Q12) are blended
    with current_cmds_weight 0.5; reset_memory() restores first-call behaviour."""
    p, poses, cmds, people, speed, costmap, od = _scene("soc_work_obst", 2, seed=5)
    opt = opt_for(p)
    ok1, path1, cmds1, _, _ = opt.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05, od)
    assert ok1
    ok2, path2, cmds2, proj2, info2 = opt.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05,
                                                   od)
    robot, proj_ref, batch, out = _reference_tick(oracle, p, poses, cmds, (path1.tolist(), cmds1.tolist()), people,
                                                  speed, costmap, od)
    # u0 of block 1 is the 0.5/0.5 blend of the seed cmd and the previous optimised cmd (reference :538-547)
    assert batch.arrays["u0"][0, 1, 0] == pytest.approx(0.5 * cmds[0][0] + 0.5 * cmds1[0][0], rel=1e-14)
    assert ok2 == bool(out["usable"][0])
    assert np.abs(cmds2 - out["cmds"][0]).max() <= 1e-6
    assert info2["cost_final"] == pytest.approx(out["cost_final"][0], rel=1e-8)
    opt.reset_memory()
    ok3, path3, cmds3, _, _ = opt.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05, od)
    assert np.array_equal(cmds3, cmds1) and np.array_equal(path3, path1)


def test_optimize_failure_modes(opt_for):
    from nav2_social_mpc_controller_b200 import _lib
    p, poses, cmds, people, speed, costmap, od = _scene("soc_work_obst", 2, seed=7)
    opt = opt_for(p)
    # path with fewer than 2 poses -> false (reference :158-162)
    ok, *_ = opt.optimize(poses[:1], cmds[:1], people, speed, p.time_step, costmap, (0.0, 0.0), 0.05, od)
    assert not ok
    # 100x100 obstacle grid "is NOT valid": every person is dropped -> all projected agents invalid.
    # Ceres >= 2.1: FAILURE on the NaN proxemics Jacobian -> optimize returns false. Ceres 2.0.0
    # (ceres_compat 200, the default): the Jet minimum distance starts at 0, the evaluation is finite, the solve runs.
    od100 = dict(od, width=100, height=100, distances=np.zeros(10000, np.float32), indexes=np.zeros(10000, np.uint32))
    ok, path, new_cmds, proj, info = opt.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05,
                                                  od100)
    assert ok and info["termination"] <= 4
    assert np.all(proj[1:, :, 3] == -1.0)
    p220 = sc.make_params("soc_work_obst", ceres_compat=220)
    opt220 = opt_for(p220)
    ok, path, new_cmds, proj, info = opt220.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05,
                                                     od100)
    assert not ok and info["termination"] == 6
    assert np.all(proj[1:, :, 3] == -1.0)
    # empty obstacle grid -> std::runtime_error in the reference -> call error here
    bad = dict(od, distances=np.zeros(0, np.float32), indexes=np.zeros(0, np.uint32))
    with pytest.raises(_lib.SmpcError):
        opt.optimize(poses, cmds, people, speed, p.time_step, costmap, (0.0, 0.0), 0.05, bad)


def _od_grid(W=80, H=80, res=0.05):
    rows = np.arange(H)[:, None] * np.ones((1, W), dtype=int)
    cols = np.ones((H, 1), dtype=int) * np.arange(W)[None, :]
    near = np.where(np.abs(rows - 12) <= np.abs(rows - 68), 12, 68)
    return dict(width=W, height=H, resolution=res, origin_x=0.0, origin_y=0.0,
                distances=(np.abs(rows - near) * res).astype(np.float32).ravel(),
                indexes=(near * W + cols).astype(np.uint32).ravel())


@pytest.mark.parametrize("A,n_valid", [(3, 3), (3, 1), (7, 5), (20, 20)])
def test_batched_gpu_project_people_matches_reference(opt_for, A, n_valid):
    """GPU project_people (one warp per problem) vs the numpy restatement of reference src/optimizer.cpp:554-671 +
    sfm.hpp, including compaction of the valid people, goals reached on the way and padded columns."""
    rng = np.random.default_rng(A * 10 + n_valid)
    p = sc.make_params("soc_work_obst")
    opt = opt_for(p)
    B = 12
    od = _od_grid()
    robots, inits, want = [], [], []
    for b in range(B):
        pose = np.array([[1.0 + rng.uniform(0, 0.5), 2.0 + rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3)]])
        gpath = sc._straight_path(1, pose[:, 0], np.array([2.0]))
        poses, cmds = sc.pure_pursuit_seed(gpath, pose, p)
        robot, _ = ref.format_to_optimize(poses[0].tolist(), cmds[0].tolist(), poses[0].tolist(), cmds[0].tolist(),
                                          (0.3, 0.0), p.current_path_w, p.current_cmds_w, p.max_time, p.time_step)
        people = []
        for k in range(n_valid):
            px, py = rng.uniform(0.8, 3.2), rng.uniform(1.0, 3.0)
            h, v = rng.uniform(-np.pi, np.pi), rng.uniform(0.0, 0.9) if k % 3 else 0.08  # slow walkers reach their goal
            people.append([px, py, h, 0.0, v, 0.0])
        while len(people) < A:
            people.append([0.0, 0.0, 0.0, -1.0, 0.0, 0.0])
        perm = rng.permutation(A) if b % 2 else np.arange(A)  # invalid columns anywhere, not only at the end
        people = [people[i] for i in perm]
        robots.append(robot)
        inits.append(people)
        want.append(np.transpose(np.array(ref.project_people(people, robot, od, p.max_time, p.time_step)), (1, 2, 0)))
    od_b = dict(width=80, height=80, resolution=0.05, origins=[[0.0, 0.0]], indexes=od["indexes"])
    got, status = opt.project_people_batch(np.array(robots), np.array(inits), od_b, p.max_time, p.time_step)
    assert np.all(status == 0)
    want = np.array(want)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-10, np.abs(got - want).max()


def test_gpu_project_people_grid_quirks(opt_for):
    """100x100 grids drop every person; leaving the grid is reported in status (the reference throws)."""
    p = sc.make_params("soc_work_obst")
    opt = opt_for(p)
    S = 28
    robot = np.zeros((2, S + 1, 6))
    robot[:, :, 0] = 1.0 + 0.03 * np.arange(S + 1)
    robot[:, :, 1] = 2.0
    robot[:, :, 4] = 0.6
    init = np.array([[[2.5, 2.0, np.pi, 0.0, 0.5, 0.0], [0, 0, 0, -1.0, 0, 0], [0, 0, 0, -1.0, 0, 0]]] * 2)
    od100 = dict(width=100, height=100, resolution=0.05, origins=[[0.0, 0.0]], indexes=np.zeros(10000, np.uint32))
    got, status = opt.project_people_batch(robot, init, od100, p.max_time, p.time_step)
    assert np.all(status == 0) and np.all(got[:, :, 3, 1:] == -1.0) and np.all(got[:, 0, 3, 0] == 0.0)
    init[1, 0, 0] = 3.97  # walks out of the 4 m grid within the horizon
    init[1, 0, 2] = 0.0
    od = _od_grid()
    got, status = opt.project_people_batch(robot, init, dict(width=80, height=80, resolution=0.05, origins=[[0.0, 0.0]],
                                                             indexes=od["indexes"]), p.max_time, p.time_step)
    assert status[0] == 0 and status[1] == 1


def test_fleet_tick_matches_per_robot_optimize(opt_for):
    """FleetOptimizer (every stage a kernel, B robots per call) vs the single-robot level-2 entry, two ticks so that
    the per-robot warm-start memory (previous path / cmds) is exercised."""
    from nav2_social_mpc_controller_b200.fleet import FleetOptimizer
    from nav2_social_mpc_controller_b200.optimizer import Optimizer
    B, A = 6, 3
    scenes = [_scene("soc_work_obst", n_people=(b % 3) + 1, seed=20 + b) for b in range(B)]
    p = scenes[0][0]
    od = scenes[0][6]
    costmap = scenes[0][5]
    n_in = scenes[0][1].shape[0]
    poses = np.stack([s[1] for s in scenes])
    cmds = np.stack([s[2] for s in scenes])
    speed = np.array([[0.3, 0.05]] * B)
    people_raw = np.zeros((B, A, 5))
    n_people = np.zeros(B, dtype=np.int32)
    for b, s in enumerate(scenes):
        k = min(A, s[3].shape[0])
        people_raw[b, :k] = s[3][:k]
        n_people[b] = k
    fleet = FleetOptimizer(p, n_robots=B, n_agents=A)
    singles = []
    for b in range(B):
        o = Optimizer(0)
        o.initialize(p)
        singles.append(o)
    try:
        od_b = dict(width=od["width"], height=od["height"], resolution=od["resolution"], origins=[[0.0, 0.0]],
                    indexes=od["indexes"])
        for tick in range(2):
            got = fleet.optimize_batch(poses, cmds, people_raw, n_people, speed, costmap[None], np.zeros((1, 2)), 0.05,
                                       od_b)
            assert np.all(got["project_status"] == 0)
            n_b = got["n_out"]
            for b in range(B):
                ok, path, new_cmds, proj, info = singles[b].optimize(poses[b], cmds[b], scenes[b][3][:A], speed[b],
                                                                      p.time_step, costmap, (0.0, 0.0), 0.05, od)
                assert bool(got["optimized"][b]) == ok
                n = int(n_b[b])
                assert n == path.shape[0]
                assert np.abs(got["people_proj"][b][:, :, :n] - np.transpose(proj, (1, 2, 0))).max() <= 1e-9
                assert got["termination"][b] == info["termination"]
                assert np.abs(got["cmds"][b][:n] - new_cmds).max() <= 1e-6, (tick, b)
                assert np.abs(got["path"][b][:n, :2] - path[:, :2]).max() <= 1e-6
                assert got["cost_final"][b] == pytest.approx(info["cost_final"], rel=1e-8)
    finally:
        fleet.close()
        for o in singles:
            o.close()
    assert n_in >= 2


def test_fleet_tick_with_mixed_path_lengths_matches_reference_pipeline(oracle):
    """smpc_optimize_batch with per-robot horizons: robots near their goal get shorter seeds from the trajectorizer
    (reference src/path_trajectorizer.cpp:152) and with them fewer steps, a shorter control horizon and fewer parameter
    blocks (src/optimizer.cpp:248-249). Every robot is checked against the reference pipeline run for it alone —
    tests/presolve_ref.py (numpy people_to_status / format_to_optimize / project_people) + the CPU oracle — over two
    ticks, so that the per-robot warm-start memory with its own length is exercised. A robot with a one-pose path is
    reported not optimized and left untouched (:158-162)."""
    from nav2_social_mpc_controller_b200.fleet import FleetOptimizer
    lengths = [31, 24, 17, 12, 8, 5, 3, 2, 1, 31]
    B, A = len(lengths), 3
    scenes = [_scene("soc_work_obst", n_people=(b % 3) + 1, seed=40 + b) for b in range(B)]
    p, od, costmap = scenes[0][0], scenes[0][6], scenes[0][5]
    n_full = scenes[0][1].shape[0]
    poses = np.stack([s[1] for s in scenes])
    cmds = np.stack([s[2] for s in scenes])
    n_poses = np.minimum(np.array(lengths), n_full).astype(np.int32)
    speed = np.array([[0.3, 0.05]] * B)
    people_raw = np.zeros((B, A, 5))
    n_people = np.zeros(B, dtype=np.int32)
    for b, s in enumerate(scenes):
        k = min(A, s[3].shape[0])
        people_raw[b, :k] = s[3][:k]
        n_people[b] = k
    od_b = dict(width=od["width"], height=od["height"], resolution=od["resolution"], origins=[[0.0, 0.0]],
                indexes=od["indexes"])
    fleet = FleetOptimizer(p, n_robots=B, n_agents=A)
    prev = [None] * B
    try:
        for tick in range(2):
            got = fleet.optimize_batch(poses, cmds, people_raw, n_people, speed, costmap[None], np.zeros((1, 2)), 0.05,
                                       od_b, n_poses=n_poses)
            for b in range(B):
                nb_in = int(n_poses[b])
                if nb_in < 2:
                    assert not got["optimized"][b] and got["n_out"][b] == nb_in
                    assert np.array_equal(got["path"][b], poses[b]) and np.array_equal(got["cmds"][b][:30], cmds[b][:30])
                    continue
                robot, proj_ref, batch, out = _reference_tick(oracle, p, poses[b][:nb_in], cmds[b][:nb_in - 1], prev[b],
                                                              scenes[b][3][:A], speed[b], costmap, od)
                n = len(robot)
                assert got["n_out"][b] == n, (tick, b)
                want_proj = np.transpose(np.array(proj_ref), (1, 2, 0))
                assert np.abs(got["people_proj"][b][:, :, :n] - want_proj).max() <= 1e-9, (tick, b)
                assert bool(got["optimized"][b]) == bool(out["usable"][0])
                assert got["termination"][b] == out["termination"][0] and got["iterations"][b] == out["iterations"][0]
                assert got["cost_final"][b] == pytest.approx(out["cost_final"][0], rel=1e-8)
                assert np.abs(got["cmds"][b][:n] - out["cmds"][0]).max() <= 1e-6, (tick, b)
                assert np.abs(got["path"][b][:n, :2] - out["path"][0][:, :2]).max() <= 1e-6
                if out["usable"][0]:
                    prev[b] = (out["path"][0].tolist(), out["cmds"][0].tolist())
                elif prev[b] is None:
                    prev[b] = (poses[b][:nb_in].tolist(), cmds[b][:nb_in - 1].tolist())
    finally:
        fleet.close()


def test_fleet_maps_stay_resident_and_person_leaving_the_grid_is_reported():
    """Costmaps / obstacle grids are re-sent only when they change (maps_version); a person who walks out of the obstacle
    grid makes the reference throw — the fleet reports that robot as not optimized and keeps its memory."""
    from nav2_social_mpc_controller_b200.fleet import FleetOptimizer
    B, A = 4, 3
    scenes = [_scene("soc_work_obst", n_people=1, seed=60 + b) for b in range(B)]
    p, od, costmap = scenes[0][0], scenes[0][6], scenes[0][5]
    poses = np.stack([s[1] for s in scenes])
    cmds = np.stack([s[2] for s in scenes])
    people_raw = np.zeros((B, A, 5))
    for b, s in enumerate(scenes):
        people_raw[b, 0] = s[3][0]
    people_raw[2, 0] = [3.97, 2.0, 0.8, 0.0, 0.0]  # leaves the 4 m grid within the horizon
    n_people = np.ones(B, dtype=np.int32)
    speed = np.array([[0.3, 0.05]] * B)
    od_b = dict(width=od["width"], height=od["height"], resolution=od["resolution"], origins=[[0.0, 0.0]],
                indexes=od["indexes"])
    maps = costmap[None].copy()
    fleet = FleetOptimizer(p, n_robots=B, n_agents=A)
    try:
        a = fleet.optimize_batch(poses, cmds, people_raw, n_people, speed, maps, np.zeros((1, 2)), 0.05, od_b)
        v1 = fleet._maps_version
        fleet.reset_memory()
        b = fleet.optimize_batch(poses, cmds, people_raw, n_people, speed, maps, np.zeros((1, 2)), 0.05, od_b)
        assert fleet._maps_version == v1  # same arrays: the maps were not re-sent
        for k in ("cmds", "path", "optimized", "termination", "cost_final"):
            assert np.array_equal(a[k], b[k]), k
        assert a["project_status"][2] == 1 and not a["optimized"][2]
        assert np.array_equal(a["cmds"][2][:cmds.shape[1]], cmds[2])  # the seed cmds are kept, as the caller would
        assert a["optimized"][[0, 1, 3]].all()
    finally:
        fleet.close()


def test_fleet_edge_cases_short_paths_and_changing_lengths_keep_the_memory():
    """Robots whose seed path has fewer than 2 poses return optimized = False (src/optimizer.cpp:158-162) without
    disturbing their neighbours; a robot whose path gets shorter from one tick to the next (approaching the goal) keeps
    its warm-start memory — the second tick of the fleet equals the second tick of a one-robot optimizer fed the same
    two paths (the reference's TrajectoryMemory survives a change of the path length)."""
    from nav2_social_mpc_controller_b200.fleet import FleetOptimizer
    from nav2_social_mpc_controller_b200.optimizer import Optimizer
    B, A = 3, 3
    scenes = [_scene("soc_work_obst", n_people=1, seed=80 + b) for b in range(B)]
    p, od, costmap = scenes[0][0], scenes[0][6], scenes[0][5]
    poses = np.stack([s[1] for s in scenes])
    cmds = np.stack([s[2] for s in scenes])
    n = poses.shape[1]
    people_raw = np.zeros((B, A, 5))
    for b, s in enumerate(scenes):
        people_raw[b, 0] = s[3][0]
    n_people = np.ones(B, dtype=np.int32)
    speed = np.array([[0.3, 0.05]] * B)
    od_b = dict(width=od["width"], height=od["height"], resolution=od["resolution"], origins=[[0.0, 0.0]],
                indexes=od["indexes"])
    maps = costmap[None].copy()
    fleet = FleetOptimizer(p, n_robots=B, n_agents=A)
    single = Optimizer(0)
    single.initialize(p)
    try:
        n1 = np.array([n, 1, n], dtype=np.int32)  # robot 1 has no path to optimise
        t1 = fleet.optimize_batch(poses, cmds, people_raw, n_people, speed, maps, np.zeros((1, 2)), 0.05, od_b, n_poses=n1)
        assert t1["optimized"].tolist() == [True, False, True]
        n2 = np.array([n, 1, n - 6], dtype=np.int32)  # robot 2's path got shorter
        t2 = fleet.optimize_batch(poses, cmds, people_raw, n_people, speed, maps, np.zeros((1, 2)), 0.05, od_b, n_poses=n2)
        assert t2["optimized"].tolist() == [True, False, True] and t2["n_out"][2] <= n - 6
        # the same two ticks for robot 2 alone through the one-robot entry
        s = scenes[2]
        ok1, *_ = single.optimize(s[1], s[2], s[3][:1], speed[2], p.time_step, costmap, (0.0, 0.0), 0.05, od)
        ok2, path2, cmds2, _, _ = single.optimize(s[1][:n - 6], s[2][:n - 6], s[3][:1], speed[2], p.time_step, costmap,
                                                  (0.0, 0.0), 0.05, od)
        assert ok1 and ok2
        k = int(t2["n_out"][2])
        assert k == path2.shape[0]
        assert np.array_equal(t2["cmds"][2][:k], cmds2[:k]) and np.array_equal(t2["path"][2][:k], path2[:k])
    finally:
        fleet.close()
        single.close()


def test_gpu_trajectorize_matches_numpy_restatement():
    """Seed generation kernel vs scenarios.pure_pursuit_seed (numpy restatement of reference
    src/path_trajectorizer.cpp:120-288), diff-drive branch, goals beyond the horizon; plus the early stop near the
    goal and the omnidirectional branch against a scalar Python restatement."""
    from nav2_social_mpc_controller_b200.fleet import FleetOptimizer
    rng = np.random.default_rng(4)
    B = 64
    p = sc.make_params("soc_work_obst")
    pose = np.stack([rng.uniform(0.5, 1.0, B), 2.0 + rng.uniform(-0.4, 0.4, B), rng.uniform(-2.5, 2.5, B)], axis=1)
    gpath = sc._straight_path(B, np.full(B, 0.6), np.full(B, 2.0))
    want_poses, want_cmds = sc.pure_pursuit_seed(gpath, pose, p)
    fleet = FleetOptimizer(p, n_robots=B)
    try:
        poses, cmds, n_steps = fleet.trajectorize_batch(gpath, pose)
        assert np.all(n_steps == want_cmds.shape[1])
        assert np.abs(poses[:, :, :2] - want_poses[:, :, :2]).max() <= 1e-12
        assert np.abs(np.cos(poses[:, :, 2] - want_poses[:, :, 2]) - 1).max() <= 1e-12
        assert np.abs(cmds[:, :, 0] - want_cmds[:, :, 0]).max() <= 1e-12 and np.all(cmds[:, :, 1] == 0.0)
        assert np.abs(cmds[:, :, 2] - want_cmds[:, :, 1]).max() <= 1e-11
        # early stop: a goal 0.5 m ahead is reached (<= 0.2 m) after a few steps
        short = sc._straight_path(B, pose[:, 0], pose[:, 1], length=0.5)
        pose0 = pose.copy()
        pose0[:, 2] = 0.0
        _, _, n_short = fleet.trajectorize_batch(short, pose0)
        want_n = int(np.ceil((0.5 - 0.2) / (0.6 * 0.05) - 1e-9))
        assert np.all(n_short == want_n), (n_short[:4], want_n)
    finally:
        fleet.close()
    po = sc.make_params("soc_work_obst", omnidirectional=1)
    fleet = FleetOptimizer(po, n_robots=4)
    try:
        poses, cmds, n_steps = fleet.trajectorize_batch(gpath[:4], pose[:4])
        for b in range(4):
            rx, ry, rth = pose[b, 0], pose[b, 1], float(sc.yaw_roundtrip(pose[b, 2]))
            for s in range(3):
                d = np.hypot(rx - gpath[b, :, 0], ry - gpath[b, :, 1])
                within = np.nonzero(d <= po.lookahead_dist)[0]
                wp = within[-1] if within.size else (len(d) - 1 - int(np.argmin(d[::-1])))
                dx = (gpath[b, wp, 0] - rx) * math.cos(rth) + (gpath[b, wp, 1] - ry) * math.sin(rth)
                dy = -(gpath[b, wp, 0] - rx) * math.sin(rth) + (gpath[b, wp, 1] - ry) * math.cos(rth)
                dth = math.atan2(dy, dx)
                vx, vy = 0.6 * math.cos(dth), 0.6 * math.sin(dth)
                assert cmds[b, s, 0] == pytest.approx(vx, abs=1e-12) and cmds[b, s, 1] == pytest.approx(vy, abs=1e-12)
                assert cmds[b, s, 2] == 0.0
                rx += (vx * math.cos(rth) + vy * math.cos(math.pi / 2 + rth)) * 0.05
                ry += (vx * math.sin(rth) + vy * math.sin(math.pi / 2 + rth)) * 0.05
                assert poses[b, s + 1, 0] == pytest.approx(rx, abs=1e-12)
                assert poses[b, s + 1, 1] == pytest.approx(ry, abs=1e-12)
    finally:
        fleet.close()


def test_fov_people_filter_matches_reference_arithmetic():
    """Batched FOV filter vs a scalar restatement of reference src/social_mpc_controller.cpp:198-214 with its float
    roundings (angle_to_person, robot_yaw, relative_angle are floats; fov_angle is a double)."""
    from nav2_social_mpc_controller_b200.fleet import FleetOptimizer
    rng = np.random.default_rng(8)
    B, K = 64, 7
    p = sc.make_params("soc_work_obst")
    pose = np.stack([rng.uniform(1.0, 3.0, B), rng.uniform(1.0, 3.0, B), rng.uniform(-3.1, 3.1, B)], axis=1)
    people = np.zeros((B, K, 5))
    people[:, :, 0] = rng.uniform(-0.5, 4.5, (B, K))  # some outside the 4 x 4 m costmap
    people[:, :, 1] = rng.uniform(-0.5, 4.5, (B, K))
    people[:, :, 2:5] = rng.normal(0, 0.5, (B, K, 3))
    n_people = rng.integers(0, K + 1, B).astype(np.int32)
    fleet = FleetOptimizer(p, n_robots=B, n_agents=3)
    try:
        got, n_got = fleet.filter_people_fov(people, n_people, pose, np.zeros((1, 2)), 80, 80, 0.05)
    finally:
        fleet.close()
    fov = float(p.fov_angle)
    for b in range(B):
        yaw = np.float32(sc.yaw_roundtrip(pose[b, 2]))
        keep = []
        for k in range(n_people[b]):
            wx, wy = people[b, k, 0], people[b, k, 1]
            if wx < 0.0 or wy < 0.0 or not (int(wx / 0.05) < 80 and int(wy / 0.05) < 80):
                continue
            ang = np.float32(math.atan2(wy - pose[b, 1], wx - pose[b, 0]))
            d = float(ang) - float(yaw)
            rel = np.float32(math.fmod(math.fmod(d + math.pi, 2 * math.pi) + 2 * math.pi, 2 * math.pi) - math.pi)
            if abs(float(rel)) < fov:
                keep.append(k)
        assert n_got[b] == len(keep), b
        for j, k in enumerate(keep[:3]):
            assert np.array_equal(got[b, j], people[b, k])
        assert np.all(got[b, min(len(keep), 3):] == 0.0)

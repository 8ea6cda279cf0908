"""TEST INFRASTRUCTURE — numpy / pure-Python restatement of the reference's pre-solve stages for ONE robot, used as
the oracle of the level-2 entry smpc_optimize (the solve itself is checked against oracle/liboracle.so).
Follows reference src/optimizer.cpp:454-482 (people_to_status), :484-551 (format_to_optimize), :554-671
(project_people), :673-728 (computeObstacle) and include/nav2_social_mpc_controller/sfm.hpp:188-323, 462-560.
Independent of the product's C++ implementation (csrc/smpc_optimize.cu). PARITY UNPINNED (no reference tests)."""
import math

import numpy as np

from nav2_social_mpc_controller_b200 import abi, scenarios as sc


def f32(x):
    return float(np.float32(x))


def wrap_pi(a):
    while a <= -math.pi:
        a += 2 * math.pi
    while a > math.pi:
        a -= 2 * math.pi
    return a


def rt(yaw):
    h = yaw * 0.5
    qz, qw = math.sin(h), math.cos(h)
    return math.atan2(2 * (qw * qz), qw * qw - qz * qz)


def people_to_status(people):
    out = []
    for p in people:
        out.append([p[0], p[1], math.atan2(p[3], p[2]), 0.0, math.sqrt(p[2] * p[2] + p[3] * p[3]), p[4]])
    while len(out) < 3:
        out.append([0.0, 0.0, 0.0, -1.0, 0.0, 0.0])
    return out[:3]


def format_to_optimize(poses, cmds, prev_poses, prev_cmds, speed, wpath, wcmd, maxtime, timestep):
    maxsize = int(round(f32(np.float32(maxtime) / np.float32(timestep))))
    poses = [list(p) for p in poses]
    if len(poses) > maxsize:
        poses = poses[: maxsize - 1]
    wpath, wcmd = f32(wpath), f32(wcmd)
    robot = []
    for i, p in enumerate(poses):
        x, y, yaw = p
        if i < len(prev_poses):
            x = wpath * p[0] + (1.0 - wpath) * prev_poses[i][0]
            y = wpath * p[1] + (1.0 - wpath) * prev_poses[i][1]
            yaw = rt(wpath * p[2] + (1.0 - wpath) * prev_poses[i][2])
            poses[i] = [x, y, yaw]
        t = float(np.float32(i) * np.float32(timestep))
        if i == 0:
            lv, av = speed
        else:
            pc = prev_cmds[i - 1] if i - 1 < len(prev_cmds) else cmds[i - 1]
            lv = wcmd * cmds[i - 1][0] + (1.0 - wcmd) * pc[0]
            av = wcmd * cmds[i - 1][1] + (1.0 - wcmd) * pc[1]
        robot.append([x, y, yaw, t, lv, av])
    return robot, poses


def compute_obstacle(pos, od):
    res = f32(od["resolution"])
    xc = int(math.floor((pos[0] - od["origin_x"]) / res))
    yc = int(math.floor((pos[1] - od["origin_y"]) / res))
    assert 0 <= xc < od["width"] and 0 <= yc < od["height"]
    ob = int(od["indexes"][xc + yc * od["width"]])
    oy, ox = ob // od["width"], ob % od["width"]
    x = f32(np.float32(ox) * np.float32(res) + od["origin_x"]) if False else float(np.float32(float(np.float32(ox) * np.float32(res)) + od["origin_x"]))
    y = float(np.float32(float(np.float32(oy) * np.float32(res)) + od["origin_y"]))
    return [pos[0] - x, pos[1] - y]


def _norm(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1])


def _normalized(v):
    z = v[0] * v[0] + v[1] * v[1]
    if z > 0:
        s = math.sqrt(z)
        return [v[0] / s, v[1] / s]
    return list(v)


def sfm_forces(agents):
    for i, a in enumerate(agents):
        if a["goal"] is not None and _norm([a["goal"][0] - a["pos"][0], a["goal"][1] - a["pos"][1]]) > 0.25:
            d = _normalized([a["goal"][0] - a["pos"][0], a["goal"][1] - a["pos"][1]])
            des = [2.0 * (d[0] * a["vdes"] - a["vel"][0]) / 0.5, 2.0 * (d[1] * a["vdes"] - a["vel"][1]) / 0.5]
        else:
            des = [-a["vel"][0] / 0.5, -a["vel"][1] / 0.5]
        obs = [0.0, 0.0]
        if a["obstacle"] is not None:
            md = [a["pos"][0] - a["obstacle"][0], a["pos"][1] - a["obstacle"][1]]
            dist = _norm(md) - a["radius"]
            n = _normalized(md)
            k = 20 * math.exp(-dist / 0.2)
            obs = [k * n[0], k * n[1]]
        soc = [0.0, 0.0]
        for k2, o in enumerate(agents):
            if k2 == i:
                continue
            diff = [o["pos"][0] - a["pos"][0], o["pos"][1] - a["pos"][1]]
            dd = _normalized(diff)
            vd = [a["vel"][0] - o["vel"][0], a["vel"][1] - o["vel"][1]]
            iv = [2.0 * vd[0] + dd[0], 2.0 * vd[1] + dd[1]]
            il = _norm(iv)
            idr = [iv[0] / il, iv[1] / il]
            a1 = wrap_pi(math.atan2(idr[1], idr[0]))
            a2 = wrap_pi(math.atan2(dd[1], dd[0]))
            th = wrap_pi(a2 - a1)
            B = 0.35 * il
            fv = -math.exp(-_norm(diff) / B - (3.0 * B * th) ** 2)
            sgn = 0.0 if th == 0 else (1.0 if th > 0 else -1.0)
            fa = -sgn * math.exp(-_norm(diff) / B - (2.0 * B * th) ** 2)
            soc[0] += 2.1 * (fv * idr[0] + fa * (-idr[1]))
            soc[1] += 2.1 * (fv * idr[1] + fa * idr[0])
        a["force"] = [des[0] + soc[0] + obs[0], des[1] + soc[1] + obs[1]]


def sfm_update(agents, dt):
    for a in agents:
        a["vel"] = [a["vel"][0] + a["force"][0] * dt, a["vel"][1] + a["force"][1] * dt]
        if _norm(a["vel"]) > a["vdes"]:
            n = _normalized(a["vel"])
            a["vel"] = [n[0] * a["vdes"], n[1] * a["vdes"]]
        y0 = a["yaw"]
        a["yaw"] = wrap_pi(math.atan2(a["vel"][1], a["vel"][0]))
        a["av"] = wrap_pi(a["yaw"] - y0) / dt
        a["pos"] = [a["pos"][0] + a["vel"][0] * dt, a["pos"][1] + a["vel"][1] * dt]
        a["lv"] = _norm(a["vel"])
        if a["goal"] is not None and _norm([a["goal"][0] - a["pos"][0], a["goal"][1] - a["pos"][1]]) <= 0.25:
            a["goal"] = None


def project_people(init_people, robot, od, maxtime, timestep):
    maxtime, timestep = f32(maxtime), f32(timestep)
    traj = [[list(s) for s in init_people]]
    agents = []
    for s in init_people:
        if s[3] == -1:
            continue
        vel = [s[4] * math.cos(s[2]), s[4] * math.sin(s[2])]
        a = dict(pos=[s[0], s[1]], yaw=s[2], lv=s[4], av=s[5], vel=vel, vdes=0.5, radius=0.5,
                 goal=[s[0] + maxtime * vel[0], s[1] + maxtime * vel[1]], obstacle=None, force=[0, 0])
        if od["width"] == 100 and od["height"] == 100:
            continue
        a["obstacle"] = compute_obstacle(a["pos"], od)
        agents.append(a)
    for i in range(len(robot) - 1):
        r = robot[i]
        rb = dict(pos=[r[0], r[1]], yaw=r[2], lv=r[4], av=r[5], vel=[r[4] * math.cos(r[2]), r[4] * math.sin(r[2])],
                  vdes=0.6, radius=0.5, goal=[robot[-1][0], robot[-1][1]], obstacle=None, force=[0, 0])
        agents.append(rb)
        sfm_forces(agents)
        sfm_update(agents, timestep)
        agents.pop()
        for a in agents:
            a["obstacle"] = compute_obstacle(a["pos"], od)
        humans = [[a["pos"][0], a["pos"][1], a["yaw"], float(np.float32(i + 1) * np.float32(timestep)), a["lv"], a["av"]]
                  for a in agents]
        while len(humans) < len(init_people):
            humans.append([0.0, 0.0, 0.0, -1.0, 0.0, 0.0])
        traj.append(humans)
    return traj


def build_level1_batch(params: abi.SmpcParams, robot, proj, has_people, costmap, origin, resolution, timestep):
    """Unpacking of reference src/optimizer.cpp:197-237 into the include/smpc.h layout (one problem)."""
    S = len(robot) - 1
    ch, bl, nb, _ = abi.problem_dims(params.control_horizon, params.parameter_block_length, S)
    robot = np.array(robot)
    arrays = dict(
        pose0=np.array([[robot[0, 0], robot[0, 1], rt(robot[0, 2])]]),
        u0=np.ascontiguousarray(robot[None, :nb, 4:6]),
        path_xy=np.ascontiguousarray(np.stack([robot[:, 0], robot[:, 1]])[None]),
        goal_yaw=np.array([robot[-1, 2]]),
        agents=np.ascontiguousarray(np.transpose(np.array(proj), (1, 2, 0))[None]),
        has_people=np.array([1 if has_people else 0], dtype=np.uint8),
        costmaps=np.ascontiguousarray(costmap, dtype=np.uint8)[None],
        costmap_origin=np.array([origin], dtype=np.float64),
        costmap_index=None,
    )
    return sc.Batch(params=params, n_problems=1, n_steps=S, n_agents=3, n_costmaps=1, size_x=costmap.shape[1],
                    size_y=costmap.shape[0], resolution=resolution, dt=f32(timestep), arrays=arrays)

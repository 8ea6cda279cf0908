"""50-digit mpmath restatement of the residual set of the social-MPC problem — TEST INFRASTRUCTURE.

Written from SURVEY.md Appendix D (formulas) and Appendix B (bicubic), not from oracle/critics.hpp: plain Python
functions over mpmath numbers, one residual list per problem, Jacobian by central differences at 50 digits (step
1e-20: truncation error ~1e-40, no cancellation at this precision). tests/test_oracle_residuals.py requires the C++
oracle's double residuals and Jet Jacobians to agree with it to round-off.
Models the functors' mathematical value, i.e. Ceres >= 2.1 semantics for ProxemicsCost (ceres_compat 220).
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50
PI = mp.pi


def _wrap_to_pi(a):  # social_work_cost_function.hpp:39-46
    while a > PI:
        a -= 2 * PI
    while a <= -PI:
        a += 2 * PI
    return a


def _wrap_atan(a):  # atan2(sin a, cos a)
    return mp.atan2(mp.sin(a), mp.cos(a))


def _block_of(j, ch, bl):
    return j // bl if j < ch else (ch - 1) // bl


def rollout(pose0, u, i, dt, ch, bl):
    """update_state.hpp:37-63: pose after steps 0..i."""
    x, y, th = (mp.mpf(v) for v in pose0)
    for j in range(i + 1):
        b = _block_of(j, ch, bl)
        v, w = u[2 * b], u[2 * b + 1]
        x, y, th = x + v * mp.cos(th) * dt, y + v * mp.sin(th) * dt, th + w * dt
    return x, y, th


def _social_force(me_xy, me_vel, other_xy, other_vel):
    """One (me <- other) interaction, SURVEY Appendix D row 1."""
    dx, dy = me_xy[0] - other_xy[0], me_xy[1] - other_xy[1]
    if mp.sqrt(dx * dx + dy * dy) < mp.mpf("1e-6"):
        dx, dy = mp.mpf("1e-6"), mp.mpf(0)
    dist = mp.sqrt(dx * dx + dy * dy)
    ex, ey = dx / dist, dy / dist
    wx, wy = me_vel[0] - other_vel[0], me_vel[1] - other_vel[1]
    ix, iy = 2 * wx + ex, 2 * wy + ey
    il = mp.sqrt(ix * ix + iy * iy)
    ux, uy = ix / il, iy / il
    theta = _wrap_to_pi(mp.atan2(ey, ex) - mp.atan2(uy, ux))
    B = mp.mpf("0.35") * il
    fv = -mp.exp(-dist / B - (3 * B * theta) ** 2)
    sgn = 1 if theta > 0 else -1
    fa = -sgn * mp.exp(-dist / B - (2 * B * theta) ** 2)
    return mp.mpf("2.1") * (fv * ux + fa * (-uy)), mp.mpf("2.1") * (fv * uy + fa * ux)


def _hermite(p0, p1, p2, p3, x):
    a = (-p0 + 3 * p1 - 3 * p2 + p3) / 2
    b = (2 * p0 - 5 * p1 + 4 * p2 - p3) / 2
    c = (-p0 + p2) / 2
    return p1 + x * (c + x * (b + x * a))


def bicubic(cmap, r, c):
    """ceres::BiCubicInterpolator<Grid2D<u_char>>::Evaluate value, clamped borders."""
    H, W = cmap.shape
    row, col = int(mp.floor(r)), int(mp.floor(c))

    def g(rr, cc):
        return mp.mpf(int(cmap[min(max(rr, 0), H - 1), min(max(cc, 0), W - 1)]))
    rows = [_hermite(g(row - 1 + k, col - 1), g(row - 1 + k, col), g(row - 1 + k, col + 1), g(row - 1 + k, col + 2),
                     c - col) for k in range(4)]
    return _hermite(rows[0], rows[1], rows[2], rows[3], r - row)


def residuals(prob, u):
    """Residual vector in AddResidualBlock order (reference src/optimizer.cpp:251-371). prob: dict with pose0, dt, S,
    ch, bl, nbd, px, py, goal_yaw, agents [A][6][S+1] or None, has_people, cmap, origin, res, weights w_*."""
    u = [mp.mpf(v) for v in u]
    S, ch, bl, dt = prob["S"], prob["ch"], prob["bl"], mp.mpf(prob["dt"])
    x0, y0, yaw0 = (mp.mpf(v) for v in prob["pose0"])
    out = []
    for i in range(S):
        X, Y, Th = rollout(prob["pose0"], u, i, dt, ch, bl)
        b = _block_of(i, ch, bl)
        lv = u[2 * b]
        if prob["has_people"]:
            ag = prob["agents"]
            A = ag.shape[0]
            col = [[mp.mpf(float(ag[k, c, i + 1])) for c in range(6)] for k in range(A)]
            # --- AgentAngle
            best, closest = None, -1
            for k in range(A):
                d2 = (col[k][0] - x0) ** 2 + (col[k][1] - y0) ** 2
                if (best is None or d2 < best) and col[k][4] > mp.mpf("0.05"):
                    best, closest = d2, k
            r = mp.mpf(0)
            if closest >= 0 and best <= 4:
                a = col[closest]
                phi = mp.atan2(a[1] - y0, a[0] - x0)
                h = _wrap_atan(a[2] - yaw0)
                if h <= -5 * PI / 6 or h >= PI / 6:
                    if not (_wrap_atan(phi - yaw0) < 0):
                        r = prob["w_agent_angle"] * _wrap_atan(Th - (yaw0 - PI / 6)) ** 2
                else:
                    if not (_wrap_atan(phi - yaw0) > 0):
                        r = prob["w_agent_angle"] * _wrap_atan(Th - (yaw0 + PI / 6)) ** 2
            out.append(r)
            # --- SocialWork
            rv = (lv * mp.cos(Th), lv * mp.sin(Th))
            fx = fy = mp.mpf(0)
            for k in range(A):
                if col[k][3] == -1:
                    continue
                av = (col[k][4] * mp.cos(col[k][2]), col[k][4] * mp.sin(col[k][2]))
                f = _social_force((X, Y), rv, (col[k][0], col[k][1]), av)
                fx, fy = fx + f[0], fy + f[1]
            wr = fx * fx + fy * fy
            wp = mp.mpf(0)
            for k in range(A):  # all columns, padded ones too
                av = (col[k][4] * mp.cos(col[k][2]), col[k][4] * mp.sin(col[k][2]))
                f = _social_force((col[k][0], col[k][1]), av, (X, Y), rv)
                wp += f[0] * f[0] + f[1] * f[1]
            out.append(prob["w_social"] * (wr + wp + mp.mpf("1e-6")))
            # --- Proxemics
            dmin = None
            for k in range(A):
                if col[k][3] == -1:
                    continue
                d2 = (X - col[k][0]) ** 2 + (Y - col[k][1]) ** 2
                dmin = d2 if dmin is None or d2 < dmin else dmin
            out.append(prob["w_prox"] * 3 * mp.exp(-dmin / mp.mpf("0.25")) if dmin is not None else mp.mpf(0))
        out.append(prob["w_velocity"] * (mp.mpf("0.6") - u[2 * (i // bl)]) ** 2 if i < ch else mp.mpf(0))
        out.append(prob["w_goal"] * _wrap_atan(mp.mpf(prob["goal_yaw"]) - Th) ** 2)
        ex, ey = X - mp.mpf(float(prob["px"][S])), Y - mp.mpf(float(prob["py"][S]))
        out.append(prob["w_distance"] * (ex * ex + ey * ey) ** 2)
        ex, ey = X - mp.mpf(float(prob["px"][i + 1])), Y - mp.mpf(float(prob["py"][i + 1]))
        out.append(prob["w_angle"] * (ex * ex + ey * ey) ** 2)
        fxw, fyw = X + mp.mpf("0.25") * mp.cos(Th), Y + mp.mpf("0.25") * mp.sin(Th)
        gx = (fxw - mp.mpf(prob["origin"][0])) / mp.mpf(prob["res"])
        gy = (fyw - mp.mpf(prob["origin"][1])) / mp.mpf(prob["res"])
        out.append(prob["w_obstacle"] * bicubic(prob["cmap"], gy, gx))
        if i != 0 and i < ch // bl:
            out.append(prob["w_vf"] * (u[2 * i] - u[2 * i - 2]) ** 2 + prob["w_vf"] * (u[2 * i + 1] - u[2 * i - 1]) ** 2)
    return out


def jacobian(prob, u, h="1e-20"):
    h = mp.mpf(h)
    u = [mp.mpf(v) for v in u]
    cols = []
    for c in range(len(u)):
        up, um = list(u), list(u)
        up[c] += h
        um[c] -= h
        rp, rm = residuals(prob, up), residuals(prob, um)
        cols.append([(a - b) / (2 * h) for a, b in zip(rp, rm)])
    return [[cols[c][k] for c in range(len(u))] for k in range(len(cols[0]))]


def problem_from_batch(batch, b):
    """Problem dict for problem b of a level-1 batch (include/smpc.h layout)."""
    from nav2_social_mpc_controller_b200 import abi
    p = batch.params
    S = batch.n_steps
    ch, bl, nb, nbd = abi.problem_dims(p.control_horizon, p.parameter_block_length, S)
    a = batch.arrays
    mi = int(a["costmap_index"][b]) if a.get("costmap_index") is not None else b % batch.n_costmaps
    w = {k: mp.mpf(float(v)) for k, v in dict(
        w_distance=p.distance_w, w_social=p.socialwork_w, w_velocity=p.velocity_w, w_angle=p.angle_w,
        w_agent_angle=p.agent_angle_w, w_prox=p.proxemics_w, w_vf=p.velocity_feasibility_w, w_obstacle=p.obstacle_w,
        w_goal=p.goal_align_w).items()}
    return dict(pose0=[float(v) for v in a["pose0"][b]], dt=float(batch.dt), S=S, ch=ch, bl=bl, nbd=nbd,
                px=a["path_xy"][b, 0], py=a["path_xy"][b, 1], goal_yaw=float(a["goal_yaw"][b]),
                agents=None if a.get("agents") is None else a["agents"][b],
                has_people=bool(a["has_people"][b]) and a.get("agents") is not None, cmap=a["costmaps"][mi],
                origin=[float(v) for v in a["costmap_origin"][mi]], res=float(batch.resolution), **w)

"""CPU-only checks that pin the oracle's residuals and Jacobians (the reference has no golden
vectors, so the oracle is pinned by finite differences, closed-form known answers and structure checks)."""
import math

import numpy as np
import pytest

from nav2_social_mpc_controller_b200 import scenarios as sc

K_AGENT_ANGLE, K_SOCIAL, K_PROX, K_VELOCITY, K_GOAL, K_PATH_FOLLOW, K_PATH_ALIGN, K_OBSTACLE, K_VEL_FEAS = range(9)


def _fd_jac(oracle, batch, b, x, h=1e-6):
    P = x.size
    m = oracle.num_residuals(batch, b)
    J = np.zeros((m, P))
    for c in range(P):
        xp, xm = x.copy(), x.copy()
        xp[c] += h
        xm[c] -= h
        J[:, c] = (oracle.evaluate(batch, b, xp, want_jac=False)["residuals"]
                   - oracle.evaluate(batch, b, xm, want_jac=False)["residuals"]) / (2 * h)
    return J


@pytest.mark.parametrize("name,S,P,m", [("readme", 13, 6, 105), ("soc_work_obst", 28, 6, 226),
                                         ("params_yaml", 38, 10, 308)])
def test_problem_sizes_match_survey_table(oracle, name, S, P, m):
    """SURVEY §8 derived-size table (reference src/optimizer.cpp:248-249,251-371)."""
    b = sc.single(name)
    assert b.n_steps == S
    assert 2 * b.n_blocks == P
    assert oracle.num_residuals(b) == m


def test_residual_order_per_step(oracle):
    """AddResidualBlock order: AgentAngle, SocialWork, Proxemics, Velocity, GoalAlign, PathFollow, PathAlign,
    Obstacle, [VelFeas for 0 < i < ch/bl] (reference src/optimizer.cpp:292-294,323-324,361-363,368)."""
    b = sc.single("soc_work_obst")
    kinds, steps = oracle.layout(b)
    assert list(kinds[:8]) == list(range(8))
    assert list(kinds[8:17]) == list(range(8)) + [K_VEL_FEAS]
    assert (kinds == K_VEL_FEAS).sum() == b.dims[3] - 1
    assert steps[-1] == b.n_steps - 1
    nb = sc.corridor(B=2)
    kinds, _ = oracle.layout(nb)
    assert set(kinds.tolist()) == {K_VELOCITY, K_GOAL, K_PATH_FOLLOW, K_PATH_ALIGN, K_OBSTACLE, K_VEL_FEAS}


@pytest.mark.parametrize("name", ["readme", "soc_work_obst", "params_yaml"])
def test_jet_jacobian_matches_central_differences(oracle, name):
    # Ceres >= 2.1 semantics: every functor differentiates to the Jacobian of its own value (under Ceres 2.0.0 the
    # proxemics rows do not: test_proxemics_under_ceres_200_jets below)
    b = sc.single(name, seed_offset=3, ceres_compat=220)
    rng = np.random.default_rng(5)
    P = 2 * b.n_blocks
    for trial in range(3):
        x = b.arrays["u0"][0].ravel() + rng.normal(0, 0.05, P)
        x[0::2] = np.clip(x[0::2], 0.05, 0.55)
        e = oracle.evaluate(b, 0, x)
        assert e["ok"]
        J = _fd_jac(oracle, b, 0, x)
        scale = np.maximum(1.0, np.abs(e["jac"]).max())
        assert np.abs(J - e["jac"]).max() / scale < 2e-5
        assert np.allclose(e["grad"], e["jac"].T @ e["residuals"], rtol=1e-12, atol=1e-9)
        assert math.isclose(e["cost"], 0.5 * float(e["residuals"] @ e["residuals"]), rel_tol=1e-13)


def test_causal_jacobian_structure(oracle):
    """Residual of step i only sees blocks 0..b(i) (reference update_state.hpp:46-61)."""
    b = sc.single("soc_work_obst")
    ch, bl, nb, _ = b.dims
    e = oracle.evaluate(b, 0, b.arrays["u0"][0].ravel())
    kinds, steps = oracle.layout(b)
    for k, (kind, i) in enumerate(zip(kinds, steps)):
        if kind == K_VEL_FEAS:
            continue
        last = (i // bl if i < ch else (ch - 1) // bl)
        assert np.all(e["jac"][k, 2 * (last + 1):] == 0.0)


def test_velocity_only_known_answer(oracle):
    """Only VelocityCost active: minimiser is v = 0.6 for every bounded block, residual w*(0.6-v)^2."""
    b = sc.corridor(B=1, distance_w=0.0, angle_w=0.0, goal_align_w=0.0, obstacle_w=0.0,
                    velocity_feasibility_w=0.0, velocity_w=10.0)
    x = b.arrays["u0"][0].ravel().copy()
    e = oracle.evaluate(b, 0, x)
    kinds, steps = oracle.layout(b)
    ch, bl, nb, _ = b.dims
    for k in np.nonzero(kinds == K_VELOCITY)[0]:
        i = steps[k]
        want = 10.0 * (0.6 - x[2 * (i // bl)]) ** 2 if i < ch else 0.0
        assert e["residuals"][k] == pytest.approx(want, rel=1e-15, abs=0.0)
    out = oracle.solve_batch(b)
    assert out["usable"][0] == 1
    assert np.allclose(out["u"][0, :, 0], 0.6, atol=5e-4)  # quartic cost: fn_tol stops just short of 0.6
    assert out["cost_final"][0] < 1e-10


def test_proxemics_under_ceres_200_jets(oracle):
    """proxemics_cost_function.hpp:128 initialises min_distance with std::numeric_limits<T>::max(). Ceres 2.0.0 has no
    numeric_limits specialisation for Jets (added in 2.1), so under T = Jet the primary template returns Jet() = 0:
    every DIFFERENTIATED evaluation sees the constant residual w * 3 * exp(-0) with a zero Jacobian row, while the
    cost-only evaluation sees the true minimum distance. ceres_compat = 200 (default) models exactly that split."""
    b = sc.single("soc_work_obst", n_people=3)
    assert b.params.ceres_compat == 200
    x = b.arrays["u0"][0].ravel()
    kinds, _ = oracle.layout(b)
    prox = kinds == K_PROX
    plain = oracle.evaluate(b, 0, x, want_jac=False)
    diff = oracle.evaluate(b, 0, x, want_jac=True)
    assert plain["ok"] and diff["ok"]
    w = b.params.proxemics_w
    assert np.all(diff["residuals"][prox] == 3.0 * w)
    assert np.all(diff["jac"][prox] == 0.0)
    assert np.all(plain["residuals"][prox] < 3.0 * w) and np.all(plain["residuals"][prox] >= 0.0)
    # every other residual is the same number in both evaluations up to round-off (Jet division multiplies by the
    # reciprocal, ceres/jet.h, where the double evaluation divides)
    assert np.allclose(plain["residuals"][~prox], diff["residuals"][~prox], rtol=1e-13, atol=0)
    # Ceres >= 2.1: both evaluations agree on the proxemics rows too
    b2 = sc.single("soc_work_obst", n_people=3, ceres_compat=220)
    p2 = oracle.evaluate(b2, 0, x, want_jac=False)
    d2 = oracle.evaluate(b2, 0, x, want_jac=True)
    assert np.allclose(p2["residuals"], d2["residuals"], rtol=1e-13, atol=0)
    assert np.array_equal(p2["residuals"], plain["residuals"])
    # no valid agent: Jet(0) stays the minimum, the evaluation is finite and the solve runs (no FAILURE as under >= 2.1)
    b3 = sc.single("soc_work_obst", n_people=0)
    b3.arrays["has_people"][:] = 1
    d3 = oracle.evaluate(b3, 0, x, want_jac=True)
    assert d3["ok"] and np.all(d3["residuals"][prox] == 3.0 * w)
    assert oracle.solve_batch(b3)["usable"][0] == 1


def test_proxemics_and_social_phantom_quirks(oracle):
    """with every agent column invalid the social residual is w*(1e-6 + sum_k |F(phantom_k <- robot)|^2)
    and the proxemics VALUE is exp(-DBL_MAX/0.25) = 0 (under Ceres >= 2.1 its jet derivative is NaN, so a
    differentiated evaluation fails, which makes Ceres terminate with FAILURE)."""
    b = sc.single("soc_work_obst", n_people=0, ceres_compat=220)
    b.arrays["has_people"][:] = 1
    x = b.arrays["u0"][0].ravel()
    e = oracle.evaluate(b, 0, x, want_jac=False)
    kinds, _ = oracle.layout(b)
    assert e["ok"]
    assert np.all(e["residuals"][kinds == K_PROX] == 0.0)
    soc = e["residuals"][kinds == K_SOCIAL]
    assert np.all(soc > 120.0 * 1e-6)
    ej = oracle.evaluate(b, 0, x, want_jac=True)
    assert not ej["ok"]
    out = oracle.solve_batch(b)
    assert out["usable"][0] == 0 and out["termination"][0] == 6


def test_obstacle_residual_is_bicubic_of_costmap(oracle):
    b = sc.corridor(B=1)
    x = b.arrays["u0"][0].ravel()
    e = oracle.evaluate(b, 0, x, want_jac=False)
    kinds, steps = oracle.layout(b)
    # roll the unicycle out by hand (reference update_state.hpp:46-61)
    ch, bl, nb, _ = b.dims
    px, py, th = b.arrays["pose0"][0]
    for k in np.nonzero(kinds == K_OBSTACLE)[0]:
        i = steps[k]
        blk = i // bl if i < ch else (ch - 1) // bl
        px += x[2 * blk] * math.cos(th) * b.dt
        py += x[2 * blk] * math.sin(th) * b.dt
        th += x[2 * blk + 1] * b.dt
        fx, fy = px + 0.25 * math.cos(th), py + 0.25 * math.sin(th)
        f, _, _ = oracle.bicubic(b.arrays["costmaps"][0], fy / b.resolution, fx / b.resolution)
        assert e["residuals"][k] == pytest.approx(0.13 * f, rel=1e-12, abs=1e-12)


def test_bicubic_interpolates_grid_nodes_and_clamps(oracle):
    rng = np.random.default_rng(1)
    cm = rng.integers(0, 255, (12, 9)).astype(np.uint8)
    for r in range(12):
        for c in range(9):
            f, _, _ = oracle.bicubic(cm, float(r), float(c))
            assert f == float(cm[r, c])
    f, dr, dc = oracle.bicubic(cm, -7.3, 100.2)  # far outside: every tap clamps to the corner value
    assert f == float(cm[0, 8]) and dr == 0.0 and dc == 0.0
    # derivative vs finite differences inside a cell
    r, c, h = 4.37, 3.81, 1e-6
    f, dr, dc = oracle.bicubic(cm, r, c)
    assert dr == pytest.approx((oracle.bicubic(cm, r + h, c)[0] - oracle.bicubic(cm, r - h, c)[0]) / (2 * h), rel=1e-6)
    assert dc == pytest.approx((oracle.bicubic(cm, r, c + h)[0] - oracle.bicubic(cm, r, c - h)[0]) / (2 * h), rel=1e-6)


@pytest.mark.parametrize("name,n_people", [("readme", 3), ("soc_work_obst", 2), ("obst_only", 0)])
def test_residuals_and_jet_jacobian_against_50_digit_mpmath(oracle, name, n_people):
    """tests/critics_mpmath.py restates every functor from the formulas of SURVEY Appendix D in 50-digit
    arithmetic (Jacobian by central differences with a 1e-20 step). The oracle's double residuals and its Jet<4>
    Jacobian must agree with it to double round-off — for the differentiable semantics, i.e. Ceres >= 2.1
    (under Ceres 2.0.0 the proxemics rows of a differentiated evaluation are the constant 3 w: checked separately)."""
    from tests import critics_mpmath as cm
    b = sc.single(name, n_people=n_people, seed_offset=2, ceres_compat=220) if name != "obst_only" else \
        sc.corridor(B=1, ceres_compat=220)
    rng = np.random.default_rng(21)
    P = 2 * b.n_blocks
    x = b.arrays["u0"][0].ravel() + rng.normal(0, 0.04, P)
    x[0::2] = np.clip(x[0::2], 0.05, 0.55)
    e = oracle.evaluate(b, 0, x)
    assert e["ok"]
    prob = cm.problem_from_batch(b, 0)
    r_mp = np.array([float(v) for v in cm.residuals(prob, x)])
    assert r_mp.size == e["residuals"].size
    assert np.allclose(e["residuals"], r_mp, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(r_mp).max()))
    J_mp = np.array([[float(v) for v in row] for row in cm.jacobian(prob, x)])
    scale = max(1.0, np.abs(J_mp).max())
    assert np.abs(e["jac"] - J_mp).max() <= 1e-10 * scale, np.abs(e["jac"] - J_mp).max() / scale

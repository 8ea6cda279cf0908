"""CPU-only tests of the oracle's trust-region LM restatement: LM invariants on the iteration
trace, the line-search polynomial minimiser against numpy, and the minimum against scipy's bounded TRF solver."""
import numpy as np
import pytest

from nav2_social_mpc_controller_b200 import scenarios as sc


def test_trace_invariants(oracle):
    for name in ("readme", "soc_work_obst", "params_yaml"):
        b = sc.single(name)
        tr = oracle.solve_trace(b)
        t = tr["trace"]
        it, cost, succ, radius = t[:, 0], t[:, 1], t[:, 9].astype(bool), t[:, 6]
        assert np.array_equal(it, np.arange(len(it)))
        acc = cost[succ]
        assert np.all(np.diff(acc) <= 0.0), "accepted-step costs must be monotone (monotonic TR)"
        assert tr["cost_final"] == pytest.approx(cost.min(), rel=0, abs=0)
        assert tr["cost_initial"] == cost[0]
        assert len(it) - 1 <= b.params.max_iterations
        # rejected steps shrink the radius by 2, 4, 8, ... (StepRejected), accepted ones follow the rho formula
        for k in range(1, len(it)):
            if not succ[k]:
                assert radius[k] < radius[k - 1]
            else:
                rho = t[k, 5]
                want = min(1e16, radius[k - 1] / max(1.0 / 3.0, 1.0 - (2 * rho - 1) ** 3))
                assert radius[k] == pytest.approx(want, rel=1e-12)
        ls = t[1:, 7]
        assert np.all((ls == -1.0) | ((ls > 0) & (ls <= 1.0)))


def test_bounds_are_respected_and_unbounded_block_exists(oracle):
    """ch = 13, bl = 6 -> three blocks, only two bounded."""
    b = sc.single("readme")
    out = oracle.solve_batch(b)
    u = out["u"][0]
    assert np.all(u[:2, 0] >= 0) and np.all(u[:2, 0] <= 0.6) and np.all(np.abs(u[:2, 1]) <= 1.4)
    assert b.dims == (13, 6, 3, 2)


def test_poly_roots_and_minimiser_against_numpy(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        c = rng.normal(size=5)
        got = np.sort(oracle.poly_roots(c))
        want = np.sort(np.roots(c).real)
        assert np.allclose(got, want, rtol=1e-8, atol=1e-8)
    # cubic through (0, f0, g0), (t, f1, g1): minimiser on [1e-3 t, 0.6 t]
    for _ in range(50):
        f0, g0, t = rng.uniform(1, 10), -rng.uniform(0.1, 5), rng.uniform(0.1, 1)
        f1, g1 = f0 + rng.uniform(0, 5), rng.uniform(-5, 20)
        lo, hi = 1e-3 * t, 0.6 * t
        x = oracle.poly_min([[0, f0, g0, 1, 1], [t, f1, g1, 1, 1]], lo, hi)
        A = np.array([[0, 0, 0, 1], [0, 0, 1, 0], [t ** 3, t ** 2, t, 1], [3 * t ** 2, 2 * t, 1, 0]], dtype=float)
        coef = np.linalg.solve(A, [f0, g0, f1, g1])
        grid = np.linspace(lo, hi, 20001)
        assert lo <= x <= hi
        assert np.polyval(coef, x) <= np.polyval(coef, grid).min() + 1e-9 * abs(f0)


def test_interior_solutions_are_local_minima_for_scipy_trf(oracle):
    """Independent check of the MINIMUM (not of the iterate path): restarting scipy's bounded
    trust-region-reflective solver from an oracle solution that has no active bound must not lower the cost.
    (Solutions sitting on a bound are excluded: Ceres' project-the-LM-step handling of bounds is known to stall
    there with a large projected gradient, and the restatement reproduces that.)"""
    from scipy.optimize import least_squares
    b = sc.corridor(B=48, velocity_feasibility_w=5.0)
    b.params.fn_tol, b.params.max_iterations = 1e-13, 400
    out = oracle.solve_batch(b)
    P = 2 * b.n_blocks
    nbd = b.dims[3]
    lo = np.full(P, -np.inf)
    hi = np.full(P, np.inf)
    lo[0:2 * nbd:2], hi[0:2 * nbd:2] = 0.0, 0.6
    lo[1:2 * nbd:2], hi[1:2 * nbd:2] = -1.4, 1.4
    checked = 0
    for k in range(b.n_problems):
        x = out["u"][k].ravel()
        if np.any(x - lo < 1e-6) or np.any(hi - x < 1e-6):
            continue

        def fun(z):
            return oracle.evaluate(b, k, z, want_jac=False)["residuals"]

        def jac(z):
            return oracle.evaluate(b, k, z, want_jac=True)["jac"]
        r = least_squares(fun, x, jac=jac, bounds=(lo, hi), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12,
                          max_nfev=500)
        assert r.cost >= out["cost_final"][k] * (1 - 2e-3), (k, r.cost, out["cost_final"][k])
        checked += 1
    assert checked >= 2


def test_ceres_compat_switch_changes_only_early_termination(oracle):
    b = sc.corridor(B=16)
    a = oracle.solve_batch(b)
    b.params.ceres_compat = 220
    c = oracle.solve_batch(b)
    assert np.all(c["iterations"] >= a["iterations"])
    assert np.all(c["cost_final"] <= a["cost_final"] * (1 + 1e-12))


def test_post_solve_expansion(oracle):
    """reference src/optimizer.cpp:390-446: cmds[S+1] hold block i/bl for i < ch then the last block; the path is
    the Euler rollout of those cmds (pose0 excluded)."""
    b = sc.single("readme")
    out = oracle.solve_batch(b)
    ch, bl, nb, nbd = b.dims
    u, cmds, path = out["u"][0], out["cmds"][0], out["path"][0]
    S = b.n_steps
    for i in range(S + 1):
        blk = i // bl if i < ch else nb - 1
        assert np.array_equal(cmds[i], u[min(blk, nb - 1)])
    x, y, th = b.arrays["pose0"][0]
    for i in range(S + 1):
        x, y = x + cmds[i, 0] * np.cos(th) * b.dt, y + cmds[i, 0] * np.sin(th) * b.dt
        th = th + cmds[i, 1] * b.dt
        assert path[i, 0] == pytest.approx(x, abs=1e-12) and path[i, 1] == pytest.approx(y, abs=1e-12)
        assert np.cos(path[i, 2] - th) == pytest.approx(1.0, abs=1e-12)


def _numpy_trlm_on(oracle, batch, b):
    """Run tests/trlm_numpy.py (the second, numpy TR-LM restatement) on problem b with the oracle's functor evaluation
    as a black box."""
    from tests import trlm_numpy as tn
    each = batch.arrays.get("n_steps_each")
    S_b = batch.n_steps if each is None else int(each[b])
    ch, bl, nb, nbd = sc.abi.problem_dims(batch.params.control_horizon, batch.params.parameter_block_length, S_b)
    P = 2 * nb
    lo, hi = np.full(P, -np.inf), np.full(P, np.inf)
    lo[0:2 * nbd:2], hi[0:2 * nbd:2] = 0.0, 0.6
    lo[1:2 * nbd:2], hi[1:2 * nbd:2] = -1.4, 1.4

    def evaluate(x, differentiated):
        e = oracle.evaluate(batch, b, x, want_jac=differentiated)
        return e if e["ok"] else None

    p = batch.params
    opt = tn.Options(max_iterations=p.max_iterations, function_tolerance=p.fn_tol, gradient_tolerance=p.gradient_tol,
                     parameter_tolerance=p.param_tol, ceres_compat=p.ceres_compat or 200)
    return tn.solve(evaluate, batch.arrays["u0"][b].ravel()[:P], lo, hi, opt), P


def _oracle_eval_rows(oracle, batch, b):
    import ctypes as C
    rows = np.zeros((512, 8))
    st = batch.struct()
    oracle.lib.smpc_oracle_solve_evals.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    oracle.lib.smpc_oracle_solve_evals.restype = C.c_int
    n = oracle.lib.smpc_oracle_solve_evals(C.byref(batch.params), C.byref(st), b, rows.ctypes.data, 512)
    return rows[:n]


def test_independent_numpy_trlm_reproduces_the_oracle_trace_on_the_golden_cases(oracle):
    """tests/trlm_numpy.py is a second TR-LM restatement with a different structure (numpy QR on
    the augmented system instead of Cholesky on the normal equations, Vandermonde solve + numpy.roots = companion
    eigenvalues for the line-search polynomial, one class per Ceres component). Fed with the same functor evaluations
    it must walk the same sequence of trial points as oracle/solver.hpp — same phases, same step sizes, same costs,
    same termination, same iteration count, same solution — on every golden problem (both Ceres versions, mixed
    horizons included)."""
    from tests import golden_lib
    n_checked = 0
    for name, (batch, gold, _) in sorted(golden_lib.load().items()):
        for b in range(min(batch.n_problems, 4)):
            res, P = _numpy_trlm_on(oracle, batch, b)
            ref = _oracle_eval_rows(oracle, batch, b)
            assert res["termination"] == gold["termination"][b], (name, b)
            assert res["iterations"] == gold["iterations"][b], (name, b)
            assert bool(res["usable"]) == bool(gold["usable"][b])
            assert np.abs(res["x"] - gold["u"][b].ravel()[:P]).max() <= 1e-7, (name, b)  # QR vs Cholesky round-off
            assert res["cost_final"] == pytest.approx(gold["cost_final"][b], rel=1e-10)
            rows = res["rows"]
            assert len(rows) == len(ref), (name, b, len(rows), len(ref))
            for (it, phase, t, c_diff, c_plain), r in zip(rows, ref):
                assert it == r[0] and phase == r[1], (name, b)
                assert t == pytest.approx(r[2], rel=1e-5, abs=1e-12), (name, b, it)
                if c_diff is not None and not np.isnan(r[3]):
                    assert c_diff == pytest.approx(r[3], rel=1e-8)
                if c_plain is not None and not np.isnan(r[4]):
                    assert c_plain == pytest.approx(r[4], rel=1e-8)
            n_checked += 1
    assert n_checked >= 20


def test_perturbed_oracle_variants_build_and_agree_on_well_conditioned_problems(oracle):
    """tools/oracle_sensitivity.py measures the noise floor of the restated algorithm with three variants of the SAME
    oracle source (FMA contraction, double instead of long-double line-search polynomial, reverse summation order). They
    must build, and on well-conditioned people-free problems they must land on the oracle's own result."""
    import ctypes as C
    import os
    import subprocess
    from tests import oracle_lib
    here = os.path.dirname(os.path.abspath(__file__))
    odir = os.path.join(os.path.dirname(here), "oracle")
    names = ["liboracle_fma.so", "liboracle_polydouble.so", "liboracle_revsum.so"]
    subprocess.run(["make", "-C", odir, "-s"] + names, check=True)
    batch = sc.corridor(B=24)
    base = oracle.solve_batch(batch, n_threads=4)
    for name in names:
        other = oracle_lib.Oracle(C.CDLL(os.path.join(odir, name))).solve_batch(batch, n_threads=4)
        assert np.array_equal(other["termination"], base["termination"]), name
        assert np.array_equal(other["iterations"], base["iterations"]), name
        assert np.abs(other["u"] - base["u"]).max() <= 1e-6, name

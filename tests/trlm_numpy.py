"""Second, independently structured restatement of the bounded trust-region Levenberg-Marquardt loop that
ceres::Solve runs for the reference's options (src/optimizer.cpp:117-131) — TEST INFRASTRUCTURE.

Written from SURVEY.md Appendix A in numpy, on purpose NOT following oracle/solver.hpp: dense Jacobian algebra through
numpy (QR solve of the augmented least-squares system [Js; D] y = [r; 0] instead of Cholesky on the normal equations),
the line-search polynomial fitted with numpy.linalg.solve on the Vandermonde system and minimised through numpy.roots
(companion-matrix eigenvalues, which is what ceres/polynomial.cc does), and an explicit object per Ceres component
(strategy, step evaluator, line search) instead of one loop. The residual functors themselves are a black box here
(`evaluate(x, differentiated)`); tests/test_oracle_solver.py feeds it the C++ oracle's evaluation and requires the two
TR-LM restatements to produce the same iterate trace.
"""
from __future__ import annotations

import dataclasses

import numpy as np

DBL_MAX = np.finfo(np.float64).max

# enum smpc_termination
CONV_GRADIENT, CONV_PARAMETER, CONV_FUNCTION, CONV_RADIUS, NO_CONVERGENCE, FAIL_INVALID, FAIL_EVAL = range(7)


@dataclasses.dataclass
class Options:
    max_iterations: int = 100
    function_tolerance: float = 1e-7
    gradient_tolerance: float = 1e-10
    parameter_tolerance: float = 1e-15
    ceres_compat: int = 200
    # Ceres defaults the reference never touches
    initial_radius: float = 1e4
    max_radius: float = 1e16
    min_radius: float = 1e-32
    min_relative_decrease: float = 1e-3
    min_lm_diagonal: float = 1e-6
    max_lm_diagonal: float = 1e32
    max_consecutive_invalid_steps: int = 5
    sufficient_decrease: float = 1e-4
    max_step_contraction: float = 1e-3
    min_step_contraction: float = 0.6
    max_line_search_iterations: int = 20
    min_line_search_step: float = 1e-9


class Box:
    """ParameterBlock::Plus with bounds: x + delta, clamped to [lo, hi] per coordinate."""

    def __init__(self, lo, hi):
        self.lo, self.hi = np.asarray(lo, float), np.asarray(hi, float)

    def plus(self, x, delta):
        return np.minimum(np.maximum(x + delta, self.lo), self.hi)


class LevenbergMarquardt:
    """ceres LevenbergMarquardtStrategy: radius bookkeeping and the damped step."""

    def __init__(self, opt: Options):
        self.o = opt
        self.radius = opt.initial_radius
        self.decrease_factor = 2.0
        self.reuse = False
        self.diagonal = None

    def step(self, Js, r):
        if not self.reuse:
            self.diagonal = np.clip((Js * Js).sum(axis=0), self.o.min_lm_diagonal, self.o.max_lm_diagonal)
        self.reuse = True
        D = np.sqrt(self.diagonal / self.radius)
        A = np.vstack([Js, np.diag(D)])
        b = np.concatenate([r, np.zeros(Js.shape[1])])
        try:
            y = np.linalg.lstsq(A, b, rcond=None)[0]
        except np.linalg.LinAlgError:
            return None
        if not np.all(np.isfinite(y)):
            return None
        return -y

    def accepted(self, rho):
        self.radius = min(self.o.max_radius, self.radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
        self.decrease_factor = 2.0
        self.reuse = False

    def rejected(self):
        self.radius /= self.decrease_factor
        self.decrease_factor *= 2.0
        self.reuse = True


class StepEvaluator:
    """ceres TrustRegionStepEvaluator with max_consecutive_nonmonotonic_steps = 0 (monotonic)."""

    def __init__(self, initial_cost):
        self.current = initial_cost

    def quality(self, cost, model_cost_change):
        if cost >= DBL_MAX:
            return -DBL_MAX
        return (self.current - cost) / model_cost_change

    def accepted(self, cost):
        self.current = cost


def _fit(samples):
    """Interpolating polynomial (highest power first) through samples [(x, value or None, gradient or None)]."""
    n = sum((v is not None) + (g is not None) for _, v, g in samples)
    deg = n - 1
    A, b = [], []
    for x, v, g in samples:
        if v is not None:
            A.append([x ** (deg - j) for j in range(deg + 1)])
            b.append(v)
        if g is not None:
            A.append([(deg - j) * x ** (deg - j - 1) if j < deg else 0.0 for j in range(deg + 1)])
            b.append(g)
    return np.linalg.solve(np.array(A, float), np.array(b, float))


def minimize_interpolating_polynomial(samples, lo, hi):
    """ceres MinimizeInterpolatingPolynomial: midpoint, both ends, real parts of the derivative's roots inside
    [lo, hi], then any sample abscissa inside [lo, hi] whose polynomial value is lower."""
    poly = _fit(samples)
    val = lambda t: float(np.polyval(poly, t))
    best_x = 0.5 * (lo + hi)
    best_v = val(best_x)
    for t in (lo, hi):
        if val(t) < best_v:
            best_x, best_v = t, val(t)
    d = np.polyder(poly)
    d = np.trim_zeros(d, "f")
    if d.size >= 2:
        for root in np.roots(d):
            t = float(np.real(root))
            if lo <= t <= hi and val(t) < best_v:
                best_x, best_v = t, val(t)
    for x, _, _ in samples:
        if lo <= x <= hi and val(x) < best_v:
            best_x, best_v = x, val(x)
    return best_x


class ArmijoLineSearch:
    """ceres ArmijoLineSearch with CUBIC interpolation along the projected direction."""

    def __init__(self, opt: Options, evaluate, box: Box):
        self.o, self.evaluate, self.box = opt, evaluate, box

    def search(self, x, direction, cost0, slope0):
        o = self.o
        dmax = float(np.max(np.abs(direction)))
        previous = None
        t = 1.0
        iters = 0
        rows = []
        while True:
            e = self.evaluate(self.box.plus(x, t * direction), True)
            ok_v = e is not None and np.isfinite(e["cost"])
            value = e["cost"] if ok_v else None
            slope = float(direction @ e["grad"]) if ok_v else None
            if slope is not None and not np.isfinite(slope):
                slope = None
            passed = ok_v and not (value > cost0 + o.sufficient_decrease * slope0 * t)
            rows.append((t, value, passed))
            if passed:
                return t, iters, rows
            iters += 1
            if iters >= o.max_line_search_iterations:
                return None, iters, rows
            lo, hi = o.max_step_contraction * t, o.min_step_contraction * t
            if not ok_v:
                t_new = min(max(0.5 * t, lo), hi)
            else:
                samples = [(0.0, cost0, slope0), (t, value, slope)]
                if previous is not None:
                    samples.append(previous)
                t_new = minimize_interpolating_polynomial(samples, lo, hi)
            if t_new * dmax < o.min_line_search_step:
                return None, iters, rows
            previous = (t, value, slope) if ok_v else None
            t = t_new


def solve(evaluate, x0, lo, hi, opt: Options):
    """evaluate(x, differentiated) -> dict(cost, residuals, jac, grad) (jac / grad only when differentiated) or None
    when the evaluation is invalid. Returns dict(x, termination, iterations, cost_initial, cost_final, usable, rows)
    with rows = one (iteration, phase, t, differentiated cost, plain cost) tuple per trial point."""
    box = Box(lo, hi)
    x = box.plus(np.asarray(x0, float), 0.0)
    rows = []
    e = evaluate(x, True)
    if e is None:
        return dict(x=np.asarray(x0, float), termination=FAIL_EVAL, iterations=0, usable=False, rows=rows,
                    cost_initial=np.nan, cost_final=np.nan)
    scale = 1.0 / (1.0 + np.sqrt((e["jac"] ** 2).sum(axis=0)))

    def at_new_point(ev):
        g = ev["grad"]
        return ev["cost"], ev["residuals"], ev["jac"] * scale, g, float(np.max(np.abs(x - box.plus(x, -g))))

    cost, r, Js, g, gmax = at_new_point(e)
    rows.append((0, 1, 0.0, cost, None))
    cost_initial = final_cost = cost
    strategy = LevenbergMarquardt(opt)
    quality = StepEvaluator(cost)
    line_search = ArmijoLineSearch(opt, evaluate, box)
    best_x, minimum_cost = x.copy(), DBL_MAX
    x_norm = float(np.linalg.norm(x))
    iteration, n_invalid = 0, 0
    successful, it_cost, any_success = True, cost, False

    def done(term):
        usable = term <= NO_CONVERGENCE
        return dict(x=best_x if usable else np.asarray(x0, float), termination=term, iterations=iteration, usable=usable,
                    rows=rows, cost_initial=cost_initial, cost_final=final_cost)

    while True:
        if successful and cost < minimum_cost:
            minimum_cost, best_x = cost, x.copy()
        final_cost = min(final_cost, it_cost)
        if iteration >= opt.max_iterations:
            return done(NO_CONVERGENCE)
        if successful and gmax <= opt.gradient_tolerance:
            return done(CONV_GRADIENT)
        if strategy.radius <= opt.min_radius:
            return done(CONV_RADIUS)
        iteration += 1

        s = strategy.step(Js, r)
        model_change = None
        if s is not None:
            m = Js @ s
            model_change = float(-(m @ (r + m / 2.0)))
        if s is None or not (model_change > 0.0):
            n_invalid += 1
            if n_invalid >= opt.max_consecutive_invalid_steps:
                iteration -= 1
                return done(FAIL_INVALID)
            strategy.rejected()
            successful, it_cost = False, cost
            continue
        n_invalid = 0
        delta = s * scale

        t, _, ls_rows = line_search.search(x, delta, cost, float(g @ delta))
        for (tt, vv, _) in ls_rows:
            rows.append((iteration, 2, tt, vv, None))
        if t is not None:
            delta = t * delta
        else:
            rows.append((iteration, 3, 1.0, None, None))

        cand = box.plus(x, delta)
        ec = evaluate(cand, False)
        cand_cost = ec["cost"] if ec is not None else DBL_MAX
        rows[-1] = rows[-1][:4] + (cand_cost,)
        armed = opt.ceres_compat < 210 or any_success
        if armed and np.linalg.norm(x - cand) <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance):
            iteration -= 1
            return done(CONV_PARAMETER)
        if armed and abs(cost - cand_cost) <= opt.function_tolerance * cost:
            iteration -= 1
            return done(CONV_FUNCTION)
        rho = quality.quality(cand_cost, model_change)
        if rho > opt.min_relative_decrease:
            x = cand
            x_norm = float(np.linalg.norm(x))
            e = evaluate(x, True)
            if e is None:
                iteration -= 1
                return done(FAIL_EVAL)
            cost, r, Js, g, gmax = at_new_point(e)
            rows[-1] = rows[-1][:3] + (cost, cand_cost)
            strategy.accepted(rho)
            quality.accepted(cand_cost)
            successful, it_cost, any_success = True, cost, True
        else:
            strategy.rejected()
            successful, it_cost = False, cand_cost

// TEST INFRASTRUCTURE — CPU oracle for the social-MPC solve path. Not shipped, not on the product path.
//
// solver.hpp — (1) the residual-block list Optimizer::optimize assembles (src/optimizer.cpp:251-379),
// (2) a ProgramEvaluator-shaped evaluation (cost = 1/2 sum r^2, gradient = J^T r, dense J) in which every
// residual block re-rolls-out the trajectory under Jet<4> stride passes like DynamicAutoDiffCostFunction,
// and (3) a restatement of the bounded trust-region Levenberg-Marquardt loop ceres::Solve runs for the
// options set at src/optimizer.cpp:117-131 (all other options: Ceres defaults).
//
// Ceres itself (libceres-dev, unpinned: 2.0.0 / 2.2.0) is not in /root/reference nor installable here; the
// algorithm below restates its published trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
// line_search.cc (ArmijoLineSearch), polynomial.cc and parameter_block.h (bounds projection in Plus).
// PARITY UNPINNED (no reference tests or golden vectors exist) — see critics.hpp.
#pragma once
#include <cfloat>
#include <cmath>
#include <complex>
#include <limits>
#include <vector>

#include "critics.hpp"

namespace smpc_oracle {

enum Kind { K_AGENT_ANGLE = 0, K_SOCIAL, K_PROX, K_VELOCITY, K_GOAL, K_PATH_FOLLOW, K_PATH_ALIGN, K_OBSTACLE, K_VEL_FEAS };

struct ResidualBlock {
  Kind kind;
  int step;    // i of src/optimizer.cpp:251
  int nblk;    // number of 2-wide parameter blocks this cost function was given
  int blk[2];  // K_VEL_FEAS only: parameter blocks (i, i-1), src/optimizer.cpp:368-369
};

// src/optimizer.cpp:251-371, in AddResidualBlock order.
inline std::vector<ResidualBlock> assemble(const ProblemView& p) {
  std::vector<ResidualBlock> out;
  out.reserve(static_cast<size_t>(p.S) * 9);
  for (int i = 0; i < p.S; ++i) {
    const int nblk = p.blocks_seen(i);
    if (p.has_people) {
      out.push_back({K_AGENT_ANGLE, i, nblk, {0, 0}});
      out.push_back({K_SOCIAL, i, nblk, {0, 0}});
      out.push_back({K_PROX, i, nblk, {0, 0}});
    }
    out.push_back({K_VELOCITY, i, nblk, {0, 0}});
    out.push_back({K_GOAL, i, nblk, {0, 0}});
    out.push_back({K_PATH_FOLLOW, i, nblk, {0, 0}});
    out.push_back({K_PATH_ALIGN, i, nblk, {0, 0}});
    out.push_back({K_OBSTACLE, i, nblk, {0, 0}});
    if (i != 0 && i < p.ch / p.bl) out.push_back({K_VEL_FEAS, i, 2, {i, i - 1}});
  }
  return out;
}

template <class T>
inline T eval_block(const ProblemView& p, const ResidualBlock& rb, const T* const* u) {
  switch (rb.kind) {
    case K_AGENT_ANGLE: return agent_angle_residual<T>(p, u, rb.step);
    case K_SOCIAL: return social_work_residual<T>(p, u, rb.step);
    case K_PROX: return proxemics_residual<T>(p, u, rb.step);
    case K_VELOCITY: return velocity_residual<T>(p, u, rb.step);
    case K_GOAL: return goal_align_residual<T>(p, u, rb.step);
    case K_PATH_FOLLOW: return distance_residual<T>(p, u, rb.step, p.w_distance, p.px[p.S], p.py[p.S]);
    case K_PATH_ALIGN: return distance_residual<T>(p, u, rb.step, p.w_angle, p.px[rb.step + 1], p.py[rb.step + 1]);
    case K_OBSTACLE: return obstacle_residual<T>(p, u, rb.step);
    case K_VEL_FEAS: return vel_feasibility_residual<T>(p, u[0], u[1]);
  }
  return T(0.0);
}

struct EvalCounters {
  long n_jacobian = 0;  // evaluations that differentiated (Jacobian or gradient requested)
  long n_cost = 0;      // residual-only evaluations
};

constexpr int kMaxParams = 64;

// Evaluate the whole program at x[P]. residuals[m], gradient[P], jac[m*P] (row-major) are optional.
// Returns false when any residual / Jacobian entry is non-finite (ceres ResidualBlock::Evaluate validity check).
inline bool evaluate(const ProblemView& p, const std::vector<ResidualBlock>& blocks, const double* x, double* cost,
                     double* residuals, double* gradient, double* jac, EvalCounters* cnt) {
  const int D = p.dof, P = D * p.nb;
  const bool need_d = (gradient != nullptr) || (jac != nullptr);
  if (cnt) (need_d ? cnt->n_jacobian : cnt->n_cost)++;
  double total = 0.0;
  if (gradient)
    for (int c = 0; c < P; ++c) gradient[c] = 0.0;
  using J4 = Jet<4>;
  for (size_t kk = 0; kk < blocks.size(); ++kk) {
#ifdef SMPC_ORACLE_REVERSE_SUM  // tools/oracle_sensitivity.py: cost and gradient summed over the residual blocks in the
    const size_t k = blocks.size() - 1 - kk;  // opposite order — an equally valid rounding of the same sums (Ceres' own
#else                                         // order depends on its thread count)
    const size_t k = kk;
#endif
    const ResidualBlock& rb = blocks[k];
    int gidx[kMaxParams];  // local parameter -> global column
    const int np = D * rb.nblk;
    if (rb.kind == K_VEL_FEAS) {
      for (int k = 0; k < D; ++k) {
        gidx[k] = D * rb.blk[0] + k;
        gidx[D + k] = D * rb.blk[1] + k;
      }
    } else {
      for (int l = 0; l < np; ++l) gidx[l] = l;
    }
    double r = 0.0;
    double row[kMaxParams];
    if (!need_d) {
      const double* ptr[kMaxParams / 2];
      for (int b = 0; b < rb.nblk; ++b) ptr[b] = x + gidx[D * b];
      r = eval_block<double>(p, rb, ptr);
      if (!std::isfinite(r)) return false;
    } else {
      // DynamicAutoDiffCostFunction: ceil(np/4) passes, 4 tangent directions each.
      J4 jets[kMaxParams];
      const J4* ptr[kMaxParams / 2];
      for (int b = 0; b < rb.nblk; ++b) ptr[b] = jets + D * b;
      for (int start = 0; start < np; start += 4) {
        for (int l = 0; l < np; ++l) jets[l] = J4(x[gidx[l]], l - start);
        J4 out = eval_block<J4>(p, rb, ptr);
        r = out.a;
        for (int l = start; l < np && l < start + 4; ++l) row[l] = out.v[l - start];
      }
      if (!std::isfinite(r)) return false;
      for (int l = 0; l < np; ++l)
        if (!std::isfinite(row[l])) return false;
      if (jac) {
        double* jr = jac + k * P;
        for (int c = 0; c < P; ++c) jr[c] = 0.0;
        for (int l = 0; l < np; ++l) jr[gidx[l]] = row[l];
      }
      if (gradient)
        for (int l = 0; l < np; ++l) gradient[gidx[l]] += row[l] * r;
    }
    if (residuals) residuals[k] = r;
    total += 0.5 * r * r;
  }
  *cost = total;
  return true;
}

// ---------------------------------------------------------------------------------------------
// polynomial.cc restatement (interpolating polynomial + its minimiser on an interval)
// ---------------------------------------------------------------------------------------------
struct Sample {
  double x = 0, value = 0, gradient = 0;
  bool value_ok = false, gradient_ok = false;
};

// Precision note. Ceres fits the interpolating polynomial by solving the Vandermonde-type system with Eigen's
// FullPivLU in double and finds the critical points as eigenvalues of the companion matrix. With line-search samples
// at t ~ 1e-3 .. 1 that system is badly conditioned: the minimiser Ceres returns carries rounding noise of 1e-7 .. 1e-5
// relative that depends on Eigen's exact operation order and is not reproducible by ANY other implementation (measured:
// tests/trlm_numpy.py — numpy.linalg.solve + numpy.roots in double — differs from a double full-pivot LU by exactly that
// much, and the differences feed back into the iterate path). The oracle therefore computes the SAME polynomial and the
// SAME candidate set (midpoint, end points, real parts of the derivative's roots, in-range samples) in long double, i.e.
// the noise-free value Ceres' result scatters around; parity checks then measure the solver under test, not the noise.
#ifdef SMPC_ORACLE_POLY_DOUBLE  // tools/oracle_sensitivity.py: the same polynomial code in double, as Ceres computes it
using real = double;
#else
using real = long double;
#endif

inline real poly_eval(const std::vector<real>& c, real x) {  // Horner, highest degree first
  real v = 0.0L;
  for (real ck : c) v = v * x + ck;
  return v;
}

// Full-pivot LU solve (Eigen::FullPivLU with threshold 0 in Ceres).
inline std::vector<real> solve_full_pivot(std::vector<std::vector<real>> a, std::vector<real> b) {
  const int n = static_cast<int>(b.size());
  std::vector<int> colperm(n);
  for (int i = 0; i < n; ++i) colperm[i] = i;
  for (int k = 0; k < n; ++k) {
    int pr = k, pc = k;
    real best = -1.0L;
    for (int i = k; i < n; ++i)
      for (int j = k; j < n; ++j)
        if (std::fabs(a[i][j]) > best) {
          best = std::fabs(a[i][j]);
          pr = i;
          pc = j;
        }
    if (best == 0.0L) break;
    std::swap(a[k], a[pr]);
    std::swap(b[k], b[pr]);
    if (pc != k) {
      for (int i = 0; i < n; ++i) std::swap(a[i][k], a[i][pc]);
      std::swap(colperm[k], colperm[pc]);
    }
    for (int i = k + 1; i < n; ++i) {
      const real f = a[i][k] / a[k][k];
      if (f == 0.0L) continue;
      for (int j = k; j < n; ++j) a[i][j] -= f * a[k][j];
      b[i] -= f * b[k];
    }
  }
  std::vector<real> y(n, 0.0L);
  for (int i = n - 1; i >= 0; --i) {
    real s = b[i];
    for (int j = i + 1; j < n; ++j) s -= a[i][j] * y[j];
    y[i] = (a[i][i] != 0.0L) ? s / a[i][i] : 0.0L;
  }
  std::vector<real> xsol(n, 0.0L);
  for (int i = 0; i < n; ++i) xsol[colperm[i]] = y[i];
  return xsol;
}

inline std::vector<real> interpolating_polynomial(const std::vector<Sample>& s) {
  int nc = 0;
  for (const Sample& q : s) nc += (q.value_ok ? 1 : 0) + (q.gradient_ok ? 1 : 0);
  const int degree = nc - 1;
  std::vector<std::vector<real>> lhs(nc, std::vector<real>(nc, 0.0L));
  std::vector<real> rhs(nc, 0.0L);
  int row = 0;
  for (const Sample& q : s) {
    const real x = q.x;
    if (q.value_ok) {
      for (int j = 0; j <= degree; ++j) lhs[row][j] = std::pow(x, static_cast<real>(degree - j));
      rhs[row++] = q.value;
    }
    if (q.gradient_ok) {
      for (int j = 0; j < degree; ++j) lhs[row][j] = (degree - j) * std::pow(x, static_cast<real>(degree - j - 1));
      rhs[row++] = q.gradient;
    }
  }
  return solve_full_pivot(lhs, rhs);
}

// Real parts of all roots. Degree <= 2 in closed form as Ceres does (FindLinear/QuadraticPolynomialRoots);
// higher degrees: Ceres takes eigenvalues of the balanced companion matrix — here Aberth-Ehrlich in long double,
// an independent method converging to the same roots.
inline std::vector<real> real_parts_of_roots(std::vector<real> c) {
  size_t lead = 0;
  while (lead + 1 < c.size() && c[lead] == 0.0L) ++lead;
  c.erase(c.begin(), c.begin() + lead);
  const int degree = static_cast<int>(c.size()) - 1;
  std::vector<real> roots;
  if (degree <= 0) return roots;
  if (degree == 1) {
    roots.push_back(-c[1] / c[0]);
    return roots;
  }
  if (degree == 2) {
    const real a = c[0], b = c[1], cc = c[2];
    const real D = b * b - 4 * a * cc;
    const real sD = std::sqrt(std::fabs(D));
    if (D >= 0) {
      if (b >= 0) {
        roots.push_back((-b - sD) / (2.0L * a));
        roots.push_back((2.0L * cc) / (-b - sD));
      } else {
        roots.push_back((2.0L * cc) / (-b + sD));
        roots.push_back((-b + sD) / (2.0L * a));
      }
    } else {
      roots.push_back(-b / (2.0L * a));
      roots.push_back(-b / (2.0L * a));
    }
    return roots;
  }
  using cld = std::complex<long double>;
  std::vector<long double> m(degree + 1);
  for (int i = 0; i <= degree; ++i) m[i] = c[i] / c[0];
  long double radius = 0.0L;
  for (int i = 1; i <= degree; ++i) radius = std::max(radius, std::pow(std::fabs(m[i]), 1.0L / i));
  radius = 2.0L * radius + 1e-30L;
  std::vector<cld> z(degree);
  for (int i = 0; i < degree; ++i) z[i] = std::polar(radius * 0.7L, 2.0L * 3.14159265358979323846L * i / degree + 0.35L);
  long double previous_moved = 1.0L;
  for (int it = 0; it < 500; ++it) {
    long double moved = 0.0L;
    for (int i = 0; i < degree; ++i) {
      cld pv = m[0], dv = 0.0L;
      for (int j = 1; j <= degree; ++j) {
        dv = dv * z[i] + pv;
        pv = pv * z[i] + m[j];
      }
      if (std::abs(pv) == 0.0L) continue;
      cld ratio = pv / dv;
      cld sum = 0.0L;
      for (int j = 0; j < degree; ++j)
        if (j != i) sum += cld(1.0L) / (z[i] - z[j]);
      cld step = ratio / (cld(1.0L) - ratio * sum);
      z[i] -= step;
      moved = std::max(moved, std::abs(step) / (std::abs(z[i]) + 1e-300L));
    }
    // converged, or down in the rounding noise of p(z) / p'(z) (a few long-double epsilons relative to |z|, more for
    // clustered roots) where the steps stop shrinking. (Until round 2 the test was `moved < 1e-19` alone — below the
    // long-double epsilon, so every call ran all 500 sweeps: 10x the CPU time of a people-free solve.)
    if (moved < 1e-18L || (moved < 1e-13L && moved >= previous_moved)) break;
    previous_moved = moved;
  }
  for (int i = 0; i < degree; ++i) roots.push_back(z[i].real());
  return roots;
}

inline void minimize_polynomial(const std::vector<real>& poly, real x_min, real x_max, real* opt_x, real* opt_v) {
  *opt_x = (x_min + x_max) / 2.0L;
  *opt_v = poly_eval(poly, *opt_x);
  const real vmin = poly_eval(poly, x_min);
  if (vmin < *opt_v) {
    *opt_v = vmin;
    *opt_x = x_min;
  }
  const real vmax = poly_eval(poly, x_max);
  if (vmax < *opt_v) {
    *opt_v = vmax;
    *opt_x = x_max;
  }
  if (poly.size() <= 2) return;
  const int degree = static_cast<int>(poly.size()) - 1;
  std::vector<real> d(degree);
  for (int j = 0; j < degree; ++j) d[j] = (degree - j) * poly[j];
  for (real root : real_parts_of_roots(d)) {
    if (root < x_min || root > x_max) continue;
    const real v = poly_eval(poly, root);
    if (v < *opt_v) {
      *opt_v = v;
      *opt_x = root;
    }
  }
}

inline double minimize_interpolating_polynomial(const std::vector<Sample>& s, double x_min, double x_max) {
  const std::vector<real> poly = interpolating_polynomial(s);
  real ox, ov;
  minimize_polynomial(poly, x_min, x_max, &ox, &ov);
  for (const Sample& q : s) {
    if (q.x < x_min || q.x > x_max) continue;
    const real v = poly_eval(poly, q.x);
    if (v < ov) {
      ov = v;
      ox = q.x;
    }
  }
  return static_cast<double>(ox);
}

// ---------------------------------------------------------------------------------------------
// Trust-region LM
// ---------------------------------------------------------------------------------------------
struct SolveOptions {
  int max_iterations = 100;
  double fn_tol = 1e-7, gradient_tol = 1e-10, param_tol = 1e-15;
  int ceres_compat = 200;  // 200: tolerance tests from the first iteration; >= 210: only after a successful step
  // Ceres defaults
  double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
  double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  int max_consecutive_invalid_steps = 5;
  double ls_sufficient_decrease = 1e-4, ls_max_contraction = 1e-3, ls_min_contraction = 0.6, ls_min_step = 1e-9;
  int ls_max_iterations = 20;
};

enum Termination {
  T_CONVERGENCE_GRADIENT = 0,
  T_CONVERGENCE_PARAMETER = 1,
  T_CONVERGENCE_FUNCTION = 2,
  T_CONVERGENCE_RADIUS = 3,
  T_NO_CONVERGENCE = 4,
  T_FAILURE_INVALID_STEPS = 5,
  T_FAILURE_EVALUATION = 6
};

struct IterationRecord {
  int iteration;
  double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius, line_search_t;
  bool valid, successful;
};

// One row per distinct trial point, in evaluation order (what tools/flip_log.py aligns with the GPU solver's trace):
// phase 1 = iteration zero, 2 = line-search sample, 3 = un-shortened step after a failed line search.
// code: bit 0 = the sample passed the Armijo test, bit 1 = the point became the candidate, bit 2 = step accepted,
// bit 3 = the solve terminated on this row, bits 4.. = termination type (when bit 3 is set).
// aux: rejected line-search sample -> cost - (cost(x) + 1e-4 g0 t) (> 0 = rejected); candidate -> relative decrease.
struct EvalRecord {
  double iteration, phase, t, cost_diff, cost_plain, aux, code, radius;
};

struct SolveSummary {
  std::vector<EvalRecord> eval_rows;
  double initial_cost = 0, final_cost = 0;
  int termination = T_NO_CONVERGENCE;
  bool usable = false;
  int iterations = 0;
  int num_successful = 0, num_unsuccessful = 0, num_line_search_steps = 0;
  EvalCounters evals;
  std::vector<IterationRecord> trace;
};

struct Bounds {
  double lo[kMaxParams], hi[kMaxParams];
};

// ParameterBlock::Plus with box projection.
inline void plus(const Bounds& bd, int P, const double* x, const double* delta, double* out) {
  for (int c = 0; c < P; ++c) {
    double v = x[c] + delta[c];
    v = std::max(v, bd.lo[c]);
    v = std::min(v, bd.hi[c]);
    out[c] = v;
  }
}

// src/optimizer.cpp:373-379: only the first ch/bl blocks are bounded.
inline Bounds make_bounds(const ProblemView& p) {
  Bounds b;
  for (int c = 0; c < kMaxParams; ++c) {
    b.lo[c] = -std::numeric_limits<double>::infinity();
    b.hi[c] = std::numeric_limits<double>::infinity();
  }
  const int D = p.dof;
  for (int k = 0; k < p.n_bounded; ++k) {
    b.lo[D * k] = 0.0;
    b.hi[D * k] = 0.6;
    if (D == 3) {  // omnidirectional extension: the lateral velocity is bounded by the same speed
      b.lo[D * k + 1] = -0.6;
      b.hi[D * k + 1] = 0.6;
    }
    b.lo[D * k + D - 1] = -1.4;
    b.hi[D * k + D - 1] = 1.4;
  }
  return b;
}

// Dense Cholesky (LLT) solve of A y = b, A symmetric P x P row-major. false on a non-positive pivot (Eigen LLT).
inline bool cholesky_solve(int P, std::vector<double> A, const double* b, double* y) {
  for (int j = 0; j < P; ++j) {
    double d = A[j * P + j];
    for (int k = 0; k < j; ++k) d -= A[j * P + k] * A[j * P + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[j * P + j] = d;
    for (int i = j + 1; i < P; ++i) {
      double s = A[i * P + j];
      for (int k = 0; k < j; ++k) s -= A[i * P + k] * A[j * P + k];
      A[i * P + j] = s / d;
    }
  }
  std::vector<double> z(P);
  for (int i = 0; i < P; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i * P + k] * z[k];
    z[i] = s / A[i * P + i];
  }
  for (int i = P - 1; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < P; ++k) s -= A[k * P + i] * y[k];
    y[i] = s / A[i * P + i];
  }
  return true;
}

// ceres::Solve for this problem. x: in = initial block values, out = solution (only written when usable,
// like Solver: parameters are copied back only if IsSolutionUsable()).
inline SolveSummary solve(const ProblemView& p, const SolveOptions& opt, double* x_inout, bool keep_trace = false) {
  SolveSummary sum;
  const std::vector<ResidualBlock> blocks = assemble(p);
  const int P = p.dof * p.nb;
  const int m = static_cast<int>(blocks.size());
  const Bounds bd = make_bounds(p);

  std::vector<double> x(x_inout, x_inout + P), best(x_inout, x_inout + P), cand(P), delta(P), step(P);
  std::vector<double> residuals(m), gradient(P), jac(static_cast<size_t>(m) * P), scale(P, 1.0), diag(P), zero(P, 0.0);
  std::vector<double> neg_g(P), proj(P), model_res(m);

  // IterationZero: project the start point onto the box.
  plus(bd, P, x.data(), zero.data(), cand.data());
  x = cand;
  double x_norm = 0.0;
  for (double v : x) x_norm += v * v;
  x_norm = std::sqrt(x_norm);

  double x_cost = std::numeric_limits<double>::max();
  double minimum_cost = x_cost;
  double gmax = 0.0;
  int iteration = 0;

  auto eval_grad_jac = [&](bool first) -> bool {
    if (!evaluate(p, blocks, x.data(), &x_cost, residuals.data(), gradient.data(), jac.data(), &sum.evals)) return false;
    if (first) {
      for (int c = 0; c < P; ++c) {
        double s2 = 0.0;
        for (int k = 0; k < m; ++k) s2 += jac[static_cast<size_t>(k) * P + c] * jac[static_cast<size_t>(k) * P + c];
        scale[c] = 1.0 / (1.0 + std::sqrt(s2));
      }
    }
    for (int k = 0; k < m; ++k)
      for (int c = 0; c < P; ++c) jac[static_cast<size_t>(k) * P + c] *= scale[c];
    for (int c = 0; c < P; ++c) neg_g[c] = -gradient[c];
    plus(bd, P, x.data(), neg_g.data(), proj.data());
    gmax = 0.0;
    for (int c = 0; c < P; ++c) gmax = std::max(gmax, std::fabs(x[c] - proj[c]));
    return true;
  };

  if (!eval_grad_jac(true)) {
    sum.termination = T_FAILURE_EVALUATION;
    sum.usable = false;
    sum.initial_cost = sum.final_cost = x_cost;
    return sum;
  }
  sum.initial_cost = x_cost;
  sum.final_cost = x_cost;

  double radius = opt.initial_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int n_invalid = 0;
  bool it_successful = true;
  double it_cost = x_cost;
  bool any_success = false;
  IterationRecord rec{0, x_cost, 0, gmax, 0, 0, radius, 0, true, true};

  // TrustRegionStepEvaluator::current_cost_ (trust_region_step_evaluator.cc): the cost the step quality is measured
  // from. It starts as the iteration-zero cost and, after an accepted step, becomes that step's CANDIDATE cost — the
  // cost-only (double) evaluation — not the cost of the differentiated re-evaluation x_cost. The two are the same
  // number unless a functor evaluates differently under Jets (proxemics with Ceres 2.0.0, critics.hpp).
  double se_cost = x_cost;
  if (keep_trace) sum.eval_rows.push_back({0.0, 1.0, 0.0, x_cost, NAN, NAN, 0.0, opt.initial_radius});
  auto finish = [&](int term) {
    if (keep_trace && !sum.eval_rows.empty()) {
      EvalRecord& e = sum.eval_rows.back();
      e.code = static_cast<double>((static_cast<int>(e.code) | 8) + 16 * term);
      e.radius = radius;
    }
    sum.termination = term;
    sum.usable = (term <= T_NO_CONVERGENCE);
    sum.iterations = iteration;
    if (sum.usable)
      for (int c = 0; c < P; ++c) x_inout[c] = best[c];
    return sum;
  };

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it_successful) {
      ++sum.num_successful;
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        best = x;
      }
    } else {
      ++sum.num_unsuccessful;
    }
    sum.final_cost = std::min(sum.final_cost, it_cost);
    rec.radius = radius;
    if (keep_trace && !sum.eval_rows.empty()) sum.eval_rows.back().radius = radius;
    if (keep_trace) sum.trace.push_back(rec);
    if (iteration >= opt.max_iterations) return finish(T_NO_CONVERGENCE);
    if (it_successful && gmax <= opt.gradient_tol) return finish(T_CONVERGENCE_GRADIENT);
    if (radius <= opt.min_radius) return finish(T_CONVERGENCE_RADIUS);

    ++iteration;
    rec = IterationRecord{iteration, 0, 0, gmax, 0, 0, radius, 0, false, false};

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int c = 0; c < P; ++c) {
        double s2 = 0.0;
        for (int k = 0; k < m; ++k) s2 += jac[static_cast<size_t>(k) * P + c] * jac[static_cast<size_t>(k) * P + c];
        diag[c] = std::min(std::max(s2, opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
    }
    std::vector<double> lhs(static_cast<size_t>(P) * P, 0.0), rhs(P, 0.0);
    for (int k = 0; k < m; ++k) {
      const double* jr = &jac[static_cast<size_t>(k) * P];
      for (int a = 0; a < P; ++a) {
        if (jr[a] == 0.0) continue;
        rhs[a] += jr[a] * residuals[k];
        for (int b2 = 0; b2 < P; ++b2) lhs[a * P + b2] += jr[a] * jr[b2];
      }
    }
    for (int c = 0; c < P; ++c) {
      const double lm = std::sqrt(diag[c] / radius);
      lhs[c * P + c] += lm * lm;
    }
    reuse_diagonal = true;
    bool step_ok = cholesky_solve(P, lhs, rhs.data(), step.data());
    if (step_ok)
      for (int c = 0; c < P; ++c) step_ok = step_ok && std::isfinite(step[c]);
    bool valid = false;
    double model_cost_change = 0.0;
    if (step_ok) {
      for (int c = 0; c < P; ++c) step[c] = -step[c];
      // model_cost_change = -(J s)'(r + J s / 2)
      for (int k = 0; k < m; ++k) {
        double v = 0.0;
        for (int c = 0; c < P; ++c) v += jac[static_cast<size_t>(k) * P + c] * step[c];
        model_res[k] = v;
      }
      for (int k = 0; k < m; ++k) model_cost_change -= model_res[k] * (residuals[k] + model_res[k] / 2.0);
      valid = model_cost_change > 0.0;
    }
    rec.valid = valid;
    if (!valid) {
      // HandleInvalidStep
      if (++n_invalid >= opt.max_consecutive_invalid_steps) {
        --iteration;  // Minimize() returns before this iteration is recorded
        return finish(T_FAILURE_INVALID_STEPS);
      }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      it_successful = false;
      it_cost = x_cost;
      rec.cost = x_cost;
      continue;
    }
    for (int c = 0; c < P; ++c) delta[c] = step[c] * scale[c];
    n_invalid = 0;

    // DoLineSearch: projected Armijo with cubic interpolation (the problem is bounds constrained).
    {
      const std::vector<double> direction = delta;
      double g0 = 0.0, dmax = 0.0;
      for (int c = 0; c < P; ++c) {
        g0 += gradient[c] * direction[c];
        dmax = std::max(dmax, std::fabs(direction[c]));
      }
      Sample initial;
      initial.x = 0.0;
      initial.value = x_cost;
      initial.gradient = g0;
      initial.value_ok = initial.gradient_ok = true;
      Sample previous, current;
      std::vector<double> sd(P), xt(P), gt(P);
      auto ls_eval = [&](double t, Sample* out) {
        *out = Sample();
        out->x = t;
        for (int c = 0; c < P; ++c) sd[c] = t * direction[c];
        plus(bd, P, x.data(), sd.data(), xt.data());
        double v;
        if (!evaluate(p, blocks, xt.data(), &v, nullptr, gt.data(), nullptr, &sum.evals) || !std::isfinite(v)) return;
        out->value = v;
        out->value_ok = true;
        double gd = 0.0;
        for (int c = 0; c < P; ++c) gd += direction[c] * gt[c];
        out->gradient = gd;
        if (!std::isfinite(gd)) return;
        out->gradient_ok = true;
      };
      int ls_iters = 0;
      bool ls_success = false;
      ls_eval(1.0, &current);
      for (;;) {
        const double armijo_margin = current.value - (x_cost + opt.ls_sufficient_decrease * g0 * current.x);
        if (keep_trace)
          sum.eval_rows.push_back({static_cast<double>(iteration), 2.0, current.x, current.value_ok ? current.value : NAN, NAN,
                               armijo_margin, 0.0, radius});
        if (current.value_ok && !(current.value > x_cost + opt.ls_sufficient_decrease * g0 * current.x)) {
          ls_success = true;
          if (keep_trace) sum.eval_rows.back().code = 1.0;
          break;
        }
        ++ls_iters;
        if (ls_iters >= opt.ls_max_iterations) break;
        const double lo = opt.ls_max_contraction * current.x, hi = opt.ls_min_contraction * current.x;
        double t_new;
        if (!current.value_ok) {
          t_new = std::min(std::max(current.x * 0.5, lo), hi);
        } else {
          std::vector<Sample> s{initial, current};
          if (previous.value_ok) s.push_back(previous);
          t_new = minimize_interpolating_polynomial(s, lo, hi);
        }
        if (t_new * dmax < opt.ls_min_step) break;
        previous = current;
        ls_eval(t_new, &current);
      }
      sum.num_line_search_steps += ls_iters;
      if (ls_success) {
        for (int c = 0; c < P; ++c) delta[c] *= current.x;
        rec.line_search_t = current.x;
      } else {
        rec.line_search_t = -1.0;
        if (keep_trace) sum.eval_rows.push_back({static_cast<double>(iteration), 3.0, 1.0, NAN, NAN, NAN, 0.0, radius});
      }
    }

    // ComputeCandidatePointAndEvaluateCost
    plus(bd, P, x.data(), delta.data(), cand.data());
    double cand_cost;
    if (!evaluate(p, blocks, cand.data(), &cand_cost, nullptr, nullptr, nullptr, &sum.evals))
      cand_cost = std::numeric_limits<double>::max();
    if (keep_trace) {
      EvalRecord& e = sum.eval_rows.back();
      e.cost_plain = cand_cost;
      e.code = static_cast<double>(static_cast<int>(e.code) | 2);
    }

    const bool tol_armed = (opt.ceres_compat < 210) || any_success;
    // ParameterToleranceReached
    double step_norm = 0.0;
    for (int c = 0; c < P; ++c) step_norm += (x[c] - cand[c]) * (x[c] - cand[c]);
    step_norm = std::sqrt(step_norm);
    rec.step_norm = step_norm;
    if (tol_armed && step_norm <= opt.param_tol * (x_norm + opt.param_tol)) {
      if (keep_trace) sum.eval_rows.back().aux = step_norm / (opt.param_tol * (x_norm + opt.param_tol)) - 1.0;
      --iteration;  // this iteration is not recorded by Ceres
      return finish(T_CONVERGENCE_PARAMETER);
    }
    // FunctionToleranceReached
    const double cost_change = x_cost - cand_cost;
    rec.cost_change = cost_change;
    if (tol_armed && std::fabs(cost_change) <= opt.fn_tol * x_cost) {
      if (keep_trace) sum.eval_rows.back().aux = std::fabs(cost_change) - opt.fn_tol * x_cost;
      --iteration;
      return finish(T_CONVERGENCE_FUNCTION);
    }
    // IsStepSuccessful
    double rho;
    if (cand_cost >= std::numeric_limits<double>::max())
      rho = std::numeric_limits<double>::lowest();
    else
      rho = (se_cost - cand_cost) / model_cost_change;  // TrustRegionStepEvaluator::StepQuality (monotonic steps)
    rec.relative_decrease = rho;
    if (keep_trace) sum.eval_rows.back().aux = rho;
    if (rho > opt.min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      x_norm = 0.0;
      for (double v : x) x_norm += v * v;
      x_norm = std::sqrt(x_norm);
      if (!eval_grad_jac(false)) {
        --iteration;
        return finish(T_FAILURE_EVALUATION);
      }
      any_success = true;
      it_successful = true;
      it_cost = x_cost;
      se_cost = cand_cost;  // step_evaluator_->StepAccepted(candidate_cost_, model_cost_change_)
      if (keep_trace) {
        EvalRecord& e = sum.eval_rows.back();
        e.cost_diff = x_cost;
        e.code = static_cast<double>(static_cast<int>(e.code) | 4);
      }
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(opt.max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      rec.successful = true;
      rec.cost = x_cost;
      rec.gradient_max_norm = gmax;
    } else {
      it_successful = false;
      it_cost = cand_cost;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      rec.successful = false;
      rec.cost = cand_cost;
    }
  }
}

}  // namespace smpc_oracle

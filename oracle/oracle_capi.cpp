// TEST INFRASTRUCTURE — CPU oracle for the social-MPC solve path. Not shipped, not on the product path.
//
// oracle_capi.cpp — extern "C" entry points of liboracle.so, bound with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs. It consumes the same
// smpc_params / smpc_batch / smpc_result PODs as libsmpc.so (include/smpc.h) so parity tests hand identical
// buffers to both.  PARITY UNPINNED — see critics.hpp.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/smpc.h"
#include "solver.hpp"

#ifdef SMPC_ORACLE_COUNT_OPS
namespace smpc_oracle {
thread_local unsigned long long g_jet_flops = 0;
}
#endif

using namespace smpc_oracle;

namespace {

bool derive_dims(const smpc_params* p, int S, int* ch, int* bl, int* nb, int* n_bounded) {
  if (S < 1 || p->control_horizon < 1 || p->parameter_block_length < 1) return false;
  *ch = std::min(p->control_horizon, S);           // src/optimizer.cpp:248
  *bl = std::min(p->parameter_block_length, *ch);  // src/optimizer.cpp:249
  *nb = (*ch + *bl - 1) / *bl;
  *n_bounded = *ch / *bl;
  return true;
}

// Row width of u0 / u: the parameter blocks of the LONGEST horizon of the batch (a shorter problem uses the first
// nb_b of them, the rest stay untouched).
int batch_blocks(const smpc_params* p, const smpc_batch* in) {
  int ch, bl, nb, nbd;
  if (!derive_dims(p, in->n_steps, &ch, &bl, &nb, &nbd)) return 0;
  return nb;
}

bool make_view(const smpc_params* p, const smpc_batch* in, int b, ProblemView* v) {
  const int S1 = in->n_steps + 1, A = in->n_agents;
  const int S = in->n_steps_each ? in->n_steps_each[b] : in->n_steps;  // per-problem horizon (include/smpc.h)
  if (S < 1 || S > in->n_steps) return false;
  v->S = S;
  v->S1 = S1;
  v->A = A;
  v->ceres_compat = p->ceres_compat ? p->ceres_compat : 200;
  v->dof = p->omni_solve ? 3 : 2;
  if (!derive_dims(p, S, &v->ch, &v->bl, &v->nb, &v->n_bounded)) return false;
  v->dt = in->dt;
  v->x0 = in->pose0[3 * b + 0];
  v->y0 = in->pose0[3 * b + 1];
  v->yaw0 = in->pose0[3 * b + 2];
  v->u0 = in->u0 + static_cast<size_t>(b) * batch_blocks(p, in) * (p->omni_solve ? 3 : 2);
  v->px = in->path_xy + static_cast<size_t>(b) * 2 * S1;
  v->py = v->px + S1;
  v->goal_yaw = in->goal_yaw[b];
  v->agents = (A > 0 && in->agents) ? in->agents + static_cast<size_t>(b) * A * 6 * S1 : nullptr;
  v->has_people = in->has_people ? (in->has_people[b] != 0) : false;
  if (v->agents == nullptr) v->A = 0;
  const int mi = in->costmap_index ? in->costmap_index[b] : (in->n_costmaps > 0 ? b % in->n_costmaps : 0);
  v->map = in->costmaps + static_cast<size_t>(mi) * in->size_x * in->size_y;
  v->size_x = in->size_x;
  v->size_y = in->size_y;
  v->origin_x = in->costmap_origin[2 * mi];
  v->origin_y = in->costmap_origin[2 * mi + 1];
  v->resolution = in->resolution;
  v->w_distance = p->distance_w;
  v->w_social = p->socialwork_w;
  v->w_velocity = p->velocity_w;
  v->w_angle = p->angle_w;
  v->w_agent_angle = p->agent_angle_w;
  v->w_prox = p->proxemics_w;
  v->w_vf = p->velocity_feasibility_w;
  v->w_obstacle = p->obstacle_w;
  v->w_goal = p->goal_align_w;
  return true;
}

SolveOptions make_options(const smpc_params* p) {
  SolveOptions o;
  o.max_iterations = p->max_iterations;
  o.fn_tol = p->fn_tol;
  o.gradient_tol = p->gradient_tol;
  o.param_tol = p->param_tol;
  o.ceres_compat = p->ceres_compat ? p->ceres_compat : 200;
  return o;
}

// tf2 Quaternion::setRPY(0,0,yaw) followed by tf2::getYaw.
double yaw_roundtrip(double yaw) {
  const double h = yaw * 0.5;
  const double qz = std::sin(h), qw = std::cos(h);  // q = (0, 0, sin(yaw/2), cos(yaw/2))
  const double sqw = qw * qw, sqz = qz * qz;
  // tf2 impl::getYaw: sarg = -2(qx qz - qw qy)/|q|^2 = 0 here, so the generic branch is taken.
  return std::atan2(2 * (qw * qz), sqw - sqz);
}

// src/optimizer.cpp:390-446: hold-last fill, block -> per-step expansion, Euler path rebuild. D = v.dof values per
// step: (v, w), or (vx, vy, w) for the omnidirectional extension (same expansion rule, holonomic Euler step).
void expand_outputs(const ProblemView& v, const double* u, double* cmds, double* path) {
  const int S = v.S, D = v.dof;
  // optim_velocities: entries [0, nb) hold the blocks; the reference then overwrites entries
  // [ch/bl, S) with block (ch-1)/bl (this also overwrites an unbounded trailing block's own slot with itself).
  std::vector<double> ov(static_cast<size_t>(S) * D);
  for (int i = 0; i < S; ++i) {
    const int src = (i < v.nb) ? i : 0;
    for (int k = 0; k < D; ++k) ov[D * i + k] = u[D * src + k];
  }
  const int last = (v.ch - 1) / v.bl;
  std::vector<double> lastv(u + D * last, u + D * last + D);
  for (int k = 0; k < D; ++k) lastv[k] = ov[D * last + k];
  for (int i = v.ch / v.bl; i < S; ++i)
    for (int k = 0; k < D; ++k) ov[D * i + k] = lastv[k];
  std::vector<double> sv(static_cast<size_t>(S + 1) * D);
  for (int i = 0; i < v.ch; ++i)
    for (int k = 0; k < D; ++k) sv[D * i + k] = ov[D * (i / v.bl) + k];
  for (int i = v.ch; i < S + 1; ++i)
    for (int k = 0; k < D; ++k) sv[D * i + k] = ov[D * (i - 1) + k];
  if (cmds) std::memcpy(cmds, sv.data(), sv.size() * sizeof(double));
  if (path) {
    double x = v.x0, y = v.y0, yaw = yaw_roundtrip(v.yaw0);
    for (int i = 0; i < S + 1; ++i) {
      const double vx = sv[D * i], vy = (D == 3) ? sv[D * i + 1] : 0.0, w = sv[D * i + D - 1];
      const double nx = (D == 3) ? x + (vx * std::cos(yaw) - vy * std::sin(yaw)) * v.dt : x + vx * std::cos(yaw) * v.dt;
      const double ny = (D == 3) ? y + (vx * std::sin(yaw) + vy * std::cos(yaw)) * v.dt : y + vx * std::sin(yaw) * v.dt;
      const double nyaw = yaw_roundtrip(yaw + w * v.dt);
      x = nx;
      y = ny;
      yaw = nyaw;
      path[3 * i] = x;
      path[3 * i + 1] = y;
      path[3 * i + 2] = yaw;
    }
  }
}

}  // namespace

extern "C" {

int smpc_oracle_num_residuals(const smpc_params* p, const smpc_batch* in, int b) {
  ProblemView v;
  if (!make_view(p, in, b, &v)) return -1;
  return static_cast<int>(assemble(v).size());
}

// Evaluate problem b at x[P]. Optional outputs: residuals[m], grad[P], jac[m*P].
// Returns 1 ok, 0 evaluation invalid (non-finite), <0 bad arguments.
int smpc_oracle_evaluate(const smpc_params* p, const smpc_batch* in, int b, const double* x, double* cost,
                         double* residuals, double* grad, double* jac) {
  ProblemView v;
  if (!make_view(p, in, b, &v)) return -1;
  const std::vector<ResidualBlock> blocks = assemble(v);
  double c = 0.0;
  const bool ok = evaluate(v, blocks, x, &c, residuals, grad, jac, nullptr);
  if (cost) *cost = c;
  return ok ? 1 : 0;
}

// FLOPs of the Jet arithmetic of ONE differentiated evaluation (cost + Jacobian) of problem b at x, as the
// reference-shaped evaluator executes it; -1 unless built with -DSMPC_ORACLE_COUNT_OPS (liboracle_count.so).
long long smpc_oracle_count_jet_flops(const smpc_params* p, const smpc_batch* in, int b, const double* x) {
#ifdef SMPC_ORACLE_COUNT_OPS
  ProblemView v;
  if (!make_view(p, in, b, &v)) return -1;
  const std::vector<ResidualBlock> blocks = assemble(v);
  const int P = v.dof * v.nb;
  std::vector<double> res(blocks.size()), grad(P), jac(blocks.size() * static_cast<size_t>(P));
  double c = 0.0;
  smpc_oracle::g_jet_flops = 0;
  evaluate(v, blocks, x, &c, res.data(), grad.data(), jac.data(), nullptr);
  return static_cast<long long>(smpc_oracle::g_jet_flops);
#else
  (void)p; (void)in; (void)b; (void)x;
  return -1;
#endif
}

// Residual kinds / steps in assembly order (for tests): kinds[m], steps[m].
int smpc_oracle_layout(const smpc_params* p, const smpc_batch* in, int b, int* kinds, int* steps) {
  ProblemView v;
  if (!make_view(p, in, b, &v)) return -1;
  const std::vector<ResidualBlock> blocks = assemble(v);
  for (size_t k = 0; k < blocks.size(); ++k) {
    kinds[k] = blocks[k].kind;
    steps[k] = blocks[k].step;
  }
  return static_cast<int>(blocks.size());
}

static void solve_one(const smpc_params* p, const smpc_batch* in, smpc_result* out, int b) {
  ProblemView v;
  make_view(p, in, b, &v);
  const int P = v.dof * v.nb;
  double x[kMaxParams];
  for (int c = 0; c < P; ++c) x[c] = v.u0[c];
  SolveSummary s = solve(v, make_options(p), x);
  const size_t Pw = static_cast<size_t>(v.dof) * batch_blocks(p, in);  // row width of u (longest horizon of the batch)
  if (out->u) std::memcpy(out->u + static_cast<size_t>(b) * Pw, x, sizeof(double) * P);
  if (out->cost_initial) out->cost_initial[b] = s.initial_cost;
  if (out->cost_final) out->cost_final[b] = s.final_cost;
  if (out->iterations) out->iterations[b] = s.iterations;
  if (out->termination) out->termination[b] = s.termination;
  if (out->usable) out->usable[b] = s.usable ? 1 : 0;
  if (out->n_evals) {
    out->n_evals[2 * b] = static_cast<int32_t>(s.evals.n_jacobian);
    out->n_evals[2 * b + 1] = static_cast<int32_t>(s.evals.n_cost);
  }
  if (out->cmds || out->path)
    expand_outputs(v, x, out->cmds ? out->cmds + static_cast<size_t>(b) * v.S1 * v.dof : nullptr,
                   out->path ? out->path + static_cast<size_t>(b) * v.S1 * 3 : nullptr);
}

// Loop the restated Ceres solve over problems [first, first+count) with n_threads std::threads.
int smpc_oracle_solve_batch(const smpc_params* p, const smpc_batch* in, smpc_result* out, int first, int count,
                            int n_threads) {
  ProblemView v;
  if (in->n_problems <= 0 || !make_view(p, in, 0, &v)) return -1;
  if (v.dof * v.nb > kMaxParams) return -3;
  if (first < 0 || first + count > in->n_problems) return -1;
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next(first);
  auto worker = [&]() {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= first + count) break;
      solve_one(p, in, out, b);
    }
  };
  if (n_threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  return 0;
}

// Per-iteration trace of one solve (debugging / LM-invariant tests). rows of 10 doubles:
// iteration, cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius, line_search_t, valid, successful
int smpc_oracle_solve_trace(const smpc_params* p, const smpc_batch* in, int b, double* x_out, double* trace,
                            int max_rows, int* termination, double* cost_initial, double* cost_final) {
  ProblemView v;
  if (!make_view(p, in, b, &v)) return -1;
  const int P = v.dof * v.nb;
  double x[kMaxParams];
  for (int c = 0; c < P; ++c) x[c] = v.u0[c];
  SolveSummary s = solve(v, make_options(p), x, true);
  for (int c = 0; c < P; ++c) x_out[c] = x[c];
  int rows = 0;
  for (const IterationRecord& r : s.trace) {
    if (rows >= max_rows) break;
    double* t = trace + 10 * rows++;
    t[0] = r.iteration;
    t[1] = r.cost;
    t[2] = r.cost_change;
    t[3] = r.gradient_max_norm;
    t[4] = r.step_norm;
    t[5] = r.relative_decrease;
    t[6] = r.radius;
    t[7] = r.line_search_t;
    t[8] = r.valid;
    t[9] = r.successful;
  }
  if (termination) *termination = s.termination;
  if (cost_initial) *cost_initial = s.initial_cost;
  if (cost_final) *cost_final = s.final_cost;
  return rows;
}

// Per-evaluation trace of one solve (tools/flip_log.py): rows of 8 doubles, see EvalRecord in solver.hpp.
int smpc_oracle_solve_evals(const smpc_params* p, const smpc_batch* in, int b, double* rows, int max_rows) {
  ProblemView v;
  if (!make_view(p, in, b, &v)) return -1;
  const int P = v.dof * v.nb;
  double x[kMaxParams];
  for (int c = 0; c < P; ++c) x[c] = v.u0[c];
  SolveSummary s = solve(v, make_options(p), x, true);
  int n = 0;
  for (const EvalRecord& e : s.eval_rows) {
    if (n >= max_rows) break;
    double* t = rows + 8 * n++;
    t[0] = e.iteration; t[1] = e.phase; t[2] = e.t; t[3] = e.cost_diff;
    t[4] = e.cost_plain; t[5] = e.aux; t[6] = e.code; t[7] = e.radius;
  }
  return n;
}

// polynomial.cc restatement, exposed for unit tests: samples rows of (x, value, gradient, value_ok, gradient_ok).
double smpc_oracle_poly_min(const double* samples, int n, double x_min, double x_max) {
  std::vector<Sample> s(n);
  for (int i = 0; i < n; ++i) {
    s[i].x = samples[5 * i];
    s[i].value = samples[5 * i + 1];
    s[i].gradient = samples[5 * i + 2];
    s[i].value_ok = samples[5 * i + 3] != 0.0;
    s[i].gradient_ok = samples[5 * i + 4] != 0.0;
  }
  return minimize_interpolating_polynomial(s, x_min, x_max);
}

int smpc_oracle_poly_roots(const double* coeffs, int n, double* roots) {
  std::vector<real> r = real_parts_of_roots(std::vector<real>(coeffs, coeffs + n));
  for (size_t i = 0; i < r.size(); ++i) roots[i] = static_cast<double>(r[i]);
  return static_cast<int>(r.size());
}

double smpc_oracle_bicubic(const uint8_t* map, int size_x, int size_y, double r, double c, double* dfdr, double* dfdc) {
  ProblemView v;
  v.map = map;
  v.size_x = size_x;
  v.size_y = size_y;
  double f;
  bicubic(v, r, c, &f, dfdr, dfdc);
  return f;
}

double smpc_oracle_yaw_roundtrip(double yaw) { return yaw_roundtrip(yaw); }

}  // extern "C"

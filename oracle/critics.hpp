// TEST INFRASTRUCTURE — CPU oracle for the social-MPC solve path. Not shipped, not on the product path.
//
// critics.hpp — restatement of the eight residual functors the reference instantiates, of the
// rollout they all share, and of ceres::BiCubicInterpolator<Grid2D<u_char>>. Each function is a
// template over T in {double, Jet<4>} exactly like the reference functors (Ceres autodiff).
// Arithmetic ORDER follows the cited lines so that the jets reproduce Ceres' numbers; structure,
// naming and data layout are this repo's own.
//
// PARITY UNPINNED: the reference ships no tests / golden vectors and neither Ceres nor Eigen nor ROS
// is installable here, so this oracle is pinned only by its own cross-checks
// (tests/test_oracle_*.py: finite differences, mpmath, scipy least_squares, known answers).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>

#include "jet.hpp"

namespace smpc_oracle {

// One problem, viewed through the C-ABI batch layout of include/smpc.h.
struct ProblemView {
  int S = 0;   // optimised steps N_v of THIS problem (src/optimizer.cpp:237,251)
  int S1 = 0;  // row stride of the step-indexed arrays = (largest S of the batch) + 1 (include/smpc.h n_steps_each)
  int A = 0;   // agent columns per step
  int ceres_compat = 200;  // < 210: std::numeric_limits<Jet> is NOT specialised (see proxemics_residual)
  // Parameters per block. 2 = (v, w): the reference (update_state.hpp:46-61). 3 = (vx, vy, w): the omnidirectional
  // EXTENSION of BASELINE configs[4] — the reference has no such solve (only its trajectorizer has an omni branch,
  // path_trajectorizer.hpp:106-123); the functors below are this repo's own definition of it, so for
  // dof == 3 this file is a specification, not a restatement.
  int dof = 2;
  int ch = 0;  // min(control_horizon, S)          src/optimizer.cpp:248
  int bl = 0;  // min(parameter_block_length, ch)  src/optimizer.cpp:249
  int nb = 0;  // ceil(ch/bl) parameter blocks     src/optimizer.cpp:254-261
  int n_bounded = 0;  // ch/bl (integer division)  src/optimizer.cpp:373
  double dt = 0.0;
  double x0 = 0, y0 = 0, yaw0 = 0;
  const double* u0 = nullptr;    // [nb][2]
  const double* px = nullptr;    // [S+1]
  const double* py = nullptr;    // [S+1]
  double goal_yaw = 0.0;
  const double* agents = nullptr;  // [A][6][S+1]
  bool has_people = false;
  const uint8_t* map = nullptr;  // [size_y][size_x]
  int size_x = 0, size_y = 0;
  double origin_x = 0, origin_y = 0, resolution = 1.0;
  double w_distance = 0, w_social = 0, w_velocity = 0, w_angle = 0, w_agent_angle = 0, w_prox = 0, w_vf = 0,
         w_obstacle = 0, w_goal = 0;

  double agent(int step, int k, int c) const { return agents[(static_cast<size_t>(k) * 6 + c) * S1 + step]; }
  int block_of(int j) const { return j < ch ? j / bl : (ch - 1) / bl; }
  int blocks_seen(int i) const { return block_of(i) + 1; }  // src/optimizer.cpp:271-288
};

// update_state.hpp:37-63 — pose after steps 0..i, controls held per block.
template <class T>
inline void rollout(const ProblemView& p, const T* const* u, int i, T* x, T* y, T* th) {
  *x = T(p.x0);
  *y = T(p.y0);
  *th = T(p.yaw0);
  for (int j = 0; j <= i; ++j) {
    const T* blk = u[p.block_of(j)];
    if (p.dof == 3) {  // holonomic Euler step: body velocity (vx, vy) rotated into the world frame
      const T c = cos(*th), s = sin(*th);
      *x += (blk[0] * c - blk[1] * s) * p.dt;
      *y += (blk[0] * s + blk[1] * c) * p.dt;
      *th += blk[2] * p.dt;
    } else {
      *x += blk[0] * cos(*th) * p.dt;
      *y += blk[0] * sin(*th) * p.dt;
      *th += blk[1] * p.dt;
    }
  }
}

template <class T>
struct Vec2 {
  T x, y;
};
template <class T>
inline T sq_norm(const Vec2<T>& a) {
  return a.x * a.x + a.y * a.y;
}
template <class T>
inline T norm2(const Vec2<T>& a) {
  return sqrt(sq_norm(a));
}

// social_work_cost_function.hpp:39-46
template <class T>
inline T wrap_to_pi(T a) {
  while (a > T(M_PI)) a -= T(2.0 * M_PI);
  while (a <= T(-M_PI)) a += T(2.0 * M_PI);
  return a;
}

// One (me <- other) interaction of computeSocialForce, social_work_cost_function.hpp:165-223.
// me/other are 6-vectors (x, y, yaw, t, lv, av); constants from social_work_cost_function.cpp:38-43.
template <class T>
inline Vec2<T> social_pair(const T* me, const Vec2<T>& me_vel, const T* other, const Vec2<T>& o_vel) {
  const double lambda = 2.0, gamma = 0.35, n_prime = 3.0, n = 2.0, factor = 2.1;
  Vec2<T> diff{me[0] - other[0], me[1] - other[1]};
  if (norm2(diff) < T(1e-6)) diff = Vec2<T>{T(1e-6), T(0.0)};
  // Eigen normalized(): divide by sqrt(squaredNorm) only when squaredNorm > 0
  Vec2<T> dir = diff;
  {
    T z = sq_norm(diff);
    if (z > T(0.0)) {
      T s = sqrt(z);
      dir = Vec2<T>{diff.x / s, diff.y / s};
    }
  }
  Vec2<T> dv{me_vel.x - o_vel.x, me_vel.y - o_vel.y};
  Vec2<T> iv{T(lambda) * dv.x + dir.x, T(lambda) * dv.y + dir.y};
  T ilen = norm2(iv);
  Vec2<T> idir{iv.x / ilen, iv.y / ilen};
  T theta = wrap_to_pi(atan2(dir.y, dir.x) - atan2(idir.y, idir.x));
  T B = T(gamma) * ilen;
  T fv = -exp(-norm2(diff) / B - (T(n_prime) * B * theta) * (T(n_prime) * B * theta));
  T sign = (theta > T(0)) ? T(1) : T(-1);
  T fa = -sign * exp(-norm2(diff) / B - (T(n) * B * theta) * (T(n) * B * theta));
  Vec2<T> f_vel{fv * idir.x, fv * idir.y};
  Vec2<T> f_ang{fa * (-idir.y), fa * idir.x};
  return Vec2<T>{T(factor) * (f_vel.x + f_ang.x), T(factor) * (f_vel.y + f_ang.y)};
}

template <class T>
inline void load_agent(const ProblemView& p, int step, int k, T* out) {
  for (int c = 0; c < 6; ++c) out[c] = T(p.agent(step, k, c));
}

template <class T>
inline void robot_state(const ProblemView& p, const T* const* u, int i, T* robot) {
  rollout(p, u, i, &robot[0], &robot[1], &robot[2]);
  robot[3] = T((static_cast<double>(i) + 1.0) * p.dt);  // counter_step, src/optimizer.cpp:253,262
  const T* blk = u[p.block_of(i)];                        // social_work_cost_function.hpp:114-123
  robot[4] = blk[0];
  robot[5] = blk[p.dof - 1];
}

// World-frame velocity of the robot at step i: lv (cos, sin)(yaw) in the reference (social_work_cost_function.hpp:
// 181-183 with me = robot); the omnidirectional extension rotates the body velocity (vx, vy).
template <class T>
inline Vec2<T> robot_velocity(const ProblemView& p, const T* const* u, int i, const T* robot) {
  if (p.dof == 3) {
    const T* blk = u[p.block_of(i)];
    const T c = cos(robot[2]), s = sin(robot[2]);
    return Vec2<T>{blk[0] * c - blk[1] * s, blk[0] * s + blk[1] * c};
  }
  return Vec2<T>{robot[4] * cos(robot[2]), robot[4] * sin(robot[2])};
}

// SocialWorkCost::operator(), social_work_cost_function.hpp:102-150. Agents of step i are
// people_proj[i+1] (src/optimizer.cpp:265).
template <class T>
inline T social_work_residual(const ProblemView& p, const T* const* u, int i) {
  T robot[6];
  robot_state(p, u, i, robot);
  Vec2<T> r_vel = robot_velocity(p, u, i, robot);
  Vec2<T> f_robot{T(0.0), T(0.0)};
  for (int k = 0; k < p.A; ++k) {
    T ag[6];
    load_agent(p, i + 1, k, ag);
    if (ag[3] == T(-1.0)) continue;
    Vec2<T> a_vel0{ag[4] * cos(ag[2]), ag[4] * sin(ag[2])};
    Vec2<T> f = social_pair(robot, r_vel, ag, a_vel0);
    f_robot.x += f.x;
    f_robot.y += f.y;
  }
  T wr = sq_norm(f_robot);
  T wp = T(0.0);
  for (int k = 0; k < p.A; ++k) {  // ALL columns, also padded ones
    T ag[6];
    load_agent(p, i + 1, k, ag);
    Vec2<T> a_vel{ag[4] * cos(ag[2]), ag[4] * sin(ag[2])};
    Vec2<T> f{T(0.0), T(0.0)};
    if (!(robot[3] == T(-1.0))) {
      Vec2<T> g = social_pair(ag, a_vel, robot, r_vel);
      f.x += g.x;
      f.y += g.y;
    }
    wp += sq_norm(f);
  }
  T total = wr + wp + T(1e-6);
  return T(p.w_social) * total;
}

// std::numeric_limits<T>::max() as proxemics_cost_function.hpp:128 sees it. For T = double it is DBL_MAX. For
// T = ceres::Jet the answer depends on the Ceres release: the std::numeric_limits<ceres::Jet<T, N>> specialisation
// was added in Ceres 2.1.0; with Ceres 2.0.0 (Ubuntu 22.04 / ROS 2 Humble, ceres_compat < 210) the PRIMARY template
// answers and returns a value-initialised Jet, i.e. 0 with a zero derivative.
template <class T>
struct IsJet {
  static constexpr bool value = false;
};
template <int N>
struct IsJet<Jet<N>> {
  static constexpr bool value = true;
};

// ProxemicsCost, proxemics_cost_function.hpp:83-151; alpha 3, d0 0.5 (proxemics_cost_function.cpp:37-38).
// Ceres 2.0.0 consequence of the note above: in every DIFFERENTIATED evaluation (Jacobian or gradient requested)
// min_distance starts at Jet(0), std::min keeps it (no squared distance is < 0), and the residual is the constant
// w * alpha * exp(-0) with a zero Jacobian row; cost-only (double) evaluations see the true minimum distance.
template <class T>
inline T proxemics_residual(const ProblemView& p, const T* const* u, int i) {
  T robot[6];
  robot_state(p, u, i, robot);
  T min_d = (IsJet<T>::value && p.ceres_compat < 210) ? T(0.0) : T(std::numeric_limits<double>::max());
  for (int k = 0; k < p.A; ++k) {
    if (p.agent(i + 1, k, 3) == -1.0) continue;
    Vec2<T> diff{robot[0] - T(p.agent(i + 1, k, 0)), robot[1] - T(p.agent(i + 1, k, 1))};
    T d2 = sq_norm(diff);
    min_d = (d2 < min_d) ? d2 : min_d;  // std::min(min_d, d2)
  }
  const double alpha = 3.0, d0 = 0.5;
  T cost = T(alpha) * exp(-min_d / (T(d0) * T(d0)));
  return T(p.w_prox) * cost;
}

// AgentAngleCost, agent_angle_cost_function.hpp:125-195; safe distance^2 = 4 (agent_angle_cost_function.cpp:31).
template <class T>
inline T agent_angle_residual(const ProblemView& p, const T* const* u, int i) {
  T x, y, th;
  rollout(p, u, i, &x, &y, &th);
  int closest = -1;
  double best = std::numeric_limits<double>::infinity();
  for (int k = 0; k < p.A; ++k) {
    const double dx = p.agent(i + 1, k, 0) - p.x0;
    const double dy = p.agent(i + 1, k, 1) - p.y0;
    const double d2 = dx * dx + dy * dy;
    if (d2 < best && p.agent(i + 1, k, 4) > 0.05) {
      best = d2;
      closest = k;
    }
  }
  if (closest < 0 || best > 4.0) return T(0.0);
  const double ax = p.agent(i + 1, closest, 0), ay = p.agent(i + 1, closest, 1), ayaw = p.agent(i + 1, closest, 2);
  T bearing = atan2(T(ay - p.y0), T(ax - p.x0));
  T yaw0 = T(p.yaw0);
  T heading_diff = atan2(sin(T(ayaw) - yaw0), cos(T(ayaw) - yaw0));
  auto wrap = [](const T& a) -> T { return atan2(sin(a), cos(a)); };
  const T k_thr = T(M_PI / 6.0);
  const T k_upper = T(5 * M_PI / 6.0);
  const T steer_right = -T(M_PI / 6.0);
  const T steer_left = T(M_PI / 6.0);
  T ang;
  if (heading_diff <= -k_upper || heading_diff >= k_thr) {
    if (wrap(bearing - yaw0) < 0.0) return T(0.0);
    ang = wrap(th - (yaw0 + steer_right));
  } else {
    if (wrap(bearing - yaw0) > 0.0) return T(0.0);
    ang = wrap(th - (yaw0 + steer_left));
  }
  T c = ang * ang;
  return p.w_agent_angle * c;
}

// VelocityCost, velocity_cost_function.hpp:89-99; desired 0.6 hard-coded at src/optimizer.cpp:238.
template <class T>
inline T velocity_residual(const ProblemView& p, const T* const* u, int i) {
  if (i < p.ch) {
    T d = T(0.6) - u[i / p.bl][0];
    if (p.dof == 3) {  // omnidirectional extension: track 0.6 m/s forward, no lateral velocity
      T vy = u[i / p.bl][1];
      return T(p.w_velocity) * d * d + T(p.w_velocity) * vy * vy;
    }
    return T(p.w_velocity) * d * d;
  }
  return T(0.0);
}

// GoalAlignCost, goal_align_cost_function.hpp:100-116.
template <class T>
inline T goal_align_residual(const ProblemView& p, const T* const* u, int i) {
  T x, y, th;
  rollout(p, u, i, &x, &y, &th);
  T turn = atan2(sin(p.goal_yaw - th), cos(p.goal_yaw - th));
  return T(p.w_goal) * turn * turn;
}

// DistanceCost, distance_cost_function.hpp:117-132.
template <class T>
inline T distance_residual(const ProblemView& p, const T* const* u, int i, double w, double tx, double ty) {
  T x, y, th;
  rollout(p, u, i, &x, &y, &th);
  Vec2<T> d{x - T(tx), y - T(ty)};
  return T(w) * sq_norm(d) * sq_norm(d);
}

// ceres::Grid2D<u_char>::GetValue with clamping + CubicHermiteSpline + BiCubicInterpolator::Evaluate
// (ceres/cubic_interpolation.h; recalled).
inline double grid_value(const ProblemView& p, int r, int c) {
  const int ri = std::min(std::max(0, r), p.size_y - 1);
  const int ci = std::min(std::max(0, c), p.size_x - 1);
  return static_cast<double>(p.map[static_cast<size_t>(ri) * p.size_x + ci]);
}
inline void hermite(double p0, double p1, double p2, double p3, double x, double* f, double* dfdx) {
  const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
  const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
  const double c = 0.5 * (-p0 + p2);
  const double d = p1;
  if (f) *f = d + x * (c + x * (b + x * a));
  if (dfdx) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}
inline void bicubic(const ProblemView& p, double r, double c, double* f, double* dfdr, double* dfdc) {
  const int row = static_cast<int>(std::floor(r));
  const int col = static_cast<int>(std::floor(c));
  double fr[4], dfr[4];
  for (int k = 0; k < 4; ++k) {
    const int rr = row - 1 + k;
    hermite(grid_value(p, rr, col - 1), grid_value(p, rr, col), grid_value(p, rr, col + 1), grid_value(p, rr, col + 2),
            c - col, &fr[k], &dfr[k]);
  }
  hermite(fr[0], fr[1], fr[2], fr[3], r - row, f, dfdr);
  if (dfdc) hermite(dfr[0], dfr[1], dfr[2], dfr[3], r - row, dfdc, nullptr);
}
inline void bicubic_eval(const ProblemView& p, const double& r, const double& c, double* f) {
  bicubic(p, r, c, f, nullptr, nullptr);
}
template <int N>
inline void bicubic_eval(const ProblemView& p, const Jet<N>& r, const Jet<N>& c, Jet<N>* f) {
  double v, dr, dc;
  bicubic(p, r.a, c.a, &v, &dr, &dc);
  f->a = v;
  for (int i = 0; i < N; ++i) f->v[i] = dr * r.v[i] + dc * c.v[i];
}

// ObstacleCost, obstacle_cost_function.hpp:137-167: cost of the point 0.25 m ahead of the robot.
template <class T>
inline T obstacle_residual(const ProblemView& p, const T* const* u, int i) {
  T x, y, th;
  rollout(p, u, i, &x, &y, &th);
  const T off = T(0.25);
  T fx = x + off * cos(th);
  T fy = y + off * sin(th);
  T gx = (fx - T(p.origin_x)) / T(p.resolution);
  T gy = (fy - T(p.origin_y)) / T(p.resolution);
  T val;
  bicubic_eval(p, gy, gx, &val);
  return T(p.w_obstacle) * val;
}

// VelocityFeasibilityCost, velocity_feasibility_cost_function.hpp:86-98 (blocks i and i-1).
template <class T>
inline T vel_feasibility_residual(const ProblemView& p, const T* s1, const T* s2) {
  T dv = s1[0] - s2[0];
  T dw = s1[1] - s2[1];
  if (p.dof == 3) {
    T d2 = s1[2] - s2[2];
    return T(p.w_vf) * dv * dv + T(p.w_vf) * dw * dw + T(p.w_vf) * d2 * d2;
  }
  return T(p.w_vf) * dv * dv + T(p.w_vf) * dw * dw;
}

}  // namespace smpc_oracle

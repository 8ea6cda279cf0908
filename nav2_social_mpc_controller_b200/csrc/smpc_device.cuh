// smpc_device.cuh — device-side social-MPC solver for sm_100a (B200), FP64 CUDA cores.
//
// One WARP solves one MPC problem: lane j owns horizon step j (steps j, j+32 when S > 32).
//   * rollout: heading in closed form from the block-held angular rates, positions by warp inclusive
//     scans; forward sensitivities dX/du, dY/du by the same scans (reference update_state.hpp:37-63
//     re-rolls-out 0..i inside every functor, O(S^2); here it is one O(S) pass per evaluation).
//   * residuals + ANALYTIC gradients wrt the lane's pose (X, Y, Theta, v_block) for the eight active
//     critics (reference include/nav2_social_mpc_controller/critics/*.hpp), accumulated per lane as a
//     4x4 Gauss-Newton block M = sum c c^T and q = sum c r, then H_lane = D^T M D, g_lane = D^T q with
//     the 4xP sensitivity matrix D; warp-shuffle all-reduce gives J^T J, J^T r, cost.
//   * the bounded trust-region Levenberg-Marquardt loop of ceres::Solve (reference src/optimizer.cpp:381
//     with the options of :117-131) runs warp-uniformly: Jacobi scaling, LM diagonal, PxP Cholesky,
//     projected Armijo line search with cubic / quintic interpolation, tolerance tests, radius update.
// No CPU fallback exists; this header is the product path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>
#include <cmath>

namespace smpc {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kMaxChunks = 2;  // S <= 64 steps (the reference parameter sets give 13, 28, 38)

struct DevParams {
  double w_distance, w_social, w_velocity, w_angle, w_agent_angle, w_prox, w_vf, w_obstacle, w_goal;
  double param_tol, fn_tol, gradient_tol;
  int max_iterations, ceres_compat;
  int ch, bl, nb, n_bounded;
};

struct DevBatch {
  int B, S, A, M, size_x, size_y;
  double resolution, dt;
  const double* pose0;
  const double* u0;
  const double* path_xy;
  const double* goal_yaw;
  const double* agents;
  const uint8_t* has_people;
  const uint8_t* costmaps;
  const double* costmap_origin;
  const int32_t* costmap_index;
};

struct DevResult {
  double* u;
  double* cmds;
  double* path;
  double* cost_initial;
  double* cost_final;
  int32_t* iterations;
  int32_t* termination;
  uint8_t* usable;
  int32_t* n_evals;
};

struct DevEvalOut {
  double* cost;
  double* grad;
  double* hess;
  uint8_t* ok;
};

// Warp-uniform view of one problem.
struct Prob {
  double x0, y0, yaw0, goal_yaw, fin_x, fin_y, org_x, org_y;
  const double* px;
  const double* py;
  const double* agents;  // [A][6][S+1]
  const uint8_t* map;
  bool has_people;
};

enum EvalFlags : unsigned { kResidualBad = 1u, kJacobianBad = 2u };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
  return v;
}

// ---------------------------------------------------------------------------------------------------
// Social force of `me` caused by `other` (reference social_work_cost_function.hpp:165-223, constants
// social_work_cost_function.cpp:38-43) as a function of d = me.xy - other.xy and w = me.vel - other.vel,
// with its 2x4 Jacobian wrt (d.x, d.y, w.x, w.y).
// ---------------------------------------------------------------------------------------------------
struct PairOut {
  double fx, fy;
  double dfx[4], dfy[4];
};

__device__ __forceinline__ double wrap_to_pi(double a) {
  // reference social_work_cost_function.hpp:39-46
  while (a > M_PI) a -= 2.0 * M_PI;
  while (a <= -M_PI) a += 2.0 * M_PI;
  return a;
}

__device__ __forceinline__ void social_pair(double dx, double dy, double wx, double wy, PairOut& o) {
  const double kLambda = 2.0, kGamma = 0.35, kNPrime = 3.0, kN = 2.0, kFactor = 2.1;
  double rho = sqrt(dx * dx + dy * dy);
  const bool tiny = rho < 1e-6;
  if (tiny) {  // coincident: fixed direction (1e-6, 0), a constant for the derivative
    dx = 1e-6;
    dy = 0.0;
    rho = sqrt(dx * dx);
  }
  const double inv_rho = 1.0 / rho;
  double ex = dx, ey = dy;
  if (rho * rho > 0.0) {  // Eigen normalized()
    ex = dx * inv_rho;
    ey = dy * inv_rho;
  }
  const double Ix = kLambda * wx + ex, Iy = kLambda * wy + ey;
  const double L = sqrt(Ix * Ix + Iy * Iy);
  const double inv_L = 1.0 / L;
  const double ix = Ix * inv_L, iy = Iy * inv_L;
  const double theta = wrap_to_pi(atan2(ey, ex) - atan2(iy, ix));
  const double Bq = kGamma * L;
  const double inv_B = 1.0 / Bq;
  const double t1 = kNPrime * Bq * theta, t2 = kN * Bq * theta;
  const double base = -rho * inv_B;
  const double E1 = exp(base - t1 * t1);
  const double E2 = exp(base - t2 * t2);
  const double sgn = (theta > 0.0) ? 1.0 : -1.0;
  const double fv = -E1, fa = -sgn * E2;
  o.fx = kFactor * (fv * ix - fa * iy);
  o.fy = kFactor * (fv * iy + fa * ix);

  // gradients wrt (dx, dy, wx, wy)
  const double zd = tiny ? 0.0 : 1.0;
  const double g_rho[2] = {zd * ex, zd * ey};                       // d rho (w part is 0)
  const double g_pe[2] = {-zd * ey * inv_rho, zd * ex * inv_rho};   // d phi_e (w part is 0)
  const double s_ie = -ix * ey + iy * ex;                           // i . e_perp
  const double c_ie = ix * ex + iy * ey;                            // i_perp . e_perp
  double gL[4] = {s_ie * g_pe[0], s_ie * g_pe[1], kLambda * ix, kLambda * iy};
  double gpi[4] = {c_ie * inv_L * g_pe[0], c_ie * inv_L * g_pe[1], -kLambda * iy * inv_L, kLambda * ix * inv_L};
  const double rho_B2 = rho * inv_B * inv_B * kGamma;  // d(-rho/B)/dL
  const double k1 = -2.0 * t1 * kNPrime, k2 = -2.0 * t2 * kN;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double grho = (c < 2) ? g_rho[c] : 0.0;
    const double gpe = (c < 2) ? g_pe[c] : 0.0;
    const double gB = kGamma * gL[c];
    const double gth = gpe - gpi[c];
    const double common = -grho * inv_B + rho_B2 * gL[c];
    const double inner = theta * gB + Bq * gth;
    const double gu1 = common + k1 * inner;
    const double gu2 = common + k2 * inner;
    const double gfv = -E1 * gu1;
    const double gfa = -sgn * E2 * gu2;
    o.dfx[c] = kFactor * (ix * gfv - iy * gfa) - o.fy * gpi[c];
    o.dfy[c] = kFactor * (iy * gfv + ix * gfa) + o.fx * gpi[c];
  }
}

// ---------------------------------------------------------------------------------------------------
// ceres::BiCubicInterpolator<Grid2D<u_char>>: value + d/drow + d/dcol.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hermite(double p0, double p1, double p2, double p3, double x, double& f, double& dfdx) {
  const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
  const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
  const double c = 0.5 * (-p0 + p2);
  f = p1 + x * (c + x * (b + x * a));
  dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

__device__ __forceinline__ void bicubic(const uint8_t* __restrict__ map, int size_x, int size_y, double r, double c,
                                        double& f, double& dfdr, double& dfdc) {
  // clamp the cell index so that a non-finite / huge coordinate cannot overflow the int conversion
  const double rf = floor(fmin(fmax(r, -4.0), (double)size_y + 4.0));
  const double cf = floor(fmin(fmax(c, -4.0), (double)size_x + 4.0));
  const int row = (int)rf, col = (int)cf;
  const double xr = r - (double)row, xc = c - (double)col;
  int cc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) cc[k] = min(max(col - 1 + k, 0), size_x - 1);
  double fr[4], dfr[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = min(max(row - 1 + k, 0), size_y - 1);
    const uint8_t* line = map + (size_t)rr * size_x;
    hermite((double)__ldg(line + cc[0]), (double)__ldg(line + cc[1]), (double)__ldg(line + cc[2]),
            (double)__ldg(line + cc[3]), xc, fr[k], dfr[k]);
  }
  double unused;
  hermite(fr[0], fr[1], fr[2], fr[3], xr, f, dfdr);
  hermite(dfr[0], dfr[1], dfr[2], dfr[3], xr, dfdc, unused);
}

// steps k < j that belong to block beta (blocks hold bl steps, the last one extends to the horizon end)
template <int NB>
__device__ __forceinline__ int steps_in_block_before(int j, int beta, int bl) {
  const int n = j - beta * bl;
  if (beta == NB - 1) return max(n, 0);
  return min(max(n, 0), bl);
}

// AgentAngleCost branch selection (reference agent_angle_cost_function.hpp:125-195): depends only on
// constants of the problem, so it is evaluated once per solve. Returns the steering target angle or NaN
// when the residual is identically zero for this step.
__device__ __forceinline__ double agent_angle_target(const DevBatch& bt, const Prob& pb, int step_plus_1) {
  const int stride = bt.S + 1;
  int closest = -1;
  double best = INFINITY;
  for (int k = 0; k < bt.A; ++k) {
    const double* a = pb.agents + (size_t)k * 6 * stride + step_plus_1;
    const double dx = a[0] - pb.x0, dy = a[stride] - pb.y0;
    const double d2 = dx * dx + dy * dy;
    if (d2 < best && a[4 * stride] > 0.05) {
      best = d2;
      closest = k;
    }
  }
  if (closest < 0 || best > 4.0) return NAN;
  const double* a = pb.agents + (size_t)closest * 6 * stride + step_plus_1;
  const double bearing = atan2(a[stride] - pb.y0, a[0] - pb.x0);
  const double hd_arg = a[2 * stride] - pb.yaw0;
  const double hd = atan2(sin(hd_arg), cos(hd_arg));
  const double wb_arg = bearing - pb.yaw0;
  const double wb = atan2(sin(wb_arg), cos(wb_arg));
  if (hd <= -(5 * M_PI / 6.0) || hd >= (M_PI / 6.0)) {
    if (wb < 0.0) return NAN;
    return pb.yaw0 + (-(M_PI / 6.0));
  }
  if (wb > 0.0) return NAN;
  return pb.yaw0 + (M_PI / 6.0);
}

template <int NB>
struct Normal {  // normal equations of one evaluation, warp-uniform after evaluate()
  static constexpr int P = 2 * NB;
  static constexpr int NH = P * (P + 1) / 2;
  double cost;
  double g[P];
  double H[NH];  // row-major lower triangle, H[a(a+1)/2 + b], a >= b
};

#define SMPC_UNROLL _Pragma("unroll")

// ---------------------------------------------------------------------------------------------------
// Full evaluation at block values x: cost = 1/2 sum r^2, g = J^T r, H = J^T J (all lanes get the result).
// Residual set and order of reference src/optimizer.cpp:251-371.
// ---------------------------------------------------------------------------------------------------
template <int NB>
__device__ __noinline__ unsigned evaluate(const DevParams& prm, const DevBatch& bt, const Prob& pb,
                                          const double (&aa_target)[kMaxChunks], const double (&x)[2 * NB], int lane,
                                          Normal<NB>& out) {
  constexpr int P = 2 * NB;
  constexpr int NH = P * (P + 1) / 2;
  const int S = bt.S, bl = prm.bl, ch = prm.ch;
  const double dt = bt.dt;
  const int stride = S + 1;

  double cost = 0.0;
  double g[P];
  double H[NH];
  SMPC_UNROLL for (int c = 0; c < P; ++c) g[c] = 0.0;
  SMPC_UNROLL for (int e = 0; e < NH; ++e) H[e] = 0.0;
  unsigned flags = 0;

  // carries of the inclusive scans between 32-step chunks
  double carry_x = pb.x0, carry_y = pb.y0;
  double carry_d[4 * NB];
  SMPC_UNROLL for (int e = 0; e < 4 * NB; ++e) carry_d[e] = 0.0;

  for (int chunk = 0; chunk * 32 < S; ++chunk) {
    const int j = chunk * 32 + lane;
    const bool act = j < S;
    const int bj = min(j / bl, NB - 1);
    double vj = x[0], wj = x[1];
    double th = pb.yaw0;  // heading before step j
    SMPC_UNROLL for (int b = 0; b < NB; ++b) {
      if (b == bj) {
        vj = x[2 * b];
        wj = x[2 * b + 1];
      }
      th += x[2 * b + 1] * (dt * (double)steps_in_block_before<NB>(j, b, bl));
    }
    double sn, cs;
    sincos(th, &sn, &cs);
    const double cdt = act ? cs * dt : 0.0, sdt = act ? sn * dt : 0.0;
    const double aj = vj * cdt, bjv = vj * sdt;

    // scan inputs: positions and the 4*NB sensitivities
    double sx = aj, sy = bjv;
    double sd[4 * NB];
    SMPC_UNROLL for (int b = 0; b < NB; ++b) {
      const double tau = dt * (double)steps_in_block_before<NB>(j, b, bl);  // d theta_j / d w_b
      sd[4 * b + 0] = (b == bj) ? cdt : 0.0;                                // dX/dv_b
      sd[4 * b + 1] = (b == bj) ? sdt : 0.0;                                // dY/dv_b
      sd[4 * b + 2] = -bjv * tau;                                           // dX/dw_b
      sd[4 * b + 3] = aj * tau;                                             // dY/dw_b
    }
    SMPC_UNROLL for (int d = 1; d < 32; d <<= 1) {
      const double tx = __shfl_up_sync(kFullMask, sx, d);
      const double ty = __shfl_up_sync(kFullMask, sy, d);
      if (lane >= d) {
        sx += tx;
        sy += ty;
      }
      SMPC_UNROLL for (int e = 0; e < 4 * NB; ++e) {
        const double t = __shfl_up_sync(kFullMask, sd[e], d);
        if (lane >= d) sd[e] += t;
      }
    }
    const double X = carry_x + sx, Y = carry_y + sy;
    SMPC_UNROLL for (int e = 0; e < 4 * NB; ++e) sd[e] += carry_d[e];
    carry_x = __shfl_sync(kFullMask, X, 31);
    carry_y = __shfl_sync(kFullMask, Y, 31);
    SMPC_UNROLL for (int e = 0; e < 4 * NB; ++e) carry_d[e] = __shfl_sync(kFullMask, sd[e], 31);

    if (act) {
      const double Th = th + wj * dt;  // heading after step j
      double sT, cT;
      sincos(Th, &sT, &cT);

      // per-lane Gauss-Newton block wrt (X, Y, Theta, lv): M = sum c c^T (10 entries), q = sum c r
      double mXX = 0, mXY = 0, mXT = 0, mXL = 0, mYY = 0, mYT = 0, mYL = 0, mTT = 0, mTL = 0, mLL = 0;
      double qX = 0, qY = 0, qT = 0, qL = 0;

      if (pb.has_people) {
        // --- AgentAngle (k=0): w * wrap(Theta - target)^2 ------------------------------------------
        const double tgt = aa_target[chunk];
        if (tgt == tgt) {
          const double ad = Th - tgt;
          const double del = atan2(sin(ad), cos(ad));
          const double r = prm.w_agent_angle * (del * del);
          const double cTh = 2.0 * prm.w_agent_angle * del;
          cost += 0.5 * r * r;
          mTT += cTh * cTh;
          qT += cTh * r;
        }
        // --- SocialWork (k=1) and Proxemics (k=2) ----------------------------------------------------
        const double rvx = vj * cT, rvy = vj * sT;  // robot velocity uses lv = v_{b(i)} and the NEW heading
        double Frx = 0.0, Fry = 0.0;
        double JFx[4] = {0, 0, 0, 0}, JFy[4] = {0, 0, 0, 0};
        double wp = 0.0;
        double G4[4] = {0, 0, 0, 0};  // gradient of (wr + wp) wrt (dX, dY, dvx, dvy) of the robot
        double dmin = DBL_MAX, pdx = 0.0, pdy = 0.0;
        const bool do_social = prm.w_social != 0.0;
        for (int k = 0; k < bt.A; ++k) {
          const double* a = pb.agents + (size_t)k * 6 * stride + (j + 1);
          const double ax = a[0], ay = a[stride], at = a[3 * stride];
          const bool valid = !(at == -1.0);
          const double ddx = X - ax, ddy = Y - ay;
          if (valid) {
            const double d2 = ddx * ddx + ddy * ddy;
            if (d2 < dmin) {
              dmin = d2;
              pdx = ddx;
              pdy = ddy;
            }
          }
          if (do_social) {
            const double ayaw = a[2 * stride], alv = a[4 * stride];
            double sa, ca;
            sincos(ayaw, &sa, &ca);
            const double avx = alv * ca, avy = alv * sa;
            PairOut po;
            if (valid) {  // robot <- agent k
              social_pair(ddx, ddy, rvx - avx, rvy - avy, po);
              Frx += po.fx;
              Fry += po.fy;
              SMPC_UNROLL for (int c = 0; c < 4; ++c) {
                JFx[c] += po.dfx[c];
                JFy[c] += po.dfy[c];
              }
            }
            // agent k (also a padded one) <- robot: d and w change sign
            social_pair(-ddx, -ddy, avx - rvx, avy - rvy, po);
            wp += po.fx * po.fx + po.fy * po.fy;
            SMPC_UNROLL for (int c = 0; c < 4; ++c) G4[c] -= 2.0 * (po.fx * po.dfx[c] + po.fy * po.dfy[c]);
          }
        }
        if (do_social) {
          const double wr = Frx * Frx + Fry * Fry;
          SMPC_UNROLL for (int c = 0; c < 4; ++c) G4[c] += 2.0 * (Frx * JFx[c] + Fry * JFy[c]);
          const double r = prm.w_social * (wr + wp + 1e-6);
          const double cX = prm.w_social * G4[0], cY = prm.w_social * G4[1];
          const double cL = prm.w_social * (G4[2] * cT + G4[3] * sT);
          const double cTh = prm.w_social * vj * (-G4[2] * sT + G4[3] * cT);
          cost += 0.5 * r * r;
          mXX += cX * cX; mXY += cX * cY; mXT += cX * cTh; mXL += cX * cL;
          mYY += cY * cY; mYT += cY * cTh; mYL += cY * cL;
          mTT += cTh * cTh; mTL += cTh * cL; mLL += cL * cL;
          qX += cX * r; qY += cY * r; qT += cTh * r; qL += cL * r;
        }
        {
          // Proxemics: w * 3 * exp(-dmin/0.25); with no valid agent the value is 0 but the jet derivative is
          // (-inf)*0 = NaN in Ceres, i.e. the differentiated evaluation fails.
          if (dmin == DBL_MAX) {
            flags |= kJacobianBad;
          } else {
            const double r = prm.w_prox * (3.0 * exp(-dmin / (0.5 * 0.5)));
            const double k = -r * (2.0 / (0.5 * 0.5));
            const double cX = k * pdx, cY = k * pdy;
            cost += 0.5 * r * r;
            mXX += cX * cX; mXY += cX * cY; mYY += cY * cY;
            qX += cX * r; qY += cY * r;
          }
        }
      }
      // --- Velocity (k=3): w (0.6 - v_b)^2 for i < ch ---------------------------------------------------
      if (j < ch) {
        const double e = 0.6 - vj;
        const double r = prm.w_velocity * e * e;
        const double cL = -2.0 * prm.w_velocity * e;
        cost += 0.5 * r * r;
        mLL += cL * cL;
        qL += cL * r;
      }
      // --- GoalAlign (k=4): w * wrap(psi - Theta)^2 -----------------------------------------------------
      {
        const double ga = pb.goal_yaw - Th;
        const double turn = atan2(sin(ga), cos(ga));
        const double r = prm.w_goal * turn * turn;
        const double cTh = -2.0 * prm.w_goal * turn;
        cost += 0.5 * r * r;
        mTT += cTh * cTh;
        qT += cTh * r;
      }
      // --- PathFollow (k=5, final seed point) and PathAlign (k=6, seed point i+1): w (|p - t|^2)^2 -----
      {
        const double ex = X - pb.fin_x, ey = Y - pb.fin_y;
        const double q2 = ex * ex + ey * ey;
        const double r = prm.w_distance * q2 * q2;
        const double k = 4.0 * prm.w_distance * q2;
        const double cX = k * ex, cY = k * ey;
        cost += 0.5 * r * r;
        mXX += cX * cX; mXY += cX * cY; mYY += cY * cY;
        qX += cX * r; qY += cY * r;
      }
      {
        const double ex = X - __ldg(pb.px + j + 1), ey = Y - __ldg(pb.py + j + 1);
        const double q2 = ex * ex + ey * ey;
        const double r = prm.w_angle * q2 * q2;
        const double k = 4.0 * prm.w_angle * q2;
        const double cX = k * ex, cY = k * ey;
        cost += 0.5 * r * r;
        mXX += cX * cX; mXY += cX * cY; mYY += cY * cY;
        qX += cX * r; qY += cY * r;
      }
      // --- Obstacle (k=7): w * bicubic(costmap) 0.25 m ahead ---------------------------------------------
      {
        const double fxw = X + 0.25 * cT, fyw = Y + 0.25 * sT;
        const double gx = (fxw - pb.org_x) / bt.resolution, gy = (fyw - pb.org_y) / bt.resolution;
        double f, dfdr, dfdc;
        bicubic(pb.map, bt.size_x, bt.size_y, gy, gx, f, dfdr, dfdc);
        const double r = prm.w_obstacle * f;
        const double kx = prm.w_obstacle * dfdc / bt.resolution, ky = prm.w_obstacle * dfdr / bt.resolution;
        const double cTh = 0.25 * (-kx * sT + ky * cT);
        cost += 0.5 * r * r;
        mXX += kx * kx; mXY += kx * ky; mXT += kx * cTh;
        mYY += ky * ky; mYT += ky * cTh; mTT += cTh * cTh;
        qX += kx * r; qY += ky * r; qT += cTh * r;
      }

      // --- lane block -> parameter space: rows of D are dX/du, dY/du, dTheta/du, dlv/du --------------
      double D0[P], D1[P], D2[P], D3[P];
      SMPC_UNROLL for (int b = 0; b < NB; ++b) {
        D0[2 * b] = sd[4 * b + 0];
        D1[2 * b] = sd[4 * b + 1];
        D0[2 * b + 1] = sd[4 * b + 2];
        D1[2 * b + 1] = sd[4 * b + 3];
        D2[2 * b] = 0.0;
        D2[2 * b + 1] = dt * (double)steps_in_block_before<NB>(j + 1, b, bl);
        D3[2 * b] = (b == bj) ? 1.0 : 0.0;
        D3[2 * b + 1] = 0.0;
      }
      SMPC_UNROLL for (int a = 0; a < P; ++a) {
        const double t0 = mXX * D0[a] + mXY * D1[a] + mXT * D2[a] + mXL * D3[a];
        const double t1 = mXY * D0[a] + mYY * D1[a] + mYT * D2[a] + mYL * D3[a];
        const double t2 = mXT * D0[a] + mYT * D1[a] + mTT * D2[a] + mTL * D3[a];
        const double t3 = mXL * D0[a] + mYL * D1[a] + mTL * D2[a] + mLL * D3[a];
        g[a] += qX * D0[a] + qY * D1[a] + qT * D2[a] + qL * D3[a];
        SMPC_UNROLL for (int b = 0; b <= a; ++b)
          H[a * (a + 1) / 2 + b] += t0 * D0[b] + t1 * D1[b] + t2 * D2[b] + t3 * D3[b];
      }
    }
  }

  // warp all-reduce
  cost = warp_sum(cost);
  SMPC_UNROLL for (int c = 0; c < P; ++c) g[c] = warp_sum(g[c]);
  SMPC_UNROLL for (int e = 0; e < NH; ++e) H[e] = warp_sum(H[e]);
  flags = __reduce_or_sync(kFullMask, flags);

  // --- VelocityFeasibility (k=8): w ((v_i - v_{i-1})^2 + (w_i - w_{i-1})^2), 0 < i < ch/bl, on blocks i, i-1 ----
  SMPC_UNROLL for (int i = 1; i < NB; ++i) {
    if (i < prm.n_bounded) {
      const double dv = x[2 * i] - x[2 * i - 2], dw = x[2 * i + 1] - x[2 * i - 1];
      const double r = prm.w_vf * dv * dv + prm.w_vf * dw * dw;
      const double jv = 2.0 * prm.w_vf * dv, jw = 2.0 * prm.w_vf * dw;
      cost += 0.5 * r * r;
      // row: [2i-2] = -jv, [2i-1] = -jw, [2i] = jv, [2i+1] = jw
      const int p0 = 2 * i - 2, p1 = 2 * i - 1, p2 = 2 * i, p3 = 2 * i + 1;
      g[p0] -= jv * r; g[p1] -= jw * r; g[p2] += jv * r; g[p3] += jw * r;
      H[p0 * (p0 + 1) / 2 + p0] += jv * jv;
      H[p1 * (p1 + 1) / 2 + p0] += jw * jv;
      H[p1 * (p1 + 1) / 2 + p1] += jw * jw;
      H[p2 * (p2 + 1) / 2 + p0] -= jv * jv;
      H[p2 * (p2 + 1) / 2 + p1] -= jv * jw;
      H[p2 * (p2 + 1) / 2 + p2] += jv * jv;
      H[p3 * (p3 + 1) / 2 + p0] -= jw * jv;
      H[p3 * (p3 + 1) / 2 + p1] -= jw * jw;
      H[p3 * (p3 + 1) / 2 + p2] += jw * jv;
      H[p3 * (p3 + 1) / 2 + p3] += jw * jw;
    }
  }

  out.cost = cost;
  SMPC_UNROLL for (int c = 0; c < P; ++c) out.g[c] = g[c];
  SMPC_UNROLL for (int e = 0; e < NH; ++e) out.H[e] = H[e];
  if (!isfinite(cost)) flags |= kResidualBad;
  bool jfin = true;
  SMPC_UNROLL for (int c = 0; c < P; ++c) jfin = jfin && isfinite(g[c]) && isfinite(H[c * (c + 1) / 2 + c]);
  if (!jfin) flags |= kJacobianBad;
  return flags;
}

// ---------------------------------------------------------------------------------------------------
// Interpolating-polynomial minimiser of the Armijo line search (ceres polynomial.cc).
// Warp-uniform scalar code; called only when a trial step fails the sufficient-decrease test.
// ---------------------------------------------------------------------------------------------------
struct LsSample {
  double x, value, gradient;
  bool value_ok, gradient_ok;
};

__device__ __forceinline__ double poly_eval(const double* c, int n, double x) {
  double v = 0.0;
  for (int i = 0; i < n; ++i) v = v * x + c[i];
  return v;
}

// Real parts of the roots of c[0] x^deg + ... (deg <= 4), Aberth-Ehrlich iteration in complex double
// for deg >= 3 (Ceres: eigenvalues of the companion matrix).
static __device__ __noinline__ int poly_real_roots(const double* cin, int n, double* roots) {
  int lead = 0;
  while (lead + 1 < n && cin[lead] == 0.0) ++lead;
  const double* c = cin + lead;
  const int deg = n - lead - 1;
  if (deg <= 0) return 0;
  if (deg == 1) {
    roots[0] = -c[1] / c[0];
    return 1;
  }
  if (deg == 2) {
    const double a = c[0], b = c[1], cc = c[2];
    const double D = b * b - 4 * a * cc;
    const double sD = sqrt(fabs(D));
    if (D >= 0) {
      if (b >= 0) {
        roots[0] = (-b - sD) / (2.0 * a);
        roots[1] = (2.0 * cc) / (-b - sD);
      } else {
        roots[0] = (2.0 * cc) / (-b + sD);
        roots[1] = (-b + sD) / (2.0 * a);
      }
    } else {
      roots[0] = roots[1] = -b / (2.0 * a);
    }
    return 2;
  }
  double m[5];
  for (int i = 0; i <= deg; ++i) m[i] = c[i] / c[0];
  double radius = 0.0;
  for (int i = 1; i <= deg; ++i) radius = fmax(radius, pow(fabs(m[i]), 1.0 / i));
  radius = 2.0 * radius + 1e-300;
  double zr[4], zi[4];
  for (int i = 0; i < deg; ++i) {
    double s, co;
    sincos(2.0 * M_PI * i / deg + 0.35, &s, &co);
    zr[i] = 0.7 * radius * co;
    zi[i] = 0.7 * radius * s;
  }
  for (int it = 0; it < 200; ++it) {
    double moved = 0.0;
    for (int i = 0; i < deg; ++i) {
      double pr = m[0], pi = 0.0, dr = 0.0, di = 0.0;
      for (int j = 1; j <= deg; ++j) {
        const double ndr = dr * zr[i] - di * zi[i] + pr, ndi = dr * zi[i] + di * zr[i] + pi;
        const double npr = pr * zr[i] - pi * zi[i] + m[j], npi = pr * zi[i] + pi * zr[i];
        dr = ndr; di = ndi; pr = npr; pi = npi;
      }
      if (pr == 0.0 && pi == 0.0) continue;
      const double dn = dr * dr + di * di;
      const double rr = (pr * dr + pi * di) / dn, ri = (pi * dr - pr * di) / dn;  // p / p'
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < deg; ++j) {
        if (j == i) continue;
        const double er = zr[i] - zr[j], ei = zi[i] - zi[j];
        const double en = er * er + ei * ei;
        sr += er / en;
        si -= ei / en;
      }
      const double qr = 1.0 - (rr * sr - ri * si), qi = -(rr * si + ri * sr);  // 1 - ratio*sum
      const double qn = qr * qr + qi * qi;
      const double stepr = (rr * qr + ri * qi) / qn, stepi = (ri * qr - rr * qi) / qn;
      zr[i] -= stepr;
      zi[i] -= stepi;
      moved = fmax(moved, sqrt(stepr * stepr + stepi * stepi) / (sqrt(zr[i] * zr[i] + zi[i] * zi[i]) + 1e-300));
    }
    if (moved < 1e-15) break;
  }
  for (int i = 0; i < deg; ++i) roots[i] = zr[i];
  return deg;
}

// Fit the polynomial through ns (<= 3) samples (values + gradients) with a full-pivot LU, minimise it on [lo, hi].
static __device__ __noinline__ double interpolating_poly_min(const LsSample* s, int ns, double lo, double hi) {
  int nc = 0;
  for (int i = 0; i < ns; ++i) nc += (s[i].value_ok ? 1 : 0) + (s[i].gradient_ok ? 1 : 0);
  const int degree = nc - 1;
  double a[6][6], rhs[6], poly[6];
  int colperm[6];
  for (int i = 0; i < 6; ++i) {
    rhs[i] = 0.0;
    colperm[i] = i;
    for (int j = 0; j < 6; ++j) a[i][j] = 0.0;
  }
  int row = 0;
  for (int i = 0; i < ns; ++i) {
    if (s[i].value_ok) {
      for (int j = 0; j <= degree; ++j) a[row][j] = pow(s[i].x, (double)(degree - j));
      rhs[row++] = s[i].value;
    }
    if (s[i].gradient_ok) {
      for (int j = 0; j < degree; ++j) a[row][j] = (degree - j) * pow(s[i].x, (double)(degree - j - 1));
      rhs[row++] = s[i].gradient;
    }
  }
  for (int k = 0; k < nc; ++k) {
    int pr = k, pc = k;
    double best = -1.0;
    for (int i = k; i < nc; ++i)
      for (int j = k; j < nc; ++j)
        if (fabs(a[i][j]) > best) {
          best = fabs(a[i][j]);
          pr = i;
          pc = j;
        }
    if (best == 0.0) break;
    for (int j = 0; j < nc; ++j) {
      const double t = a[k][j]; a[k][j] = a[pr][j]; a[pr][j] = t;
    }
    { const double t = rhs[k]; rhs[k] = rhs[pr]; rhs[pr] = t; }
    if (pc != k) {
      for (int i = 0; i < nc; ++i) {
        const double t = a[i][k]; a[i][k] = a[i][pc]; a[i][pc] = t;
      }
      const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
    }
    for (int i = k + 1; i < nc; ++i) {
      const double f = a[i][k] / a[k][k];
      if (f == 0.0) continue;
      for (int j = k; j < nc; ++j) a[i][j] -= f * a[k][j];
      rhs[i] -= f * rhs[k];
    }
  }
  double y[6];
  for (int i = nc - 1; i >= 0; --i) {
    double v = rhs[i];
    for (int j = i + 1; j < nc; ++j) v -= a[i][j] * y[j];
    y[i] = (a[i][i] != 0.0) ? v / a[i][i] : 0.0;
  }
  for (int i = 0; i < nc; ++i) poly[colperm[i]] = y[i];

  double ox = (lo + hi) / 2.0;
  double ov = poly_eval(poly, nc, ox);
  const double vlo = poly_eval(poly, nc, lo);
  if (vlo < ov) { ov = vlo; ox = lo; }
  const double vhi = poly_eval(poly, nc, hi);
  if (vhi < ov) { ov = vhi; ox = hi; }
  if (nc > 2) {
    double der[5], roots[4];
    for (int j = 0; j < degree; ++j) der[j] = (degree - j) * poly[j];
    const int nr = poly_real_roots(der, degree, roots);
    for (int i = 0; i < nr; ++i) {
      if (roots[i] < lo || roots[i] > hi) continue;
      const double v = poly_eval(poly, nc, roots[i]);
      if (v < ov) { ov = v; ox = roots[i]; }
    }
  }
  for (int i = 0; i < ns; ++i) {
    if (s[i].x < lo || s[i].x > hi) continue;
    const double v = poly_eval(poly, nc, s[i].x);
    if (v < ov) { ov = v; ox = s[i].x; }
  }
  return ox;
}

// Box projection of ParameterBlock::Plus: only the first n_bounded blocks carry bounds
// (reference src/optimizer.cpp:373-379: v in [0, 0.6], w in [-1.4, 1.4]).
template <int NB>
__device__ __forceinline__ void plus_project(const double (&x)[2 * NB], const double (&d)[2 * NB], double t, int n_bounded,
                                             double (&out)[2 * NB]) {
  SMPC_UNROLL for (int b = 0; b < NB; ++b) {
    double v = x[2 * b] + t * d[2 * b];
    double w = x[2 * b + 1] + t * d[2 * b + 1];
    if (b < n_bounded) {
      v = fmin(fmax(v, 0.0), 0.6);
      w = fmin(fmax(w, -1.4), 1.4);
    }
    out[2 * b] = v;
    out[2 * b + 1] = w;
  }
}

enum Termination {
  kConvGradient = 0,
  kConvParameter = 1,
  kConvFunction = 2,
  kConvRadius = 3,
  kNoConvergence = 4,
  kFailInvalidSteps = 5,
  kFailEvaluation = 6
};

struct SolveOut {
  double cost_initial, cost_final;
  int iterations, termination, n_jac, n_cost;
};

// ---------------------------------------------------------------------------------------------------
// ceres::Solve restated, warp-uniform. x: in = seed block values, out = solution
// (left at the projected seed when the solution is not usable).
// ---------------------------------------------------------------------------------------------------
template <int NB>
__device__ void solve_problem(const DevParams& prm, const DevBatch& bt, const Prob& pb,
                              const double (&aa_target)[kMaxChunks], double (&x)[2 * NB], int lane, SolveOut& so) {
  constexpr int P = 2 * NB;
  constexpr int NH = P * (P + 1) / 2;
  const int nbd = prm.n_bounded;
  so.n_jac = 0;
  so.n_cost = 0;

  double zero[P], x_seed[P];
  SMPC_UNROLL for (int c = 0; c < P; ++c) {
    zero[c] = 0.0;
    x_seed[c] = x[c];
  }
  {
    double xp[P];
    plus_project<NB>(x, zero, 0.0, nbd, xp);  // IterationZero: project the start point
    SMPC_UNROLL for (int c = 0; c < P; ++c) x[c] = xp[c];
  }
  double best[P];
  SMPC_UNROLL for (int c = 0; c < P; ++c) best[c] = x[c];

  Normal<NB> cur;  // normal equations at x
  unsigned fl = evaluate<NB>(prm, bt, pb, aa_target, x, lane, cur);
  ++so.n_jac;
  so.cost_initial = cur.cost;
  so.cost_final = cur.cost;
  so.iterations = 0;
  if (fl) {
    so.termination = kFailEvaluation;
    SMPC_UNROLL for (int c = 0; c < P; ++c) x[c] = x_seed[c];
    return;
  }
  double x_cost = cur.cost;
  double scale[P];
  SMPC_UNROLL for (int c = 0; c < P; ++c) scale[c] = 1.0 / (1.0 + sqrt(cur.H[c * (c + 1) / 2 + c]));

  auto grad_max_norm = [&](const double (&xx)[P], const double (&gg)[P]) {
    double neg[P], proj[P];
    SMPC_UNROLL for (int c = 0; c < P; ++c) neg[c] = -gg[c];
    plus_project<NB>(xx, neg, 1.0, nbd, proj);
    double m = 0.0;
    SMPC_UNROLL for (int c = 0; c < P; ++c) m = fmax(m, fabs(xx[c] - proj[c]));
    return m;
  };
  double gmax = grad_max_norm(x, cur.g);
  double x_norm = 0.0;
  SMPC_UNROLL for (int c = 0; c < P; ++c) x_norm += x[c] * x[c];
  x_norm = sqrt(x_norm);

  double radius = 1e4, decrease_factor = 2.0, minimum_cost = DBL_MAX;
  bool reuse_diagonal = false, it_successful = true, any_success = false;
  int n_invalid = 0, iteration = 0;
  double it_cost = x_cost;
  double diag[P];
  SMPC_UNROLL for (int c = 0; c < P; ++c) diag[c] = 1.0;
  int term = kNoConvergence;
  Normal<NB> trial;

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it_successful && x_cost < minimum_cost) {
      minimum_cost = x_cost;
      SMPC_UNROLL for (int c = 0; c < P; ++c) best[c] = x[c];
    }
    so.cost_final = fmin(so.cost_final, it_cost);
    if (iteration >= prm.max_iterations) { term = kNoConvergence; break; }
    if (it_successful && gmax <= prm.gradient_tol) { term = kConvGradient; break; }
    if (radius <= 1e-32) { term = kConvRadius; break; }
    ++iteration;

    // LevenbergMarquardtStrategy::ComputeStep on the column-scaled normal equations
    if (!reuse_diagonal) {
      SMPC_UNROLL for (int c = 0; c < P; ++c)
        diag[c] = fmin(fmax(scale[c] * scale[c] * cur.H[c * (c + 1) / 2 + c], 1e-6), 1e32);
    }
    reuse_diagonal = true;
    double Lc[NH], rhs[P], step[P];
    SMPC_UNROLL for (int a = 0; a < P; ++a) {
      rhs[a] = scale[a] * cur.g[a];
      SMPC_UNROLL for (int b = 0; b <= a; ++b) Lc[a * (a + 1) / 2 + b] = scale[a] * scale[b] * cur.H[a * (a + 1) / 2 + b];
      const double lm = sqrt(diag[a] / radius);
      Lc[a * (a + 1) / 2 + a] += lm * lm;
    }
    bool step_ok = true;
    SMPC_UNROLL for (int jc = 0; jc < P; ++jc) {  // Cholesky, in place, lower triangle
      double d = Lc[jc * (jc + 1) / 2 + jc];
      SMPC_UNROLL for (int k = 0; k < jc; ++k) d -= Lc[jc * (jc + 1) / 2 + k] * Lc[jc * (jc + 1) / 2 + k];
      if (!(d > 0.0)) step_ok = false;
      d = sqrt(d);
      Lc[jc * (jc + 1) / 2 + jc] = d;
      const double inv_d = 1.0 / d;
      SMPC_UNROLL for (int i = jc + 1; i < P; ++i) {
        double s = Lc[i * (i + 1) / 2 + jc];
        SMPC_UNROLL for (int k = 0; k < jc; ++k) s -= Lc[i * (i + 1) / 2 + k] * Lc[jc * (jc + 1) / 2 + k];
        Lc[i * (i + 1) / 2 + jc] = s * inv_d;
      }
    }
    SMPC_UNROLL for (int i = 0; i < P; ++i) {  // forward substitution
      double s = rhs[i];
      SMPC_UNROLL for (int k = 0; k < i; ++k) s -= Lc[i * (i + 1) / 2 + k] * step[k];
      step[i] = s / Lc[i * (i + 1) / 2 + i];
    }
    SMPC_UNROLL for (int i = P - 1; i >= 0; --i) {  // back substitution
      double s = step[i];
      SMPC_UNROLL for (int k = i + 1; k < P; ++k) s -= Lc[k * (k + 1) / 2 + i] * step[k];
      step[i] = s / Lc[i * (i + 1) / 2 + i];
    }
    SMPC_UNROLL for (int c = 0; c < P; ++c) {
      step_ok = step_ok && isfinite(step[c]);
      step[c] = -step[c];
    }
    // model_cost_change = -(Js s)'(r + Js s / 2) = -s'(Js' r) - s'(Js' Js) s / 2
    double model_cost_change = 0.0;
    if (step_ok) {
      double lin = 0.0, quad = 0.0;
      SMPC_UNROLL for (int a = 0; a < P; ++a) {
        lin += step[a] * scale[a] * cur.g[a];
        double rowv = 0.0;
        SMPC_UNROLL for (int b = 0; b < P; ++b) {
          const int hi = a > b ? a : b, lo = a > b ? b : a;
          rowv += scale[a] * scale[b] * cur.H[hi * (hi + 1) / 2 + lo] * step[b];
        }
        quad += step[a] * rowv;
      }
      model_cost_change = -lin - 0.5 * quad;
    }
    const bool valid = step_ok && (model_cost_change > 0.0);
    if (!valid) {  // HandleInvalidStep
      if (++n_invalid >= 5) {
        --iteration;
        term = kFailInvalidSteps;
        break;
      }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      it_successful = false;
      it_cost = x_cost;
      continue;
    }
    n_invalid = 0;
    double delta[P];
    SMPC_UNROLL for (int c = 0; c < P; ++c) delta[c] = step[c] * scale[c];

    // DoLineSearch: projected Armijo with cubic interpolation. Every trial point gets a full evaluation so
    // that an accepted point needs no second pass (its cost is the candidate cost, its J^T J the next iterate's).
    double g0 = 0.0, dmax = 0.0;
    SMPC_UNROLL for (int c = 0; c < P; ++c) {
      g0 += cur.g[c] * delta[c];
      dmax = fmax(dmax, fabs(delta[c]));
    }
    LsSample smp[3];  // [0] initial, [1] current, [2] previous
    smp[0] = {0.0, x_cost, g0, true, true};
    smp[2] = {0.0, 0.0, 0.0, false, false};
    double cand[P];
    double t = 1.0;
    int ls_iters = 0;
    bool ls_success = false;
    unsigned tfl = 0;
    for (;;) {
      plus_project<NB>(x, delta, t, nbd, cand);
      tfl = evaluate<NB>(prm, bt, pb, aa_target, cand, lane, trial);
      ++so.n_jac;
      smp[1] = {t, trial.cost, 0.0, false, false};
      if (!(tfl & kResidualBad) && !(tfl & kJacobianBad)) {
        smp[1].value_ok = true;
        double gd = 0.0;
        SMPC_UNROLL for (int c = 0; c < P; ++c) gd += delta[c] * trial.g[c];
        smp[1].gradient = gd;
        smp[1].gradient_ok = isfinite(gd);
      }
      if (smp[1].value_ok && !(smp[1].value > x_cost + 1e-4 * g0 * t)) {
        ls_success = true;
        break;
      }
      ++ls_iters;
      if (ls_iters >= 20) break;
      const double lo = 1e-3 * t, hi = 0.6 * t;
      double t_new;
      if (!smp[1].value_ok) {
        t_new = fmin(fmax(t * 0.5, lo), hi);
      } else {
        t_new = interpolating_poly_min(smp, smp[2].value_ok ? 3 : 2, lo, hi);
      }
      if (t_new * dmax < 1e-9) break;
      smp[2] = smp[1];
      t = t_new;
    }
    if (!ls_success && t != 1.0) {
      // line search failed: the TR step is kept as is (delta unchanged) -> candidate = Plus(x, delta)
      plus_project<NB>(x, delta, 1.0, nbd, cand);
      tfl = evaluate<NB>(prm, bt, pb, aa_target, cand, lane, trial);
      ++so.n_jac;
    }
    double cand_cost;
    cand_cost = (tfl & kResidualBad) ? DBL_MAX : trial.cost;

    const bool tol_armed = (prm.ceres_compat < 210) || any_success;
    double step_norm = 0.0;
    SMPC_UNROLL for (int c = 0; c < P; ++c) step_norm += (x[c] - cand[c]) * (x[c] - cand[c]);
    step_norm = sqrt(step_norm);
    if (tol_armed && step_norm <= prm.param_tol * (x_norm + prm.param_tol)) {
      --iteration;
      term = kConvParameter;
      break;
    }
    const double cost_change = x_cost - cand_cost;
    if (tol_armed && fabs(cost_change) <= prm.fn_tol * x_cost) {
      --iteration;
      term = kConvFunction;
      break;
    }
    const double rho = (cand_cost >= DBL_MAX) ? -DBL_MAX : (x_cost - cand_cost) / model_cost_change;
    if (rho > 1e-3) {  // HandleSuccessfulStep
      if (tfl & kJacobianBad) {
        --iteration;
        term = kFailEvaluation;
        break;
      }
      SMPC_UNROLL for (int c = 0; c < P; ++c) x[c] = cand[c];
      x_norm = 0.0;
      SMPC_UNROLL for (int c = 0; c < P; ++c) x_norm += x[c] * x[c];
      x_norm = sqrt(x_norm);
      cur = trial;
      x_cost = cur.cost;
      gmax = grad_max_norm(x, cur.g);
      any_success = true;
      it_successful = true;
      it_cost = x_cost;
      const double q = 2.0 * rho - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
      radius = fmin(1e16, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      it_successful = false;
      it_cost = cand_cost;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
  so.termination = term;
  so.iterations = iteration;
  const bool usable = term <= kNoConvergence;  // Solver::Summary::IsSolutionUsable
  SMPC_UNROLL for (int c = 0; c < P; ++c) x[c] = usable ? best[c] : x_seed[c];
}

// Load the warp-uniform problem view.
__device__ __forceinline__ void load_problem(const DevBatch& bt, int b, Prob& pb) {
  const int S = bt.S;
  pb.x0 = __ldg(bt.pose0 + 3 * (size_t)b);
  pb.y0 = __ldg(bt.pose0 + 3 * (size_t)b + 1);
  pb.yaw0 = __ldg(bt.pose0 + 3 * (size_t)b + 2);
  pb.goal_yaw = __ldg(bt.goal_yaw + b);
  pb.px = bt.path_xy + (size_t)b * 2 * (S + 1);
  pb.py = pb.px + (S + 1);
  pb.fin_x = __ldg(pb.px + S);
  pb.fin_y = __ldg(pb.py + S);
  pb.agents = (bt.A > 0 && bt.agents) ? bt.agents + (size_t)b * bt.A * 6 * (S + 1) : nullptr;
  pb.has_people = (bt.has_people != nullptr) && (bt.has_people[b] != 0);
  const int mi = bt.costmap_index ? __ldg(bt.costmap_index + b) : (b % bt.M);
  pb.map = bt.costmaps + (size_t)mi * bt.size_x * bt.size_y;
  pb.org_x = __ldg(bt.costmap_origin + 2 * mi);
  pb.org_y = __ldg(bt.costmap_origin + 2 * mi + 1);
}

__device__ __forceinline__ void agent_angle_setup(const DevParams& prm, const DevBatch& bt, const Prob& pb, int lane,
                                                  double (&aa_target)[kMaxChunks]) {
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    aa_target[c] = NAN;
    const int j = c * 32 + lane;
    if (pb.has_people && j < bt.S && bt.A > 0 && pb.agents != nullptr) aa_target[c] = agent_angle_target(bt, pb, j + 1);
  }
}

}  // namespace smpc

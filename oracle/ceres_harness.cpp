// TEST INFRASTRUCTURE — Ceres-box harness: the reference's own residual functors under the real ceres::Solve.
//
// CANNOT BE BUILT IN THIS REPOSITORY'S CONTAINER (no Ceres, Eigen or glog). It is kept ready so that, on
// any machine with libceres-dev, TRUE golden vectors can be produced and the "parity unpinned" caveat of DESIGN.md
// retired. It includes the reference critic headers UNMODIFIED (from the reference checkout, -I<ref>/include) and the
// reference critic .cpp files are compiled alongside; only ROS message / tf2 / costmap headers are replaced by the
// minimal stand-ins under oracle/ceres_harness_stubs/.
//
//   make -C oracle ceres_harness REF=/path/to/nav2_social_mpc_controller        (see oracle/Makefile)
//   python tools/dump_problems.py --case crowd_x8_A3 --out /tmp/p.bin            (level-1 arrays of include/smpc.h)
//   oracle/_ref/ceres_harness /tmp/p.bin > /tmp/ceres.jsonl                      (one JSON line per problem)
//   python tools/dump_problems.py --case crowd_x8_A3 --compare /tmp/ceres.jsonl  (oracle vs the real Ceres)
//
// The problem assembly below restates src/optimizer.cpp:241-379 call for call (same Create() factories, same
// AddParameterBlock counts, same AddResidualBlock order, same bounds) on the level-1 inputs, i.e. AFTER
// people_to_status / format_to_optimize / project_people: exactly what libsmpc's smpc_solve_batch consumes. Options as
// src/optimizer.cpp:117-131 (everything else Ceres defaults); max_solver_time_in_seconds is left at its default so that
// the run is deterministic. The Ceres version in use is printed with every result: it decides, among other things,
// what std::numeric_limits<Jet>::max() is (ProxemicsCost).
//
// File format (little endian): int32 B, S, A, size_x, size_y, control_horizon, block_length, max_iterations;
// double resolution, dt, fn_tol, gradient_tol, param_tol, w[9] (distance, social, velocity, angle, agent_angle,
// proxemics, velocity_feasibility, obstacle, goal_align); then per problem: pose0[3], u0[NB][2], path_xy[2][S+1],
// goal_yaw, agents[A][6][S+1], uint8 has_people, origin[2], costmap[size_y*size_x] (u8).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

#include "ceres/ceres.h"
#include "ceres/cubic_interpolation.h"
#include "nav2_social_mpc_controller/critics/agent_angle_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/distance_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/goal_align_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/obstacle_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/proxemics_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/social_work_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/velocity_cost_function.hpp"
#include "nav2_social_mpc_controller/critics/velocity_feasibility_cost_function.hpp"

using namespace nav2_social_mpc_controller;  // NOLINT

struct Header {
  int32_t B, S, A, size_x, size_y, control_horizon, block_length, max_iterations;
  double resolution, dt, fn_tol, gradient_tol, param_tol, w[9];
};
struct Vel {
  double params[2];
};

template <class T>
static bool rd(FILE* f, T* p, size_t n) {
  return std::fread(p, sizeof(T), n, f) == n;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: ceres_harness problems.bin\n");
    return 2;
  }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) {
    std::perror(argv[1]);
    return 2;
  }
  Header h;
  if (!rd(f, &h, 1)) return 2;
  const int S = h.S, S1 = S + 1, A = h.A;
  const unsigned int ch = std::min<unsigned int>(h.control_horizon, S);  // src/optimizer.cpp:248
  const unsigned int bl = std::min<unsigned int>(h.block_length, ch);    // src/optimizer.cpp:249
  const int NB = (ch + bl - 1) / bl;
  const double w_distance = h.w[0], w_social = h.w[1], w_velocity = h.w[2], w_angle = h.w[3], w_agent_angle = h.w[4],
               w_prox = h.w[5], w_vf = h.w[6], w_obstacle = h.w[7], w_goal = h.w[8];
  for (int b = 0; b < h.B; ++b) {
    double pose0[3], goal_yaw, origin[2];
    std::vector<double> u0(2 * NB), path(2 * S1), agents(static_cast<size_t>(A) * 6 * S1);
    uint8_t has_people;
    std::vector<unsigned char> cmap(static_cast<size_t>(h.size_x) * h.size_y);
    if (!rd(f, pose0, 3) || !rd(f, u0.data(), u0.size()) || !rd(f, path.data(), path.size()) || !rd(f, &goal_yaw, 1) ||
        (A > 0 && !rd(f, agents.data(), agents.size())) || !rd(f, &has_people, 1) || !rd(f, origin, 2) ||
        !rd(f, cmap.data(), cmap.size())) {
      std::fprintf(stderr, "short file\n");
      return 2;
    }

    nav2_costmap_2d::Costmap2D costmap(cmap.data(), h.size_x, h.size_y, h.resolution, origin[0], origin[1]);
    auto grid = std::make_shared<ceres::Grid2D<u_char>>(costmap.getCharMap(), 0, costmap.getSizeInCellsY(), 0,
                                                        costmap.getSizeInCellsX());             // :167-168
    auto interp = std::make_shared<ceres::BiCubicInterpolator<ceres::Grid2D<u_char>>>(*grid);  // :170
    geometry_msgs::msg::Pose robot0;
    robot0.position.x = pose0[0];
    robot0.position.y = pose0[1];
    tf2::Quaternion q;
    q.setRPY(0, 0, pose0[2]);
    robot0.orientation = tf2::toMsg(q);  // :224-226
    // optim_velocities[i]: parameter block b starts at the seed velocity of TIME index b; only the first
    // NB entries are ever handed to Ceres
    std::vector<Vel> vel(std::max(S, NB));
    for (int i = 0; i < NB; ++i) {
      vel[i].params[0] = u0[2 * i];
      vel[i].params[1] = u0[2 * i + 1];
    }
    auto people_at = [&](int step) {  // people_proj[step]: A columns (the reference: 3) of (x, y, yaw, t, lv, av)
      AgentsStates st;
      for (int k = 0; k < A; ++k) {
        AgentStatus a;
        for (int c = 0; c < 6; ++c) a[c] = agents[(static_cast<size_t>(k) * 6 + c) * S1 + step];
        st.push_back(a);
      }
      return st;
    };
    Eigen::Matrix<double, 2, 1> final_point(path[S], path[S1 + S]);  // :234-235
    Eigen::Matrix<double, 2, 1> final_heading(S * h.dt, goal_yaw);   // :298 (the functor reads [1] = yaw)

    ceres::Problem problem;
    std::vector<double*> blocks;
    double counter = 0.0;
    for (unsigned int i = 0; i < static_cast<unsigned int>(S); ++i) {  // :251-371
      counter += 1.0;
      const unsigned int block_used = i / bl;
      if (i < ch && (blocks.empty() || blocks.back() != vel[block_used].params)) blocks.push_back(vel[block_used].params);
      const double counter_step = counter * h.dt;
      const unsigned int n_seen = (i < ch) ? i / bl + 1 : (ch - 1) / bl + 1;
      if (has_people) {
        auto* social = SocialWorkCost::Create(w_social, people_at(i + 1), robot0, counter_step, i, h.dt, ch, bl);
        auto* angle = AgentAngleCost::Create(w_agent_angle, people_at(i + 1), robot0, i, h.dt, ch, bl);
        auto* prox = ProxemicsCost::Create(w_prox, people_at(i + 1), robot0, counter_step, i, h.dt, ch, bl);
        for (unsigned int j = 0; j < n_seen; ++j) {
          angle->AddParameterBlock(2);
          social->AddParameterBlock(2);
          prox->AddParameterBlock(2);
        }
        angle->SetNumResiduals(1);
        social->SetNumResiduals(1);
        prox->SetNumResiduals(1);
        problem.AddResidualBlock(angle, NULL, blocks);
        problem.AddResidualBlock(social, NULL, blocks);
        problem.AddResidualBlock(prox, NULL, blocks);
      }
      auto* velocity = VelocityCost::Create(w_velocity, 0.6, i, ch, bl);  // :238, :296-297
      auto* goal = GoalAlignCost::Create(w_goal, final_heading, robot0, i, h.dt, ch, bl);
      for (unsigned int j = 0; j < n_seen; ++j) {
        velocity->AddParameterBlock(2);
        goal->AddParameterBlock(2);
      }
      velocity->SetNumResiduals(1);
      goal->SetNumResiduals(1);
      problem.AddResidualBlock(velocity, NULL, blocks);
      problem.AddResidualBlock(goal, NULL, blocks);
      Eigen::Matrix<double, 2, 1> point(path[i + 1], path[S1 + i + 1]);  // :327
      auto* follow = DistanceCost::Create(w_distance, final_point, robot0, i, h.dt, ch, bl);
      auto* align = DistanceCost::Create(w_angle, point, robot0, i, h.dt, ch, bl);
      auto* obst = ObstacleCost::Create(w_obstacle, &costmap, interp, robot0, i, h.dt, ch, bl);
      for (unsigned int j = 0; j < n_seen; ++j) {
        follow->AddParameterBlock(2);
        align->AddParameterBlock(2);
        obst->AddParameterBlock(2);
      }
      follow->SetNumResiduals(1);
      align->SetNumResiduals(1);
      obst->SetNumResiduals(1);
      problem.AddResidualBlock(follow, NULL, blocks);
      problem.AddResidualBlock(align, NULL, blocks);
      problem.AddResidualBlock(obst, NULL, blocks);
      if (i != 0 && i < ch / bl) {
        auto* vf = VelocityFeasibilityCost::Create(w_vf, i, ch);
        problem.AddResidualBlock(vf, NULL, vel[i].params, vel[i - 1].params);  // :364-370
      }
    }
    for (unsigned int i = 0; i < ch / bl; ++i) {  // :373-379
      problem.SetParameterLowerBound(vel[i].params, 0, 0.0);
      problem.SetParameterUpperBound(vel[i].params, 0, 0.6);
      problem.SetParameterLowerBound(vel[i].params, 1, -1.4);
      problem.SetParameterUpperBound(vel[i].params, 1, 1.4);
    }
    ceres::Solver::Options options;  // :117-131
    options.linear_solver_type = ceres::DENSE_SCHUR;
    options.max_num_iterations = h.max_iterations;
    options.function_tolerance = h.fn_tol;
    options.gradient_tolerance = h.gradient_tol;
    options.parameter_tolerance = h.param_tol;
    options.logging_type = ceres::SILENT;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);

    std::printf("{\"problem\": %d, \"ceres_version\": \"%s\", \"termination_type\": %d, \"usable\": %d, "
                "\"initial_cost\": %.17g, \"final_cost\": %.17g, \"iterations\": %d, \"num_successful_steps\": %d, "
                "\"num_unsuccessful_steps\": %d, \"num_line_search_steps\": %d, \"u\": [",
                b, CERES_VERSION_STRING, static_cast<int>(summary.termination_type), summary.IsSolutionUsable() ? 1 : 0,
                summary.initial_cost, summary.final_cost,
                summary.iterations.empty() ? 0 : summary.iterations.back().iteration, summary.num_successful_steps,
                summary.num_unsuccessful_steps, summary.num_line_search_steps);
    for (int i = 0; i < NB; ++i) std::printf("%s%.17g, %.17g", i ? ", " : "", vel[i].params[0], vel[i].params[1]);
    std::printf("], \"trace\": [");
    for (size_t k = 0; k < summary.iterations.size(); ++k) {
      const ceres::IterationSummary& it = summary.iterations[k];
      std::printf("%s[%d, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %d, %d]", k ? ", " : "", it.iteration, it.cost,
                  it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius,
                  it.step_is_valid ? 1 : 0, it.step_is_successful ? 1 : 0);
    }
    std::printf("]}\n");
  }
  std::fclose(f);
  return 0;
}

// smpc_optimize.cu — level-2 entry of libsmpc.so: bool Optimizer::optimize(...) for one robot
// (reference include/nav2_social_mpc_controller/optimizer.hpp:167-170, src/optimizer.cpp:148-452).
// The pre-solve stages are the reference's host-side stages restated without ROS / Eigen types; the solve and the
// post-solve expansion are the GPU level-1 path (smpc_solve_batch). Own structure and naming; arithmetic follows
// the cited lines (float time_step / max_time / blend weights included).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smpc.h"
#include "smpc_host_state.h"

namespace {

struct V2 {
  double x, y;
};
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator*(double s, V2 a) { return {s * a.x, s * a.y}; }
inline V2 operator*(V2 a, double s) { return {a.x * s, a.y * s}; }
inline V2 operator/(V2 a, double s) { return {a.x / s, a.y / s}; }
inline double norm(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }
inline V2 normalized(V2 a) {  // Eigen normalized()
  const double z = a.x * a.x + a.y * a.y;
  if (z > 0.0) return a / std::sqrt(z);
  return a;
}
inline double wrap_pi(double a) {
  while (a <= -M_PI) a += 2 * M_PI;
  while (a > M_PI) a -= 2 * M_PI;
  return a;
}

// tf2 Quaternion::setRPY(0, 0, yaw) -> tf2::getYaw
inline double yaw_roundtrip(double yaw) {
  const double h = yaw * 0.5;
  const double qz = std::sin(h), qw = std::cos(h);
  return std::atan2(2 * (qw * qz), qw * qw - qz * qz);
}

struct Status {  // AgentStatus: x, y, yaw, t, lv, av (tools/type_definitions.hpp:6)
  double v[6];
};

// lightsfm agent as used by project_people (sfm.hpp:90-140); one obstacle, one goal, no groups
struct SfmAgent {
  V2 position, velocity;
  double yaw, desired_velocity, radius, linear_velocity, angular_velocity;
  bool has_goal;
  V2 goal;
  double goal_radius;
  bool has_obstacle;
  V2 obstacle;  // what the reference stores in obstacles1 (actually agent - obstacle)
  V2 global_force;
};

// Optimizer::computeObstacle, src/optimizer.cpp:673-728. Returns false on the std::runtime_error cases.
bool compute_obstacle(const smpc_obstacle_distance& od, V2 apos, V2* out, std::string* err) {
  if (od.distances == nullptr || od.indexes == nullptr) {
    *err = "ObstacleDistance grid is empty";
    return false;
  }
  if (od.width == 0 || od.height == 0) {
    *err = "ObstacleDistance grid has invalid size";
    return false;
  }
  if (!(od.resolution > 0.0f)) {
    *err = "ObstacleDistance grid has invalid resolution";
    return false;
  }
  const unsigned int xcell = (unsigned int)std::floor((apos.x - od.origin_x) / od.resolution);
  const unsigned int ycell = (unsigned int)std::floor((apos.y - od.origin_y) / od.resolution);
  if (xcell >= od.width || ycell >= od.height) {
    *err = "ObstacleDistance grid cell out of bounds";
    return false;
  }
  const unsigned int index = xcell + ycell * od.width;
  const unsigned int ob_idx = od.indexes[index];
  if (ob_idx >= od.width * od.height) {
    *err = "ObstacleDistance grid index out of bounds";
    return false;
  }
  const unsigned int oy = ob_idx / od.width, ox = ob_idx % od.width;
  const float x = ox * od.resolution + od.origin_x;  // float-rounded like the reference (:719-720)
  const float y = oy * od.resolution + od.origin_y;
  *out = apos - V2{(double)x, (double)y};
  return true;
}

// SocialForceModel::computeForces for ungrouped agents (sfm.hpp:188-323, 462-487)
void sfm_compute_forces(std::vector<SfmAgent>& ag) {
  const double kDesired = 2.0, kObstacle = 20.0, kSigma = 0.2, kSocial = 2.1, kLambda = 2.0, kGamma = 0.35, kN = 2.0,
               kNPrime = 3.0, kRelax = 0.5;
  for (size_t i = 0; i < ag.size(); ++i) {
    SfmAgent& a = ag[i];
    V2 desired;
    if (a.has_goal && norm(a.goal - a.position) > a.goal_radius) {
      const V2 dir = normalized(a.goal - a.position);
      desired = kDesired * (dir * a.desired_velocity - a.velocity) / kRelax;
    } else {
      desired = (-1.0 * a.velocity) / kRelax;
    }
    V2 obstacle{0.0, 0.0};
    if (a.has_obstacle) {
      const V2 min_diff = a.position - a.obstacle;
      const double distance = norm(min_diff) - a.radius;
      obstacle = obstacle + kObstacle * std::exp(-distance / kSigma) * normalized(min_diff);
      obstacle = obstacle / 1.0;
    }
    V2 social{0.0, 0.0};
    for (size_t k = 0; k < ag.size(); ++k) {
      if (k == i) continue;
      const V2 diff = ag[k].position - a.position;
      const V2 dir = normalized(diff);
      const V2 vel_diff = a.velocity - ag[k].velocity;
      const V2 iv = kLambda * vel_diff + dir;
      const double ilen = norm(iv);
      const V2 idir = iv / ilen;
      const double a1 = wrap_pi(std::atan2(idir.y, idir.x));
      const double a2 = wrap_pi(std::atan2(dir.y, dir.x));
      const double theta = wrap_pi(a2 - a1);
      const double B = kGamma * ilen;
      const double fv = -std::exp(-norm(diff) / B - (kNPrime * B * theta) * (kNPrime * B * theta));
      double sign = -1.0;
      if (theta == 0) sign = 0; else if (theta > 0) sign = 1;
      const double fa = -sign * std::exp(-norm(diff) / B - (kN * B * theta) * (kN * B * theta));
      const V2 f_vel = fv * idir;
      const V2 f_ang = fa * V2{-idir.y, idir.x};
      social = social + kSocial * (f_vel + f_ang);
    }
    a.global_force = desired + social + obstacle + V2{0.0, 0.0};
  }
}

// SocialForceModel::updatePosition (sfm.hpp:512-560)
void sfm_update(std::vector<SfmAgent>& ag, double dt) {
  for (SfmAgent& a : ag) {
    a.velocity = a.velocity + a.global_force * dt;
    if (norm(a.velocity) > a.desired_velocity) {
      a.velocity = normalized(a.velocity);
      a.velocity = a.velocity * a.desired_velocity;
    }
    const double init_yaw = a.yaw;
    a.yaw = wrap_pi(std::atan2(a.velocity.y, a.velocity.x));
    a.angular_velocity = wrap_pi(a.yaw - init_yaw) / dt;
    a.position = a.position + a.velocity * dt;
    a.linear_velocity = norm(a.velocity);
    if (a.has_goal && norm(a.goal - a.position) <= a.goal_radius) a.has_goal = false;  // goals.pop_front()
  }
}

}  // namespace

extern "C" {

int smpc_reset_memory(smpc_handle* h) {
  if (!h) return smpc_host_fail(SMPC_ERR_ARGUMENT, "handle is NULL");
  smpc_memory* m = smpc_handle_memory(h);
  m->prev_poses.clear();
  m->prev_cmds.clear();
  return SMPC_OK;
}

int smpc_optimize(smpc_handle* h, smpc_optimize_io* io) {
  if (!h || !io) return smpc_host_fail(SMPC_ERR_ARGUMENT, "NULL argument");
  if (!io->poses || !io->cmds || !io->costmap) return smpc_host_fail(SMPC_ERR_ARGUMENT, "poses / cmds / costmap missing");
  const smpc_params* prm = smpc_handle_params(h);
  smpc_memory* mem = smpc_handle_memory(h);
  io->optimized = 0;
  io->n_proj_steps = 0;

  // people_to_status, src/optimizer.cpp:454-482
  std::vector<Status> init_people;
  for (int k = 0; k < io->n_people; ++k) {
    const double* p = io->people + 5 * k;
    Status s{{p[0], p[1], std::atan2(p[3], p[2]), 0.0, std::sqrt(p[2] * p[2] + p[3] * p[3]), p[4]}};
    init_people.push_back(s);
  }
  while (init_people.size() < 3) init_people.push_back(Status{{0.0, 0.0, 0.0, -1.0, 0.0, 0.0}});
  while (init_people.size() > 3) init_people.pop_back();

  if (io->n_poses < 2) return SMPC_OK;  // "Path has less than 2 points" -> return false (:158-162)

  // TrajectoryMemory (:174-186): first call seeds the memory with the current path / cmds
  std::vector<double> cur_poses(io->poses, io->poses + 3 * (size_t)io->n_poses);
  std::vector<double> cur_cmds(io->cmds, io->cmds + 2 * (size_t)io->n_cmds);
  if (mem->prev_poses.empty()) {
    mem->prev_poses = cur_poses;
    mem->prev_cmds = cur_cmds;
  }
  const std::vector<double> prev_poses = mem->prev_poses;
  const std::vector<double> prev_cmds = mem->prev_cmds;

  // format_to_optimize, :484-551
  const float timestep = io->time_step, maxtime = prm->max_time;
  const float wpath = prm->current_path_w, wcmd = prm->current_cmds_w;
  int n_poses = io->n_poses;
  const int maxsize = (int)std::round(maxtime / timestep);
  if (n_poses > maxsize) n_poses = maxsize - 1;
  if (n_poses < 2) return smpc_host_fail(SMPC_ERR_ARGUMENT, "max_time / time_step leaves fewer than 2 poses");
  if (io->n_cmds < n_poses - 1) return smpc_host_fail(SMPC_ERR_ARGUMENT, "fewer cmds than poses - 1");
  if (io->capacity < n_poses) return smpc_host_fail(SMPC_ERR_ARGUMENT, "capacity smaller than the number of poses");
  const size_t n_prev_poses = prev_poses.size() / 3, n_prev_cmds = prev_cmds.size() / 2;
  std::vector<Status> robot(n_poses);
  for (int i = 0; i < n_poses; ++i) {
    double x = cur_poses[3 * i], y = cur_poses[3 * i + 1], yaw = cur_poses[3 * i + 2];
    if ((size_t)i < n_prev_poses) {
      x = wpath * cur_poses[3 * i] + (1.0 - wpath) * prev_poses[3 * i];
      y = wpath * cur_poses[3 * i + 1] + (1.0 - wpath) * prev_poses[3 * i + 1];
      const double smoothed = wpath * cur_poses[3 * i + 2] + (1.0 - wpath) * prev_poses[3 * i + 2];
      yaw = yaw_roundtrip(smoothed);  // setRPY -> toMsg -> getYaw (:517-525)
      cur_poses[3 * i] = x;
      cur_poses[3 * i + 1] = y;
      cur_poses[3 * i + 2] = yaw;
    }
    Status& r = robot[i];
    r.v[0] = x;
    r.v[1] = y;
    r.v[2] = yaw;
    r.v[3] = (double)((float)i * timestep);  // unsigned * float (:526)
    if (i == 0) {
      r.v[4] = io->speed_v;
      r.v[5] = io->speed_w;
    } else {
      // previous_cmds[i-1] is read unguarded by the reference; a missing entry falls back to the current cmd
      const double pv = ((size_t)(i - 1) < n_prev_cmds) ? prev_cmds[2 * (i - 1)] : cur_cmds[2 * (i - 1)];
      const double pw = ((size_t)(i - 1) < n_prev_cmds) ? prev_cmds[2 * (i - 1) + 1] : cur_cmds[2 * (i - 1) + 1];
      r.v[4] = wcmd * cur_cmds[2 * (i - 1)] + (1.0 - wcmd) * pv;
      r.v[5] = wcmd * cur_cmds[2 * (i - 1) + 1] + (1.0 - wcmd) * pw;
    }
  }

  // project_people, :554-671
  std::vector<std::vector<Status>> proj;
  proj.push_back(init_people);
  {
    std::vector<SfmAgent> agents;
    std::string err;
    for (size_t k = 0; k < init_people.size(); ++k) {
      const Status& s = init_people[k];
      if (s.v[3] == -1) continue;
      SfmAgent a{};
      a.position = {s.v[0], s.v[1]};
      a.yaw = s.v[2];
      a.linear_velocity = s.v[4];
      a.angular_velocity = s.v[5];
      a.velocity = {a.linear_velocity * std::cos(a.yaw), a.linear_velocity * std::sin(a.yaw)};
      a.desired_velocity = 0.5;
      a.radius = 0.5;
      a.has_goal = true;
      a.goal_radius = 0.25;
      a.goal = a.position + (double)maxtime * a.velocity;
      if (io->od.width == 100 && io->od.height == 100) continue;  // "grid is NOT valid" (:598-603)
      V2 ob;
      if (!compute_obstacle(io->od, a.position, &ob, &err)) return smpc_host_fail(SMPC_ERR_ARGUMENT, err);
      a.has_obstacle = true;
      a.obstacle = ob;
      agents.push_back(a);
    }
    for (int i = 0; i + 1 < n_poses; ++i) {
      SfmAgent rb{};
      rb.desired_velocity = 0.6;
      rb.radius = 0.5;
      rb.position = {robot[i].v[0], robot[i].v[1]};
      rb.yaw = robot[i].v[2];
      rb.linear_velocity = robot[i].v[4];
      rb.angular_velocity = robot[i].v[5];
      rb.velocity = {rb.linear_velocity * std::cos(rb.yaw), rb.linear_velocity * std::sin(rb.yaw)};
      rb.has_goal = true;
      rb.goal_radius = 0.25;
      rb.goal = {robot.back().v[0], robot.back().v[1]};
      rb.has_obstacle = false;
      agents.push_back(rb);
      sfm_compute_forces(agents);
      sfm_update(agents, (double)timestep);
      agents.pop_back();
      for (SfmAgent& a : agents) {
        V2 ob;
        if (!compute_obstacle(io->od, a.position, &ob, &err)) return smpc_host_fail(SMPC_ERR_ARGUMENT, err);
        a.obstacle = ob;
      }
      std::vector<Status> humans;
      for (const SfmAgent& a : agents)
        humans.push_back(Status{{a.position.x, a.position.y, a.yaw, (double)((float)(i + 1) * timestep),
                                 a.linear_velocity, a.angular_velocity}});
      while (humans.size() < init_people.size()) humans.push_back(Status{{0.0, 0.0, 0.0, -1.0, 0.0, 0.0}});
      proj.push_back(humans);
    }
  }
  io->n_proj_steps = n_poses;
  if (io->people_proj)
    for (int i = 0; i < n_poses; ++i)
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 6; ++c) io->people_proj[((size_t)i * 3 + k) * 6 + c] = proj[i][k].v[c];

  // unpack (:197-237) -> level-1 batch of one problem
  const int S = n_poses - 1;
  int nb = 0;
  int rc = smpc_problem_dims(prm, S, nullptr, nullptr, &nb, nullptr);
  if (rc != SMPC_OK) return rc;
  std::vector<double> pose0{robot[0].v[0], robot[0].v[1], yaw_roundtrip(robot[0].v[2])};  // setRPY (:224-226) + getYaw
  std::vector<double> u0(2 * (size_t)nb), path_xy(2 * (size_t)(S + 1)), agents(3 * 6 * (size_t)(S + 1));
  for (int b = 0; b < nb; ++b) {
    u0[2 * b] = robot[b].v[4];
    u0[2 * b + 1] = robot[b].v[5];
  }
  for (int i = 0; i <= S; ++i) {
    path_xy[i] = robot[i].v[0];
    path_xy[S + 1 + i] = robot[i].v[1];
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < 6; ++c) agents[((size_t)k * 6 + c) * (S + 1) + i] = proj[i][k].v[c];
  }
  double goal_yaw = robot.back().v[2];
  uint8_t has_people = io->n_people != 0;
  double origin[2] = {io->origin_x, io->origin_y};
  smpc_batch in{};
  std::memset(&in, 0, sizeof(in));
  in.n_problems = 1;
  in.n_steps = S;
  in.n_agents = 3;
  in.n_costmaps = 1;
  in.size_x = io->size_x;
  in.size_y = io->size_y;
  in.resolution = io->resolution;
  in.dt = (double)timestep;
  in.pose0 = pose0.data();
  in.u0 = u0.data();
  in.path_xy = path_xy.data();
  in.goal_yaw = &goal_yaw;
  in.agents = agents.data();
  in.has_people = &has_people;
  in.costmaps = io->costmap;
  in.costmap_origin = origin;
  std::vector<double> cmds_out(2 * (size_t)(S + 1)), path_out(3 * (size_t)(S + 1));
  uint8_t usable = 0;
  int32_t termination = 0, iterations = 0;
  double c0 = 0.0, c1 = 0.0;
  smpc_result out{};
  std::memset(&out, 0, sizeof(out));
  out.cmds = cmds_out.data();
  out.path = path_out.data();
  out.usable = &usable;
  out.termination = &termination;
  out.iterations = &iterations;
  out.cost_initial = &c0;
  out.cost_final = &c1;
  rc = smpc_solve_batch(h, &in, &out);
  if (rc != SMPC_OK) return rc;
  io->termination = termination;
  io->iterations = iterations;
  io->cost_initial = c0;
  io->cost_final = c1;
  if (!usable) {  // "Optimization failed!!!" -> return false (:384-388); the path keeps the cut / blended seed
    io->n_poses = n_poses;
    std::memcpy(io->poses, cur_poses.data(), sizeof(double) * 3 * n_poses);
    return SMPC_OK;
  }
  // :390-449
  io->n_poses = S + 1;
  io->n_cmds = S + 1;
  std::memcpy(io->poses, path_out.data(), sizeof(double) * 3 * (S + 1));
  std::memcpy(io->cmds, cmds_out.data(), sizeof(double) * 2 * (S + 1));
  mem->prev_poses = path_out;
  mem->prev_cmds = cmds_out;
  io->optimized = 1;
  return SMPC_OK;
}

}  // extern "C"

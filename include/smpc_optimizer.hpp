// smpc_optimizer.hpp — C++ host side above the C-ABI (include/smpc.h): the reference's optimizer interface for the
// solve path, same names, argument meaning and error behaviour, with the Ceres solve replaced by libsmpc.so.
//
// Mirrors (file:line under the reference tree):
//   struct OptimizerParams                      include/nav2_social_mpc_controller/optimizer.hpp:59-101
//   OptimizerParams::get                        src/optimizer.cpp:16-85          -> OptimizerParams::get(yaml, plugin)
//   Optimizer::initialize(const OptimizerParams)   optimizer.hpp:152, src/optimizer.cpp:98-132
//   bool Optimizer::optimize(path, people_proj, costmap, obstacles, cmds, people, speed, time_step)
//                                               optimizer.hpp:167-170, src/optimizer.cpp:148-452
//   AgentStatus / AgentsStates / AgentTrajectory / AgentsTrajectories   tools/type_definitions.hpp:6-9
//
// The class is a template over a message family `M` so that ONE source serves both builds:
//   * inside a ROS 2 workspace  : M = RosMsgs  (ros_shim/ros_msgs.hpp: nav_msgs, geometry_msgs, people_msgs,
//                                 obstacle_distance_msgs, nav2_costmap_2d::Costmap2D)
//   * in this repo (no ROS here): M = PlainMsgs (include/smpc_plain_msgs.hpp: PODs with the same member names)
// Only the members the reference touches are used: path.header, path.poses[i].pose.position.{x,y},
// .pose.orientation.{x,y,z,w}, cmds[i].header, cmds[i].twist.linear.{x,y}, .twist.angular.z,
// people.people[k].position.{x,y}, .velocity.{x,y,z}, speed.linear.x, speed.angular.z, obstacles.info.{width,height,
// resolution,origin.position.{x,y}}, obstacles.distances, obstacles.indexes, costmap->getCharMap(),
// getSizeInCellsX/Y(), getOriginX/Y(), getResolution().
//
// Error behaviour: initialize() throws std::runtime_error for an invalid linear_solver_type (src/optimizer.cpp:44) or
// when no sm_100 GPU / library is available (no CPU fallback exists); optimize() returns false exactly where the
// reference does (solver summary not usable, src/optimizer.cpp:384-388) and throws std::runtime_error where
// computeObstacle does (:676-713). Thread-safety: one Optimizer per caller thread, like the reference.
#pragma once

#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "smpc.h"

namespace nav2_social_mpc_controller_b200
{

// tools/type_definitions.hpp:6-9 — x, y, yaw, t, lv, av
using AgentStatus = std::array<double, 6>;
using AgentsStates = std::vector<AgentStatus>;
using AgentTrajectory = std::vector<AgentStatus>;
using AgentsTrajectories = std::vector<AgentsStates>;

struct OptimizerParams
{
  // same member names as the reference struct (optimizer.hpp:79-100)
  std::string linear_solver_type;
  double param_tol;
  double fn_tol;
  double gradient_tol;
  double socialwork_w_;
  double distance_w_;
  double velocity_w_;
  double angle_w_;
  double agent_angle_w_;
  double velocity_feasibility_w_;
  double goal_align_w_;
  double obstacle_w_;
  double proxemics_w_;
  float current_path_w;
  float current_cmds_w;
  float max_time;
  int discretization_;
  int control_horizon_;
  int parameter_block_length_;
  bool debug;
  int max_iterations;
  // not reference yaml: which Ceres release's trust-region loop to follow; 200 = libceres-dev of Humble
  int ceres_compat;

  OptimizerParams()
  {
    smpc_params p;
    smpc_params_default(&p);  // the declared defaults of OptimizerParams::get, src/optimizer.cpp:26-84
    from_c(p);
  }

  // OptimizerParams::get(node, name): here the parameters come from the same yaml the node would have loaded.
  void get(const std::string& yaml_path, const std::string& name)
  {
    smpc_params p;
    if (smpc_params_from_yaml(yaml_path.c_str(), name.c_str(), &p) != SMPC_OK)
      throw std::runtime_error(smpc_last_error());  // incl. "Invalid linear_solver_type", src/optimizer.cpp:44
    from_c(p);
  }

  void from_c(const smpc_params& p)
  {
    linear_solver_type = p.linear_solver_type;
    param_tol = p.param_tol;
    fn_tol = p.fn_tol;
    gradient_tol = p.gradient_tol;
    socialwork_w_ = p.socialwork_w;
    distance_w_ = p.distance_w;
    velocity_w_ = p.velocity_w;
    angle_w_ = p.angle_w;
    agent_angle_w_ = p.agent_angle_w;
    velocity_feasibility_w_ = p.velocity_feasibility_w;
    goal_align_w_ = p.goal_align_w;
    obstacle_w_ = p.obstacle_w;
    proxemics_w_ = p.proxemics_w;
    current_path_w = p.current_path_w;
    current_cmds_w = p.current_cmds_w;
    max_time = p.max_time;
    discretization_ = p.discretization;
    control_horizon_ = p.control_horizon;
    parameter_block_length_ = p.parameter_block_length;
    debug = p.debug != 0;
    max_iterations = p.max_iterations;
    ceres_compat = p.ceres_compat;
  }

  smpc_params to_c() const
  {
    smpc_params p;
    smpc_params_default(&p);
    std::snprintf(p.linear_solver_type, sizeof p.linear_solver_type, "%s", linear_solver_type.c_str());
    p.param_tol = param_tol;
    p.fn_tol = fn_tol;
    p.gradient_tol = gradient_tol;
    p.socialwork_w = socialwork_w_;
    p.distance_w = distance_w_;
    p.velocity_w = velocity_w_;
    p.angle_w = angle_w_;
    p.agent_angle_w = agent_angle_w_;
    p.velocity_feasibility_w = velocity_feasibility_w_;
    p.goal_align_w = goal_align_w_;
    p.obstacle_w = obstacle_w_;
    p.proxemics_w = proxemics_w_;
    p.current_path_w = current_path_w;
    p.current_cmds_w = current_cmds_w;
    p.max_time = max_time;
    p.discretization = discretization_;
    p.control_horizon = control_horizon_;
    p.parameter_block_length = parameter_block_length_;
    p.debug = debug ? 1 : 0;
    p.max_iterations = max_iterations;
    p.ceres_compat = ceres_compat;
    return p;
  }
};

namespace detail
{
// tf2::getYaw(const geometry_msgs::msg::Quaternion&) (tf2/utils.h; used at update_state.hpp:44, src/optimizer.cpp:224)
template <class Q>
inline double get_yaw(const Q& q)
{
  const double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
  const double sarg = -2.0 * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
  if (sarg <= -0.99999) return -2.0 * std::atan2(q.y, q.x);
  if (sarg >= 0.99999) return 2.0 * std::atan2(q.y, q.x);
  return std::atan2(2.0 * (q.x * q.y + q.w * q.z), sqw + sqx - sqy - sqz);
}
// tf2::Quaternion::setRPY(0, 0, yaw) + tf2::toMsg (src/optimizer.cpp:437-439)
template <class Q>
inline void set_yaw(Q& q, double yaw)
{
  q.x = 0.0;
  q.y = 0.0;
  q.z = std::sin(0.5 * yaw);
  q.w = std::cos(0.5 * yaw);
}
}  // namespace detail

template <class M>
class OptimizerT
{
public:
  using Path = typename M::Path;
  using PoseStamped = typename M::PoseStamped;
  using TwistStamped = typename M::TwistStamped;
  using Twist = typename M::Twist;
  using People = typename M::People;
  using ObstacleDistance = typename M::ObstacleDistance;
  using Costmap2D = typename M::Costmap2D;

  OptimizerT() = default;
  OptimizerT(const OptimizerT&) = delete;
  OptimizerT& operator=(const OptimizerT&) = delete;
  ~OptimizerT()
  {
    if (handle_) smpc_destroy(handle_);
  }

  // Optimizer::initialize (src/optimizer.cpp:98-132). `device` = CUDA ordinal of the B200 this controller uses.
  void initialize(const OptimizerParams params, int device = 0)
  {
    if (handle_) {
      smpc_destroy(handle_);
      handle_ = nullptr;
    }
    params_ = params;
    const smpc_params p = params.to_c();
    if (smpc_create(&p, device, &handle_) != SMPC_OK) throw std::runtime_error(smpc_last_error());
  }

  // bool Optimizer::optimize(...) — optimizer.hpp:167-170. `path` and `cmds` are the trajectorizer's seed and come
  // back optimised; `people_proj` receives the SFM projection of the (first 3) people; false <=> keep the seed.
  bool optimize(Path& path, AgentsTrajectories& people_proj, const Costmap2D* costmap, const ObstacleDistance& obstacles,
                std::vector<TwistStamped>& cmds, const People& people, const Twist& speed, const float time_step)
  {
    if (!handle_) throw std::runtime_error("Optimizer::initialize has not been called");
    const int n_poses = static_cast<int>(path.poses.size());
    const int n_cmds = static_cast<int>(cmds.size());
    const int cap = (n_poses > n_cmds ? n_poses : n_cmds) + 2;
    poses_buf_.assign(3 * static_cast<size_t>(cap), 0.0);
    cmds_buf_.assign(2 * static_cast<size_t>(cap), 0.0);
    proj_buf_.assign(18 * static_cast<size_t>(cap), 0.0);
    for (int i = 0; i < n_poses; ++i) {
      poses_buf_[3 * i] = path.poses[i].pose.position.x;
      poses_buf_[3 * i + 1] = path.poses[i].pose.position.y;
      poses_buf_[3 * i + 2] = detail::get_yaw(path.poses[i].pose.orientation);
    }
    for (int i = 0; i < n_cmds; ++i) {
      cmds_buf_[2 * i] = cmds[i].twist.linear.x;
      cmds_buf_[2 * i + 1] = cmds[i].twist.angular.z;
    }
    people_buf_.resize(5 * people.people.size());
    for (size_t k = 0; k < people.people.size(); ++k) {
      people_buf_[5 * k] = people.people[k].position.x;
      people_buf_[5 * k + 1] = people.people[k].position.y;
      people_buf_[5 * k + 2] = people.people[k].velocity.x;
      people_buf_[5 * k + 3] = people.people[k].velocity.y;
      people_buf_[5 * k + 4] = people.people[k].velocity.z;
    }

    smpc_optimize_io io;
    std::memset(&io, 0, sizeof io);
    io.capacity = cap;
    io.n_poses = n_poses;
    io.poses = poses_buf_.data();
    io.n_cmds = n_cmds;
    io.cmds = cmds_buf_.data();
    io.n_people = static_cast<int>(people.people.size());
    io.people = people_buf_.empty() ? nullptr : people_buf_.data();
    io.speed_v = speed.linear.x;
    io.speed_w = speed.angular.z;
    io.time_step = time_step;
    io.costmap = costmap->getCharMap();
    io.size_x = static_cast<int>(costmap->getSizeInCellsX());
    io.size_y = static_cast<int>(costmap->getSizeInCellsY());
    io.origin_x = costmap->getOriginX();
    io.origin_y = costmap->getOriginY();
    io.resolution = costmap->getResolution();
    io.od.width = obstacles.info.width;
    io.od.height = obstacles.info.height;
    io.od.resolution = obstacles.info.resolution;
    io.od.origin_x = obstacles.info.origin.position.x;
    io.od.origin_y = obstacles.info.origin.position.y;
    io.od.distances = obstacles.distances.empty() ? nullptr : obstacles.distances.data();
    io.od.indexes = obstacles.indexes.empty() ? nullptr : obstacles.indexes.data();
    io.people_proj = proj_buf_.data();

    const int rc = smpc_optimize(handle_, &io);
    if (rc != SMPC_OK) throw std::runtime_error(smpc_last_error());  // computeObstacle's runtime_error, CUDA errors
    last_termination_ = io.termination;
    last_iterations_ = io.iterations;
    last_cost_initial_ = io.cost_initial;
    last_cost_final_ = io.cost_final;

    // people_proj is filled before the solve in the reference (src/optimizer.cpp:186) and stays valid on failure
    people_proj.assign(static_cast<size_t>(io.n_proj_steps), AgentsStates(3));
    for (int i = 0; i < io.n_proj_steps; ++i)
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 6; ++c) people_proj[i][k][c] = proj_buf_[(static_cast<size_t>(i) * 3 + k) * 6 + c];
    if (!io.optimized) return false;  // src/optimizer.cpp:384-388

    // src/optimizer.cpp:411-446
    cmds.resize(static_cast<size_t>(io.n_cmds));
    for (int i = 0; i < io.n_cmds; ++i) {
      cmds[i].header = path.header;
      cmds[i].twist.linear.x = cmds_buf_[2 * i];
      cmds[i].twist.linear.y = 0.0;
      cmds[i].twist.angular.z = cmds_buf_[2 * i + 1];
    }
    path.poses.clear();
    PoseStamped pose;
    pose.header = path.header;
    for (int i = 0; i < io.n_poses; ++i) {
      pose.pose.position.x = poses_buf_[3 * i];
      pose.pose.position.y = poses_buf_[3 * i + 1];
      detail::set_yaw(pose.pose.orientation, poses_buf_[3 * i + 2]);
      path.poses.push_back(pose);
    }
    return true;
  }

  // Fresh TrajectoryMemory (trajectory_memory.hpp): forget the previous tick's path / cmds.
  void reset_memory()
  {
    if (handle_) smpc_reset_memory(handle_);
  }

  // Solver::Summary of the last optimize() call (the reference only logs it under debug_optimizer).
  int last_termination() const { return last_termination_; }
  int last_iterations() const { return last_iterations_; }
  double last_initial_cost() const { return last_cost_initial_; }
  double last_final_cost() const { return last_cost_final_; }
  smpc_handle* handle() const { return handle_; }
  const OptimizerParams& params() const { return params_; }

private:
  smpc_handle* handle_ = nullptr;
  OptimizerParams params_;
  std::vector<double> poses_buf_, cmds_buf_, proj_buf_, people_buf_;
  int last_termination_ = SMPC_NO_CONVERGENCE;
  int last_iterations_ = 0;
  double last_cost_initial_ = 0.0, last_cost_final_ = 0.0;
};

}  // namespace nav2_social_mpc_controller_b200

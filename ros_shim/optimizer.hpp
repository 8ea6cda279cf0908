// ros_shim/optimizer.hpp — drop-in replacement of include/nav2_social_mpc_controller/optimizer.hpp for a ROS 2
// workspace that has libsmpc.so instead of Ceres. SOURCE ONLY: ROS 2 / Nav2 / people_msgs / obstacle_distance_msgs
// are not installed in the image this repo is developed in, so this file is not compiled or tested here; the same
// OptimizerT<M> template IS compiled and tested with the ROS-free message family (tests/test_cpp_host.py).
//
// SocialMPCController (src/social_mpc_controller.cpp:48-86, :240) keeps using
//   optimizer_ = std::make_unique<Optimizer>();  optimizer_params_.get(node.get(), name);
//   optimizer_->initialize(optimizer_params_);   optimizer_->optimize(traj_path, projected_people, costmap_, od, cmds,
//                                                                     people, speed, ts);
// unchanged: class names, argument lists and the bool return are those of the reference header (:59-101, :152, :167-170).
#pragma once

#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "geometry_msgs/msg/pose_stamped.hpp"
#include "geometry_msgs/msg/twist.hpp"
#include "geometry_msgs/msg/twist_stamped.hpp"
#include "nav2_costmap_2d/costmap_2d.hpp"
#include "nav2_util/node_utils.hpp"
#include "nav_msgs/msg/path.hpp"
#include "obstacle_distance_msgs/msg/obstacle_distance.hpp"
#include "people_msgs/msg/people.hpp"
#include "rclcpp_lifecycle/lifecycle_node.hpp"

#include "smpc_optimizer.hpp"

namespace nav2_social_mpc_controller
{
namespace b200 = nav2_social_mpc_controller_b200;

using AgentStatus = b200::AgentStatus;
using AgentsStates = b200::AgentsStates;
using AgentTrajectory = b200::AgentTrajectory;
using AgentsTrajectories = b200::AgentsTrajectories;

struct RosMsgs
{
  using Path = nav_msgs::msg::Path;
  using PoseStamped = geometry_msgs::msg::PoseStamped;
  using TwistStamped = geometry_msgs::msg::TwistStamped;
  using Twist = geometry_msgs::msg::Twist;
  using People = people_msgs::msg::People;
  using ObstacleDistance = obstacle_distance_msgs::msg::ObstacleDistance;
  using Costmap2D = nav2_costmap_2d::Costmap2D;
};

struct OptimizerParams : public b200::OptimizerParams
{
  // Same parameter names, defaults and error as the reference OptimizerParams::get (src/optimizer.cpp:16-85).
  void get(rclcpp_lifecycle::LifecycleNode* node, const std::string& name)
  {
    const std::string trajectorizer = name + ".trajectorizer.";
    const std::string local_name = name + ".optimizer.";
    const std::string weights = local_name + "weights.";
    auto declare = [&](const std::string& key, const rclcpp::ParameterValue& def) {
      nav2_util::declare_parameter_if_not_declared(node, key, def);
    };
    declare(local_name + "linear_solver_type", rclcpp::ParameterValue("SPARSE_NORMAL_CHOLESKY"));
    node->get_parameter(local_name + "linear_solver_type", linear_solver_type);
    static const char* const kSolverTypes[] = { "DENSE_SCHUR", "SPARSE_SCHUR", "DENSE_NORMAL_CHOLESKY", "DENSE_QR",
                                                "SPARSE_NORMAL_CHOLESKY" };
    bool known = false;
    for (const char* t : kSolverTypes) known = known || linear_solver_type == t;
    if (!known) {
      RCLCPP_ERROR(rclcpp::get_logger("optimizer"), "Invalid linear_solver_type");
      throw std::runtime_error("Invalid parameter: linear_solver_type");
    }
    declare(local_name + "param_tol", rclcpp::ParameterValue(1e-15));
    node->get_parameter(local_name + "param_tol", param_tol);
    declare(local_name + "fn_tol", rclcpp::ParameterValue(1e-7));
    node->get_parameter(local_name + "fn_tol", fn_tol);
    declare(local_name + "gradient_tol", rclcpp::ParameterValue(1e-10));
    node->get_parameter(local_name + "gradient_tol", gradient_tol);
    declare(local_name + "max_iterations", rclcpp::ParameterValue(100));
    node->get_parameter(local_name + "max_iterations", max_iterations);
    declare(local_name + "debug_optimizer", rclcpp::ParameterValue(false));
    node->get_parameter(local_name + "debug_optimizer", debug);
    declare(weights + "distance_weight", rclcpp::ParameterValue(3.0));
    node->get_parameter(weights + "distance_weight", distance_w_);
    declare(weights + "social_weight", rclcpp::ParameterValue(1.0));
    node->get_parameter(weights + "social_weight", socialwork_w_);
    declare(weights + "velocity_weight", rclcpp::ParameterValue(0.5));
    node->get_parameter(weights + "velocity_weight", velocity_w_);
    declare(weights + "angle_weight", rclcpp::ParameterValue(0.0));
    node->get_parameter(weights + "angle_weight", angle_w_);
    declare(weights + "agent_angle_weight", rclcpp::ParameterValue(0.5));
    node->get_parameter(weights + "agent_angle_weight", agent_angle_w_);
    declare(weights + "proxemics_weight", rclcpp::ParameterValue(90.0));
    node->get_parameter(weights + "proxemics_weight", proxemics_w_);
    declare(weights + "velocity_feasibility_weight", rclcpp::ParameterValue(0.5));
    node->get_parameter(weights + "velocity_feasibility_weight", velocity_feasibility_w_);
    declare(weights + "obstacle_weight", rclcpp::ParameterValue(0.0));
    node->get_parameter(weights + "obstacle_weight", obstacle_w_);
    declare(weights + "goal_align_weight", rclcpp::ParameterValue(0.0));
    node->get_parameter(weights + "goal_align_weight", goal_align_w_);
    declare(local_name + "control_horizon", rclcpp::ParameterValue(5));
    node->get_parameter(local_name + "control_horizon", control_horizon_);
    declare(local_name + "parameter_block_length", rclcpp::ParameterValue(5));
    node->get_parameter(local_name + "parameter_block_length", parameter_block_length_);
    declare(local_name + "current_path_weight", rclcpp::ParameterValue(1.0));
    node->get_parameter(local_name + "current_path_weight", current_path_w);
    declare(local_name + "current_cmds_weight", rclcpp::ParameterValue(1.0));
    node->get_parameter(local_name + "current_cmds_weight", current_cmds_w);
    node->get_parameter(trajectorizer + "max_time", max_time);
    // not a reference parameter: which GPU this controller instance uses, and the Ceres release to follow
    declare(local_name + "cuda_device", rclcpp::ParameterValue(0));
    node->get_parameter(local_name + "cuda_device", cuda_device);
    declare(local_name + "ceres_compat", rclcpp::ParameterValue(200));
    node->get_parameter(local_name + "ceres_compat", ceres_compat);
  }
  int cuda_device = 0;
};

class Optimizer : public b200::OptimizerT<RosMsgs>
{
public:
  void initialize(const OptimizerParams params)
  {
    b200::OptimizerT<RosMsgs>::initialize(params, params.cuda_device);
  }
};

}  // namespace nav2_social_mpc_controller

// Type-checks ros_shim/optimizer.hpp against the header stand-ins of tests/cpp/ros_stubs and exercises
// OptimizerParams::get(node, name) defaults (reference src/optimizer.cpp:26-84). No GPU call is made.
#include <cstdio>

#include "optimizer.hpp"

int main()
{
  rclcpp_lifecycle::LifecycleNode node;
  node.values.emplace("FollowPath.trajectorizer.max_time", rclcpp::ParameterValue(1.5));
  node.values.emplace("FollowPath.optimizer.control_horizon", rclcpp::ParameterValue(18));
  nav2_social_mpc_controller::OptimizerParams params;
  params.get(&node, "FollowPath");
  std::printf("%s %g %g %g %d %d %d %g %g %g\n", params.linear_solver_type.c_str(), params.param_tol, params.fn_tol,
              params.gradient_tol, params.max_iterations, params.control_horizon_, params.parameter_block_length_,
              params.distance_w_, params.proxemics_w_, static_cast<double>(params.max_time));
  node.values.erase("FollowPath.optimizer.linear_solver_type");
  node.values.emplace("FollowPath.optimizer.linear_solver_type", rclcpp::ParameterValue("CGNR"));
  try {
    params.get(&node, "FollowPath");
  } catch (const std::runtime_error& e) {
    std::printf("%s\n", e.what());
  }
  // the optimizer class must instantiate (members only; initialize() would need a GPU)
  bool (nav2_social_mpc_controller::Optimizer::*fn)(nav_msgs::msg::Path&, nav2_social_mpc_controller::AgentsTrajectories&,
                                                    const nav2_costmap_2d::Costmap2D*,
                                                    const obstacle_distance_msgs::msg::ObstacleDistance&,
                                                    std::vector<geometry_msgs::msg::TwistStamped>&,
                                                    const people_msgs::msg::People&, const geometry_msgs::msg::Twist&,
                                                    const float) = &nav2_social_mpc_controller::Optimizer::optimize;
  return fn == nullptr;
}

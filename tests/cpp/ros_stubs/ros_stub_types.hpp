#pragma once
#include <cstdio>
#include <map>
#include <string>

#include "smpc_plain_msgs.hpp"

namespace geometry_msgs { namespace msg {
using PoseStamped = nav2_social_mpc_controller_b200::plain::PoseStamped;
using Twist = nav2_social_mpc_controller_b200::plain::Twist;
using TwistStamped = nav2_social_mpc_controller_b200::plain::TwistStamped;
} }
namespace nav_msgs { namespace msg { using Path = nav2_social_mpc_controller_b200::plain::Path; } }
namespace people_msgs { namespace msg { using People = nav2_social_mpc_controller_b200::plain::People; } }
namespace obstacle_distance_msgs { namespace msg {
using ObstacleDistance = nav2_social_mpc_controller_b200::plain::ObstacleDistance;
} }
namespace nav2_costmap_2d { using Costmap2D = nav2_social_mpc_controller_b200::plain::Costmap2D; }

namespace rclcpp {
struct ParameterValue {
  double d = 0; std::string s; bool is_string = false;
  ParameterValue(double v) : d(v) {}
  ParameterValue(int v) : d(v) {}
  ParameterValue(bool v) : d(v ? 1 : 0) {}
  ParameterValue(const char* v) : s(v), is_string(true) {}
};
struct Logger {};
inline Logger get_logger(const char*) { return Logger(); }
}  // namespace rclcpp
#define RCLCPP_ERROR(logger, ...) do { (void)(logger); std::fprintf(stderr, __VA_ARGS__); } while (0)

namespace rclcpp_lifecycle {
class LifecycleNode {
public:
  std::map<std::string, rclcpp::ParameterValue> values;
  template <class T>
  bool get_parameter(const std::string& key, T& out) const {
    auto it = values.find(key);
    if (it == values.end()) return false;
    out = static_cast<T>(it->second.d);
    return true;
  }
  bool get_parameter(const std::string& key, std::string& out) const {
    auto it = values.find(key);
    if (it == values.end()) return false;
    out = it->second.s;
    return true;
  }
};
}  // namespace rclcpp_lifecycle
namespace nav2_util {
inline void declare_parameter_if_not_declared(rclcpp_lifecycle::LifecycleNode* node, const std::string& key,
                                              const rclcpp::ParameterValue& v) {
  node->values.emplace(key, v);
}
}  // namespace nav2_util

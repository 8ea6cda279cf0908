// TEST INFRASTRUCTURE — stand-in for tf2_geometry_msgs: tf2::toMsg(Quaternion).
#pragma once
#include "geometry_msgs/msg/pose.hpp"
#include "tf2/LinearMath/Quaternion.h"
namespace tf2 {
inline geometry_msgs::msg::Quaternion toMsg(const Quaternion& q) {
  geometry_msgs::msg::Quaternion m;
  m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w();
  return m;
}
}

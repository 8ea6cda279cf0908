// TEST INFRASTRUCTURE — stand-in for tf2/utils.h: tf2::getYaw (tf2 impl::getYaw, generic branch and the gimbal-lock one).
#pragma once
#include <cmath>
#include "geometry_msgs/msg/pose.hpp"
namespace tf2 {
template <class Q>
inline double getYaw(const Q& q) {
  const double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
  const double sarg = -2 * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
  if (sarg <= -0.99999) return -2 * std::atan2(q.y, q.x);
  if (sarg >= 0.99999) return 2 * std::atan2(q.y, q.x);
  return std::atan2(2 * (q.x * q.y + q.w * q.z), sqw + sqx - sqy - sqz);
}
}

// TEST INFRASTRUCTURE — ROS-free stand-in for geometry_msgs/msg/twist.hpp.
#pragma once
namespace geometry_msgs { namespace msg {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Twist { Vector3 linear, angular; };
} }

// smpc_plain_msgs.hpp — ROS-free message family for OptimizerT<M> (include/smpc_optimizer.hpp).
//
// PODs with the member names of the ROS 2 messages the reference optimizer reads and writes
// (nav_msgs/Path, geometry_msgs/{PoseStamped,Twist,TwistStamped}, people_msgs/{People,Person},
// obstacle_distance_msgs/ObstacleDistance, nav2_costmap_2d::Costmap2D accessors). ROS 2 is not installed in this
// image; inside a ROS workspace use ros_shim/ros_msgs.hpp instead, the optimizer source is the same.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace nav2_social_mpc_controller_b200
{
namespace plain
{
struct Time { int32_t sec = 0; uint32_t nanosec = 0; };
struct Header { Time stamp; std::string frame_id; };
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { Header header; Pose pose; };
struct Path { Header header; std::vector<PoseStamped> poses; };
struct Twist { Vector3 linear; Vector3 angular; };
struct TwistStamped { Header header; Twist twist; };
struct Person { std::string name; Point position; Point velocity; double reliability = 0; };  // people_msgs/Person
struct People { Header header; std::vector<Person> people; };
struct MapMetaData { float resolution = 0; uint32_t width = 0, height = 0; Pose origin; };
struct ObstacleDistance  // obstacle_distance_msgs/ObstacleDistance
{
  Header header;
  MapMetaData info;
  std::vector<float> distances;
  std::vector<uint32_t> indexes;
};

// The accessors of nav2_costmap_2d::Costmap2D the optimizer uses (src/optimizer.cpp:167-170, obstacle critic).
class Costmap2D
{
public:
  Costmap2D(unsigned size_x, unsigned size_y, double resolution, double origin_x, double origin_y,
            const unsigned char* data = nullptr)
    : size_x_(size_x), size_y_(size_y), resolution_(resolution), origin_x_(origin_x), origin_y_(origin_y),
      cells_(static_cast<size_t>(size_x) * size_y, 0)
  {
    if (data) cells_.assign(data, data + cells_.size());
  }
  unsigned char* getCharMap() const { return const_cast<unsigned char*>(cells_.data()); }
  unsigned getSizeInCellsX() const { return size_x_; }
  unsigned getSizeInCellsY() const { return size_y_; }
  double getOriginX() const { return origin_x_; }
  double getOriginY() const { return origin_y_; }
  double getResolution() const { return resolution_; }
  bool worldToMap(double wx, double wy, unsigned& mx, unsigned& my) const
  {
    if (wx < origin_x_ || wy < origin_y_) return false;
    mx = static_cast<unsigned>((wx - origin_x_) / resolution_);
    my = static_cast<unsigned>((wy - origin_y_) / resolution_);
    return mx < size_x_ && my < size_y_;
  }

private:
  unsigned size_x_, size_y_;
  double resolution_, origin_x_, origin_y_;
  std::vector<unsigned char> cells_;
};
}  // namespace plain

struct PlainMsgs
{
  using Path = plain::Path;
  using PoseStamped = plain::PoseStamped;
  using TwistStamped = plain::TwistStamped;
  using Twist = plain::Twist;
  using People = plain::People;
  using ObstacleDistance = plain::ObstacleDistance;
  using Costmap2D = plain::Costmap2D;
};

}  // namespace nav2_social_mpc_controller_b200

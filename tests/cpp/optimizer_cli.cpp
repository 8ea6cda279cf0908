// optimizer_cli.cpp — test driver for the C++ host mirror (include/smpc_optimizer.hpp) with the ROS-free message
// family: reads one scene from a text file, runs n_ticks calls of Optimizer::optimize (same in-out semantics as the
// reference, optimizer.hpp:167-170) and prints every tick's outputs with 17 significant digits.
// Built and run by tests/test_cpp_host.py; not part of the product library.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "smpc_optimizer.hpp"
#include "smpc_plain_msgs.hpp"

using namespace nav2_social_mpc_controller_b200;
using Optimizer = OptimizerT<PlainMsgs>;

int main(int argc, char** argv)
{
  if (argc < 2) {
    std::fprintf(stderr, "usage: optimizer_cli scene.txt\n");
    return 2;
  }
  std::ifstream in(argv[1]);
  std::string yaml, plugin;
  int n_ticks = 0, n_poses = 0, n_cmds = 0, n_people = 0;
  in >> yaml >> plugin >> n_ticks;
  plain::Path seed_path;
  seed_path.header.frame_id = "odom";
  in >> n_poses;
  for (int i = 0; i < n_poses; ++i) {
    plain::PoseStamped p;
    double yaw;
    in >> p.pose.position.x >> p.pose.position.y >> yaw;
    detail::set_yaw(p.pose.orientation, yaw);
    seed_path.poses.push_back(p);
  }
  std::vector<plain::TwistStamped> seed_cmds;
  in >> n_cmds;
  for (int i = 0; i < n_cmds; ++i) {
    plain::TwistStamped c;
    in >> c.twist.linear.x >> c.twist.angular.z;
    seed_cmds.push_back(c);
  }
  plain::People people;
  in >> n_people;
  for (int k = 0; k < n_people; ++k) {
    plain::Person p;
    in >> p.position.x >> p.position.y >> p.velocity.x >> p.velocity.y >> p.velocity.z;
    people.people.push_back(p);
  }
  plain::Twist speed;
  float time_step;
  in >> speed.linear.x >> speed.angular.z >> time_step;
  unsigned size_x, size_y;
  double ox, oy, res;
  in >> size_x >> size_y >> ox >> oy >> res;
  std::vector<unsigned char> cells(static_cast<size_t>(size_x) * size_y);
  for (auto& c : cells) {
    int v;
    in >> v;
    c = static_cast<unsigned char>(v);
  }
  plain::Costmap2D costmap(size_x, size_y, res, ox, oy, cells.data());
  plain::ObstacleDistance od;
  size_t n_od = 0;
  in >> od.info.width >> od.info.height >> od.info.resolution >> od.info.origin.position.x >> od.info.origin.position.y >> n_od;
  od.distances.resize(n_od);
  od.indexes.resize(n_od);
  for (auto& d : od.distances) in >> d;
  for (auto& i : od.indexes) in >> i;
  if (!in) {
    std::fprintf(stderr, "scene file is truncated\n");
    return 2;
  }

  try {
    OptimizerParams params;
    params.get(yaml, plugin);
    Optimizer optimizer;
    optimizer.initialize(params);
    for (int t = 0; t < n_ticks; ++t) {
      plain::Path path = seed_path;
      std::vector<plain::TwistStamped> cmds = seed_cmds;
      AgentsTrajectories proj;
      const bool ok = optimizer.optimize(path, proj, &costmap, od, cmds, people, speed, time_step);
      std::printf("tick %d ok %d termination %d iterations %d cost_initial %.17g cost_final %.17g\n", t, ok ? 1 : 0,
                  optimizer.last_termination(), optimizer.last_iterations(), optimizer.last_initial_cost(),
                  optimizer.last_final_cost());
      std::printf("path %zu", path.poses.size());
      for (auto& p : path.poses)
        std::printf(" %.17g %.17g %.17g", p.pose.position.x, p.pose.position.y, detail::get_yaw(p.pose.orientation));
      std::printf("\ncmds %zu", cmds.size());
      for (auto& c : cmds) std::printf(" %.17g %.17g", c.twist.linear.x, c.twist.angular.z);
      std::printf("\nproj %zu", proj.size());
      for (auto& step : proj)
        for (auto& a : step)
          for (double v : a) std::printf(" %.17g", v);
      std::printf("\n");
    }
  } catch (const std::exception& e) {
    std::printf("exception %s\n", e.what());
    return 1;
  }
  return 0;
}

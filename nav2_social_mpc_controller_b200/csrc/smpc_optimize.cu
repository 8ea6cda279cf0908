// smpc_optimize.cu — level-2 entry of libsmpc.so for ONE robot: bool Optimizer::optimize(...)
// (reference include/nav2_social_mpc_controller/optimizer.hpp:167-170, src/optimizer.cpp:148-452).
// Since round 2 this is the fleet tick (smpc_optimize_batch, smpc_project.cu) called for a fleet of one: every stage —
// TrajectoryMemory seeding, people_to_status, format_to_optimize, project_people (SFM), the bounded TR-LM solve, the
// post-solve expansion, the memory update — runs in the same GPU kernels a fleet uses. This file only converts the
// reference-shaped argument block (step-major people_proj, capacity-sized arrays) and keeps its error behaviour.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smpc.h"
#include "smpc_host_state.h"

extern "C" {

int smpc_reset_memory(smpc_handle* h) {
  if (!h) return smpc_host_fail(SMPC_ERR_ARGUMENT, "handle is NULL");
  for (smpc_fleet_state* fs : {smpc_handle_fleet(h), smpc_handle_single(h)}) {
    std::lock_guard<std::mutex> lk(fs->mu);
    fs->forget();  // the next tick re-creates (zeroes) the per-robot memory: a fresh TrajectoryMemory
  }
  return SMPC_OK;
}

int smpc_optimize(smpc_handle* h, smpc_optimize_io* io) {
  if (!h || !io) return smpc_host_fail(SMPC_ERR_ARGUMENT, "NULL argument");
  if (!io->poses || !io->cmds || !io->costmap) return smpc_host_fail(SMPC_ERR_ARGUMENT, "poses / cmds / costmap missing");
  io->optimized = 0;
  io->n_proj_steps = 0;
  if (io->n_poses < 2) return SMPC_OK;  // "Path has less than 2 points" -> return false (:158-162)
  if (io->n_cmds < io->n_poses - 1) return smpc_host_fail(SMPC_ERR_ARGUMENT, "fewer cmds than poses - 1");
  if (io->capacity < io->n_poses) return smpc_host_fail(SMPC_ERR_ARGUMENT, "capacity smaller than the number of poses");
  // computeObstacle's std::runtime_error cases that do not depend on the people (:676-700)
  if (!io->od.indexes || !io->od.distances) return smpc_host_fail(SMPC_ERR_ARGUMENT, "ObstacleDistance is empty");

  const int cap = io->capacity, A = 3;  // the reference keeps exactly three people columns (:468-479)
  // Array stride of the one-robot fleet: constant from tick to tick (the trajectorizer never emits more than
  // max_time / time_step + 1 poses), so that the warm-start memory survives a change of the path length or of the
  // caller's buffer capacity — the reference's TrajectoryMemory does (ADVICE).
  const smpc_params* prm = smpc_handle_params(h);
  const float ts = io->time_step > 0.0f ? io->time_step : static_cast<float>(prm->time_step);
  const int stride = std::max(io->n_poses, static_cast<int>(std::round(static_cast<float>(prm->max_time) / ts)) + 1);
  std::vector<double> pose_rows((size_t)stride * 3, 0.0), cmd_rows((size_t)stride * 2, 0.0);
  std::memcpy(pose_rows.data(), io->poses, sizeof(double) * 3 * io->n_poses);
  std::memcpy(cmd_rows.data(), io->cmds, sizeof(double) * 2 * std::min(io->n_cmds, stride));
  std::vector<double> people((size_t)A * 5, 0.0);
  const int32_t n_people = io->n_people < A ? io->n_people : A;
  if (n_people > 0) std::memcpy(people.data(), io->people, sizeof(double) * 5 * n_people);
  // has_people follows people.people.size() != 0 of the UNtruncated list (:263): same thing for n_people >= 1
  const int32_t n_poses = io->n_poses;
  const double speed[2] = {io->speed_v, io->speed_w};
  const double origin[2] = {io->origin_x, io->origin_y}, od_origin[2] = {io->od.origin_x, io->od.origin_y};
  int32_t n_out = 0, termination = 0, iterations = 0, status = 0;
  uint8_t optimized = 0;
  double c0 = 0.0, c1 = 0.0;
  std::vector<double> proj(io->people_proj ? (size_t)A * 6 * stride : 0);

  smpc_fleet_io f;
  std::memset(&f, 0, sizeof(f));
  f.n_robots = 1;
  f.max_poses = stride;
  f.n_agents = A;
  f.time_step = io->time_step;
  f.n_poses = &n_poses;
  f.people = people.data();
  f.n_people = &n_people;
  f.speed = speed;
  f.costmaps = io->costmap;
  f.costmap_origin = origin;
  f.n_costmaps = 1;
  f.size_x = io->size_x;
  f.size_y = io->size_y;
  f.resolution = io->resolution;
  f.od_indexes = io->od.indexes;
  f.od_origin = od_origin;
  f.n_od_grids = 1;
  f.od_width = io->od.width;
  f.od_height = io->od.height;
  f.od_resolution = io->od.resolution;
  f.maps_version = 0;  // one robot: the maps travel with every call, as in the reference
  f.poses = pose_rows.data();
  f.cmds = cmd_rows.data();
  f.n_out = &n_out;
  f.optimized = &optimized;
  f.termination = &termination;
  f.iterations = &iterations;
  f.cost_initial = &c0;
  f.cost_final = &c1;
  f.project_status = &status;
  f.people_proj = io->people_proj ? proj.data() : nullptr;
  const int rc = smpc_optimize_batch_on(h, smpc_handle_single(h), &f);
  if (rc != SMPC_OK) return rc;
  if (status != 0)  // the reference throws std::runtime_error from computeObstacle (:707-713)
    return smpc_host_fail(SMPC_ERR_ARGUMENT, "Agent position is outside the ObstacleDistance grid");
  io->termination = termination;
  io->iterations = iterations;
  io->cost_initial = c0;
  io->cost_final = c1;
  if (n_out > cap) return smpc_host_fail(SMPC_ERR_ARGUMENT, "capacity smaller than the optimised path");
  std::memcpy(io->poses, pose_rows.data(), sizeof(double) * 3 * n_out);
  std::memcpy(io->cmds, cmd_rows.data(), sizeof(double) * 2 * n_out);
  io->n_poses = n_out;
  io->n_proj_steps = n_out;
  if (io->people_proj)  // [A][6][cap] -> the reference's AgentsTrajectories[step][agent][6]
    for (int i = 0; i < n_out; ++i)
      for (int k = 0; k < A; ++k)
        for (int c = 0; c < 6; ++c) io->people_proj[((size_t)i * A + k) * 6 + c] = proj[((size_t)k * 6 + c) * stride + i];
  if (optimized) {
    io->n_cmds = n_out;
    io->optimized = 1;
  }
  return SMPC_OK;
}

}  // extern "C"

// smpc_multi.cu — in-library multi-GPU dispatch of the level-1 solve: problems are independent, so a host
// batch is cut into contiguous shards, one per GPU; one host thread + one handle (own streams, own staging) per GPU
// runs the ordinary host-buffer pipeline (smpc_solve_batch) on its shard, reading from and writing to the caller's
// host arrays at the shard's offset — the "final host gather" is the shards' own D2H copies landing side by side in
// the caller's (ideally page-locked) result arrays. No collective, no peer traffic: nothing on the solve path crosses
// GPUs. Shard boundaries are multiples of `granule` (e.g. the number of starts per robot for multi-start batches, so a
// robot's arg-min never spans two GPUs).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "../../include/smpc.h"
#include "smpc_host_state.h"

struct smpc_multi {
  std::vector<smpc_handle*> handles;
  std::vector<int> devices;
  smpc_params params;
  std::vector<int32_t> index_scratch;  // explicit costmap index when shared maps are addressed by b % M
  double last_ms = 0.0;
};

namespace {

// [lo, hi) of shard r: contiguous, multiples of `granule`, sizes differing by at most one granule.
void shard_bounds(long long n, int world, int r, long long granule, long long* lo, long long* hi) {
  const long long units = n / granule;
  const long long base = units / world, extra = units % world;
  const long long lo_u = r * base + std::min<long long>(r, extra);
  const long long hi_u = lo_u + base + (r < extra ? 1 : 0);
  *lo = lo_u * granule;
  *hi = (r == world - 1) ? n : hi_u * granule;  // a ragged tail (n not a multiple of granule) goes to the last shard
}

}  // namespace

extern "C" {

int smpc_debug_shard_bounds(int n_problems, int n_shards, int shard, int granule, int* lo, int* hi) {
  if (n_problems < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || granule < 1 || !lo || !hi)
    return smpc_host_fail(SMPC_ERR_ARGUMENT, "bad shard arguments");
  long long a = 0, b = 0;
  shard_bounds(n_problems, n_shards, shard, granule, &a, &b);
  *lo = static_cast<int>(a);
  *hi = static_cast<int>(b);
  return SMPC_OK;
}

int smpc_multi_create(const smpc_params* p, int n_devices, const int* devices, smpc_multi** out) {
  if (!p || !out || n_devices < 1) return smpc_host_fail(SMPC_ERR_ARGUMENT, "multi_create: bad arguments");
  smpc_multi* m = new smpc_multi();
  m->params = *p;
  for (int i = 0; i < n_devices; ++i) {
    const int dev = devices ? devices[i] : i;  // the same device may appear twice (two handles, two stream pairs)
    smpc_handle* h = nullptr;
    const int rc = smpc_create(p, dev, &h);
    if (rc != SMPC_OK) {
      const std::string msg = smpc_last_error();
      smpc_multi_destroy(m);
      return smpc_host_fail(rc, "multi_create, device " + std::to_string(dev) + ": " + msg);
    }
    m->handles.push_back(h);
    m->devices.push_back(dev);
  }
  *out = m;
  return SMPC_OK;
}

void smpc_multi_destroy(smpc_multi* m) {
  if (!m) return;
  for (smpc_handle* h : m->handles) smpc_destroy(h);
  delete m;
}

int smpc_multi_device_count(smpc_multi* m) { return m ? static_cast<int>(m->handles.size()) : 0; }

int smpc_solve_batch_multi(smpc_multi* m, const smpc_batch* in, smpc_result* out, int granule) {
  if (!m || !in || !out) return smpc_host_fail(SMPC_ERR_ARGUMENT, "NULL argument");
  if (granule < 1) granule = 1;
  const long long B = in->n_problems;
  if (B < 0) return smpc_host_fail(SMPC_ERR_ARGUMENT, "n_problems < 0");
  if (B == 0) return SMPC_OK;
  const int world = static_cast<int>(m->handles.size());
  int nb = 0;
  int rc = smpc_problem_dims(&m->params, in->n_steps, nullptr, nullptr, &nb, nullptr);
  if (rc != SMPC_OK) return rc;
  const size_t dof = m->params.omni_solve ? 3 : 2;
  const size_t S1 = static_cast<size_t>(in->n_steps) + 1, A = in->agents ? in->n_agents : 0, P = dof * nb;
  const size_t cells = static_cast<size_t>(in->size_x) * in->size_y;
  const bool scenes = in->scenario_index != nullptr;  // per-scene arrays are shared by all shards, whole
  const bool maps_per_problem = !scenes && in->costmap_index == nullptr && in->n_costmaps == B;
  // shared maps addressed by b % M: a shard starting at `lo` would see map (b - lo) % M, so give every problem its
  // explicit index (cheap: 4 bytes per problem)
  const int32_t* index = in->costmap_index;
  if (!scenes && !maps_per_problem && index == nullptr) {
    m->index_scratch.resize(B);
    for (long long b = 0; b < B; ++b) m->index_scratch[b] = static_cast<int32_t>(b % in->n_costmaps);
    index = m->index_scratch.data();
  }
  std::vector<int> codes(world, SMPC_OK);
  std::vector<std::string> messages(world);
  std::vector<std::thread> pool;
  for (int r = 0; r < world; ++r) {
    long long lo = 0, hi = 0;
    shard_bounds(B, world, r, granule, &lo, &hi);
    if (hi <= lo) continue;
    pool.emplace_back([&, r, lo, hi]() {
      smpc_batch s = *in;
      s.n_problems = static_cast<int>(hi - lo);
      s.u0 = in->u0 + P * lo;
      if (scenes) {
        s.scenario_index = in->scenario_index + lo;
      } else {
        s.pose0 = in->pose0 + 3 * lo;
        s.path_xy = in->path_xy + 2 * S1 * lo;
        s.goal_yaw = in->goal_yaw + lo;
        s.agents = A ? in->agents + A * 6 * S1 * lo : nullptr;
        s.has_people = in->has_people ? in->has_people + lo : nullptr;
        s.n_steps_each = in->n_steps_each ? in->n_steps_each + lo : nullptr;
      }
      if (scenes) {
        s.costmap_index = in->costmap_index;  // per scene row, like every other per-scene array: unchanged
      } else if (maps_per_problem) {
        s.costmaps = in->costmaps + cells * lo;
        s.costmap_origin = in->costmap_origin + 2 * lo;
        s.n_costmaps = s.n_problems;
        s.costmap_index = nullptr;
      } else {
        s.costmap_index = index + lo;
      }
      smpc_result o = *out;
      if (out->u) o.u = out->u + P * lo;
      if (out->cmds) o.cmds = out->cmds + S1 * dof * lo;
      if (out->path) o.path = out->path + S1 * 3 * lo;
      if (out->cost_initial) o.cost_initial = out->cost_initial + lo;
      if (out->cost_final) o.cost_final = out->cost_final + lo;
      if (out->iterations) o.iterations = out->iterations + lo;
      if (out->termination) o.termination = out->termination + lo;
      if (out->usable) o.usable = out->usable + lo;
      if (out->n_evals) o.n_evals = out->n_evals + 2 * lo;
      if (out->trace) o.trace = out->trace + static_cast<size_t>(out->trace_rows) * 8 * lo;
      codes[r] = smpc_solve_batch(m->handles[r], &s, &o);
      if (codes[r] != SMPC_OK) messages[r] = smpc_last_error();  // the error text is thread-local: carry it out
    });
  }
  for (std::thread& t : pool) t.join();
  for (int r = 0; r < world; ++r)
    if (codes[r] != SMPC_OK)
      return smpc_host_fail(codes[r], "shard " + std::to_string(r) + " (device " + std::to_string(m->devices[r]) + "): " + messages[r]);
  return SMPC_OK;
}

}  // extern "C"

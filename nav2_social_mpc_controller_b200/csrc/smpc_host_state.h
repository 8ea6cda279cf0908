// smpc_host_state.h — internal accessors shared by the host translation units of libsmpc.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/smpc.h"

// per-handle stand-in for the reference's process-wide TrajectoryMemory singleton (trajectory_memory.hpp)
struct smpc_memory {
  std::vector<double> prev_poses;  // [n][3]
  std::vector<double> prev_cmds;   // [n][2]
};

// growing device allocation
struct smpc_devbuf {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// growing page-locked host allocation
struct smpc_pinbuf {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaHostAlloc(&ptr, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// device-resident state of the fleet tick (smpc_optimize_batch): per-robot TrajectoryMemory, cached maps, scratch
struct smpc_fleet_state {
  std::mutex mu;
  smpc_devbuf scratch, maps, memory;
  smpc_pinbuf pin_in, pin_out, pin_maps;  // staging of small ticks: one copy per direction (see smpc_optimize_batch_on)
  int mem_robots = 0, mem_stride = 0;
  long long maps_version = -1;
  size_t maps_bytes = 0;
  std::vector<int32_t> host_n_each, host_s_each;
  void forget() { mem_robots = mem_stride = 0; }
  void release() {
    scratch.release();
    maps.release();
    memory.release();
    pin_in.release();
    pin_out.release();
    pin_maps.release();
    forget();
    maps_version = -1;
  }
};

const smpc_params* smpc_handle_params(smpc_handle* h);
smpc_fleet_state* smpc_handle_fleet(smpc_handle* h);
smpc_fleet_state* smpc_handle_single(smpc_handle* h);
int smpc_handle_device(smpc_handle* h);
smpc_memory* smpc_handle_memory(smpc_handle* h);
int smpc_host_fail(int code, const std::string& msg);
cudaStream_t smpc_handle_stream(smpc_handle* h);
void smpc_handle_count_launch(smpc_handle* h);
// the fleet tick on an explicit state (smpc_optimize_batch uses the handle's fleet state, smpc_optimize its own)
int smpc_optimize_batch_on(smpc_handle* h, smpc_fleet_state* fs, smpc_fleet_io* io);

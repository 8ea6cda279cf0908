// smpc_host_state.h — internal accessors shared by the host translation units of libsmpc.so.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/smpc.h"

// per-handle stand-in for the reference's process-wide TrajectoryMemory singleton (trajectory_memory.hpp)
struct smpc_memory {
  std::vector<double> prev_poses;  // [n][3]
  std::vector<double> prev_cmds;   // [n][2]
};

const smpc_params* smpc_handle_params(smpc_handle* h);
smpc_memory* smpc_handle_memory(smpc_handle* h);
int smpc_host_fail(int code, const std::string& msg);
cudaStream_t smpc_handle_stream(smpc_handle* h);
void smpc_handle_count_launch(smpc_handle* h);

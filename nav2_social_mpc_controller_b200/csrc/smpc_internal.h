// smpc_internal.h — launch interface between the host library (smpc_host.cu) and the kernels (smpc_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smpc {
struct DevParams;
struct DevBatch;
struct DevResult;
struct DevEvalOut;

cudaError_t launch_solve(const DevParams& prm, const DevBatch& bt, const DevResult& rs, int* queue, int n_sm,
                         int forced_group, int forced_warps, cudaStream_t stream);
cudaError_t launch_eval(const DevParams& prm, const DevBatch& bt, const double* x, const DevEvalOut& eo, int n_sm,
                        int forced_group, cudaStream_t stream);
cudaError_t launch_argmin(int n_robots, int n_starts, int n_params, const double* cost_final, const uint8_t* usable,
                          const double* u, int32_t* best_index, double* best_cost, double* best_u, cudaStream_t stream);
int max_supported_blocks();
cudaError_t launch_pack_agents(long long n_rows, int S1, const double* agents, double* packed, uint8_t* valid,
                               cudaStream_t stream);
cudaError_t launch_polymin(int n, const double* in, double* out, cudaStream_t stream);
cudaError_t launch_math(int kind, int n, const double* in, double* out, cudaStream_t stream);
cudaError_t launch_dfma_peak(double* sink, int n_sm, int iters, cudaStream_t stream);
}  // namespace smpc

// smpc_kernels.cu — sm_100a kernels of libsmpc.so and their launchers.
//   smpc_solve_kernel<NB,G,W> (smpc_kernels_nb.inc) persistent groups pull problems from an atomic queue and run the whole bounded
//                            TR-LM solve (replaces ceres::Solve, reference src/optimizer.cpp:381) plus the
//                            post-solve expansion of reference src/optimizer.cpp:390-446.
//   smpc_eval_kernel<NB>     one evaluation (cost, J^T r, J^T J) per problem: the parity / first-slice entry.
//   smpc_argmin_kernel       per-robot arg-min over multi-start solves.
#include "smpc_device.cuh"
#include "smpc_internal.h"

namespace smpc {

#ifndef SMPC_WARPS_PER_CTA
#define SMPC_WARPS_PER_CTA 4
#endif
constexpr int kWarpsPerCta = SMPC_WARPS_PER_CTA;
constexpr int kThreads = kWarpsPerCta * 32;
#ifndef SMPC_MIN_CTAS
#define SMPC_MIN_CTAS (16 / SMPC_WARPS_PER_CTA)
#endif

// Agent records for the evaluation: (x, y, v cos yaw, v sin yaw) per agent and step, one 32-byte sector each, plus
// a validity byte. Built once per batch so that the solve never calls sincos on agent headings (61 evaluations per
// problem on average) and reads every agent with one coalesced sector instead of five strided rows.
__global__ void smpc_pack_agents_kernel(long long n_rows, int S1, const double* __restrict__ agents,
                                        double* __restrict__ packed, uint8_t* __restrict__ valid) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (problem * A + agent) * S1 + step
  if (t >= n_rows * S1) return;
  const long long row = t / S1;
  const int step = (int)(t % S1);
  const double* a = agents + row * 6 * S1 + step;
  const double yaw = a[2 * (size_t)S1], lv = a[4 * (size_t)S1];
  double sy, cy;
  sincos(yaw, &sy, &cy);
  double* o = packed + t * 4;
  o[0] = a[0];
  o[1] = a[S1];
  o[2] = lv * cy;
  o[3] = lv * sy;
  valid[t] = !(a[3 * (size_t)S1] == -1.0);
}

cudaError_t launch_pack_agents(long long n_rows, int S1, const double* agents, double* packed, uint8_t* valid,
                               cudaStream_t stream) {
  const long long n = n_rows * S1;
  smpc_pack_agents_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n_rows, S1, agents, packed, valid);
  return cudaGetLastError();
}

// Line-search polynomial minimiser exposed for unit tests: rows of (lo, hi, f0, g0, t1, f1, g1, t2, f2, g2);
// t2 <= 0 selects the two-sample (cubic) case.
__global__ void smpc_polymin_kernel(int n, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* r = in + (size_t)i * 10;
  out[i] = (r[7] > 0.0) ? quintic_interp_min(r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[0], r[1], 0, 0u, 1)
                        : cubic_interp_min(r[2], r[3], r[4], r[5], r[6], r[0], r[1]);
}

// The pair loop's own elementary functions exposed for unit tests: rows of (a, b); kind 0 exp_nonpos(a),
// 1 rsqrt_pos(a), 2 atan2_unit(a, b).
__global__ void smpc_math_kernel(int kind, int n, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = in[2 * (size_t)i], b = in[2 * (size_t)i + 1];
  out[i] = (kind == 0) ? exp_nonpos(a) : (kind == 1) ? rsqrt_pos(a) : atan2_unit(a, b);
}

// One CTA per robot: arg-min of cost_final over its n_starts consecutive solves (usable ones only; ties -> lowest index).
__global__ void smpc_argmin_kernel(int n_starts, int n_params, const double* __restrict__ cost_final,
                                   const uint8_t* __restrict__ usable, const double* __restrict__ u,
                                   int32_t* best_index, double* best_cost, double* best_u) {
  __shared__ double s_cost[32];
  __shared__ int s_idx[32];
  const int robot = blockIdx.x;
  const size_t base = (size_t)robot * n_starts;
  double bc = INFINITY;
  int bi = -1;
  for (int k = threadIdx.x; k < n_starts; k += blockDim.x) {
    const double c = cost_final[base + k];
    const bool ok = (usable == nullptr || usable[base + k]) && (c == c);
    if (ok && (c < bc || (c == bc && (bi < 0 || k < bi)))) {
      bc = c;
      bi = k;
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    const double oc = __shfl_xor_sync(kFullMask, bc, d);
    const int oi = __shfl_xor_sync(kFullMask, bi, d);
    if (oi >= 0 && (bi < 0 || oc < bc || (oc == bc && oi < bi))) {
      bc = oc;
      bi = oi;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_cost[warp] = bc;
    s_idx[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    bc = (lane < nw) ? s_cost[lane] : INFINITY;
    bi = (lane < nw) ? s_idx[lane] : -1;
    for (int d = 16; d > 0; d >>= 1) {
      const double oc = __shfl_xor_sync(kFullMask, bc, d);
      const int oi = __shfl_xor_sync(kFullMask, bi, d);
      if (oi >= 0 && (bi < 0 || oc < bc || (oc == bc && oi < bi))) {
        bc = oc;
        bi = oi;
      }
    }
    if (lane == 0) {
      best_index[robot] = (bi >= 0) ? (int32_t)(base + bi) : -1;
      best_cost[robot] = (bi >= 0) ? bc : INFINITY;
    }
    if (best_u)
      for (int c = lane; c < n_params; c += 32)
        best_u[(size_t)robot * n_params + c] = (bi >= 0) ? u[(base + bi) * n_params + c] : NAN;
  }
}

// -------------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------------
#define SMPC_DECL_NB(k)                                                                                              \
  cudaError_t launch_solve_nb##k(const DevParams&, const DevBatch&, const DevResult&, int*, int, int, int, cudaStream_t); \
  cudaError_t launch_eval_nb##k(const DevParams&, const DevBatch&, const double*, const DevEvalOut&, int, int, cudaStream_t);
SMPC_DECL_NB(1) SMPC_DECL_NB(2) SMPC_DECL_NB(3) SMPC_DECL_NB(4) SMPC_DECL_NB(5) SMPC_DECL_NB(6)
SMPC_DECL_NB(7) SMPC_DECL_NB(8) SMPC_DECL_NB(9) SMPC_DECL_NB(10) SMPC_DECL_NB(11) SMPC_DECL_NB(12)
SMPC_DECL_NB(13) SMPC_DECL_NB(14) SMPC_DECL_NB(15) SMPC_DECL_NB(16) SMPC_DECL_NB(17) SMPC_DECL_NB(18)

cudaError_t launch_solve(const DevParams& prm, const DevBatch& bt, const DevResult& rs, int* queue, int n_sm,
                         int forced_group, int forced_warps, cudaStream_t stream) {
  switch (prm.nb) {
    case 1: return launch_solve_nb1(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 2: return launch_solve_nb2(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 3: return launch_solve_nb3(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 4: return launch_solve_nb4(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 5: return launch_solve_nb5(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 6: return launch_solve_nb6(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 7: return launch_solve_nb7(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 8: return launch_solve_nb8(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 9: return launch_solve_nb9(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 10: return launch_solve_nb10(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 11: return launch_solve_nb11(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 12: return launch_solve_nb12(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 13: return launch_solve_nb13(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 14: return launch_solve_nb14(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 15: return launch_solve_nb15(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 16: return launch_solve_nb16(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 17: return launch_solve_nb17(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    case 18: return launch_solve_nb18(prm, bt, rs, queue, n_sm, forced_group, forced_warps, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_eval(const DevParams& prm, const DevBatch& bt, const double* x, const DevEvalOut& eo, int n_sm,
                        int forced_group, cudaStream_t stream) {
  switch (prm.nb) {
    case 1: return launch_eval_nb1(prm, bt, x, eo, n_sm, forced_group, stream);
    case 2: return launch_eval_nb2(prm, bt, x, eo, n_sm, forced_group, stream);
    case 3: return launch_eval_nb3(prm, bt, x, eo, n_sm, forced_group, stream);
    case 4: return launch_eval_nb4(prm, bt, x, eo, n_sm, forced_group, stream);
    case 5: return launch_eval_nb5(prm, bt, x, eo, n_sm, forced_group, stream);
    case 6: return launch_eval_nb6(prm, bt, x, eo, n_sm, forced_group, stream);
    case 7: return launch_eval_nb7(prm, bt, x, eo, n_sm, forced_group, stream);
    case 8: return launch_eval_nb8(prm, bt, x, eo, n_sm, forced_group, stream);
    case 9: return launch_eval_nb9(prm, bt, x, eo, n_sm, forced_group, stream);
    case 10: return launch_eval_nb10(prm, bt, x, eo, n_sm, forced_group, stream);
    case 11: return launch_eval_nb11(prm, bt, x, eo, n_sm, forced_group, stream);
    case 12: return launch_eval_nb12(prm, bt, x, eo, n_sm, forced_group, stream);
    case 13: return launch_eval_nb13(prm, bt, x, eo, n_sm, forced_group, stream);
    case 14: return launch_eval_nb14(prm, bt, x, eo, n_sm, forced_group, stream);
    case 15: return launch_eval_nb15(prm, bt, x, eo, n_sm, forced_group, stream);
    case 16: return launch_eval_nb16(prm, bt, x, eo, n_sm, forced_group, stream);
    case 17: return launch_eval_nb17(prm, bt, x, eo, n_sm, forced_group, stream);
    case 18: return launch_eval_nb18(prm, bt, x, eo, n_sm, forced_group, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_argmin(int n_robots, int n_starts, int n_params, const double* cost_final, const uint8_t* usable,
                          const double* u, int32_t* best_index, double* best_cost, double* best_u, cudaStream_t stream) {
  int threads = 32;
  while (threads < n_starts && threads < 256) threads <<= 1;
  smpc_argmin_kernel<<<n_robots, threads, 0, stream>>>(n_starts, n_params, cost_final, usable, u, best_index, best_cost,
                                                       best_u);
  return cudaGetLastError();
}

int max_supported_blocks() { return 18; }

cudaError_t launch_math(int kind, int n, const double* in, double* out, cudaStream_t stream) {
  smpc_math_kernel<<<(n + 127) / 128, 128, 0, stream>>>(kind, n, in, out);
  return cudaGetLastError();
}

cudaError_t launch_polymin(int n, const double* in, double* out, cudaStream_t stream) {
  smpc_polymin_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, in, out);
  return cudaGetLastError();
}

// DFMA-saturating microbenchmark: 8 independent FMA chains per thread (roofline denominator, "of measured").
__global__ void __launch_bounds__(256) smpc_dfma_peak_kernel(double* sink, int iters, double a, double b) {
  double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
      r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
  }
  const double r = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  if (r == 123.456) sink[0] = r;  // never true; keeps the chains alive
}

cudaError_t launch_dfma_peak(double* sink, int n_sm, int iters, cudaStream_t stream) {
  smpc_dfma_peak_kernel<<<n_sm * 8, 256, 0, stream>>>(sink, iters, 0.999999, 1e-7);
  return cudaGetLastError();
}

}  // namespace smpc

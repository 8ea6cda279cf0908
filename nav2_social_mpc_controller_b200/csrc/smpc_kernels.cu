// smpc_kernels.cu — sm_100a kernels of libsmpc.so and their launchers.
//   smpc_solve_kernel<NB>    persistent warps pull problems from an atomic queue and run the whole bounded
//                            TR-LM solve (replaces ceres::Solve, reference src/optimizer.cpp:381) plus the
//                            post-solve expansion of reference src/optimizer.cpp:390-446.
//   smpc_eval_kernel<NB>     one evaluation (cost, J^T r, J^T J) per problem: the parity / first-slice entry.
//   smpc_argmin_kernel       per-robot arg-min over multi-start solves.
#include "smpc_device.cuh"
#include "smpc_internal.h"

namespace smpc {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
#ifndef SMPC_MIN_CTAS
#define SMPC_MIN_CTAS 4
#endif

template <int NB, bool MULTI>
__global__ void __launch_bounds__(kThreads, SMPC_MIN_CTAS) smpc_solve_kernel(DevParams prm, DevBatch bt, DevResult rs, int* queue) {
  using L = Layout<NB>;
  constexpr int P = 2 * NB;
  __shared__ double smem[kWarpsPerCta * L::kTotal];
  const int lane = threadIdx.x & 31;
  double* ws = smem + (threadIdx.x >> 5) * L::kTotal;
  LaneConst<NB> lc;
  lane_setup<NB>(lane, prm.bl, bt.dt, lc);
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(queue, 1);
    b = __shfl_sync(kFullMask, b, 0);
    if (b >= bt.B) break;

    Prob pb;
    load_problem(bt, b, pb);
    double aa_target[kMaxChunks];
    agent_angle_setup(prm, bt, pb, lane, aa_target);
    __syncwarp();
    if (lane < P) ws[L::kX + lane] = __ldg(bt.u0 + (size_t)b * P + lane);
    __syncwarp();

    SolveOut so;
    solve_problem<NB, MULTI>(prm, bt, pb, aa_target, lc, ws, lane, so);
    const bool usable = so.termination <= kNoConvergence;  // Solver::Summary::IsSolutionUsable
    double x[P];
    SMPC_UNROLL for (int c = 0; c < P; ++c) x[c] = usable ? ws[L::kX + c] : __ldg(bt.u0 + (size_t)b * P + c);

    if (lane == 0) {
      if (rs.u) {
        SMPC_UNROLL for (int c = 0; c < P; ++c) rs.u[(size_t)b * P + c] = x[c];
      }
      if (rs.cost_initial) rs.cost_initial[b] = so.cost_initial;
      if (rs.cost_final) rs.cost_final[b] = so.cost_final;
      if (rs.iterations) rs.iterations[b] = so.iterations;
      if (rs.termination) rs.termination[b] = so.termination;
      if (rs.usable) rs.usable[b] = usable ? 1 : 0;
      if (rs.n_evals) {
        rs.n_evals[2 * b] = so.n_jac;
        rs.n_evals[2 * b + 1] = so.n_cost;
      }
    }
    // Post-solve expansion (reference src/optimizer.cpp:390-446): cmds[S+1] hold block min(i/bl, NB-1) for
    // i < ch and the last block afterwards; the path is the Euler rollout of those cmds from pose0.
    if (rs.cmds || rs.path) {
      const int S = bt.S;
      double s0, c0;
      sincos(pb.yaw0 * 0.5, &s0, &c0);
      const double yaw_rt = atan2(2.0 * (c0 * s0), c0 * c0 - s0 * s0);  // evolving_poses[0] went through setRPY/getYaw
      double carry_x = pb.x0, carry_y = pb.y0;
      for (int base = 0; base <= S; base += 32) {
        const int i = base + lane;
        const bool act = i <= S;
        const int bi = (i < prm.ch) ? min(i / prm.bl, NB - 1) : NB - 1;
        double v = x[0], w = x[1];
        double th = yaw_rt, th_next = yaw_rt;
        SMPC_UNROLL for (int bb = 0; bb < NB; ++bb) {
          if (bb == bi) {
            v = x[2 * bb];
            w = x[2 * bb + 1];
          }
          th += x[2 * bb + 1] * (bt.dt * (double)steps_in_block_before<NB>(i, bb, prm.bl));
          th_next += x[2 * bb + 1] * (bt.dt * (double)steps_in_block_before<NB>(i + 1, bb, prm.bl));
        }
        if (act && rs.cmds) {
          rs.cmds[((size_t)b * (S + 1) + i) * 2] = v;
          rs.cmds[((size_t)b * (S + 1) + i) * 2 + 1] = w;
        }
        if (rs.path) {
          double sn, cs;
          sincos(th, &sn, &cs);
          double sx = act ? v * cs * bt.dt : 0.0, sy = act ? v * sn * bt.dt : 0.0;
          SMPC_UNROLL for (int d = 1; d < 32; d <<= 1) {
            const double tx = __shfl_up_sync(kFullMask, sx, d), ty = __shfl_up_sync(kFullMask, sy, d);
            if (lane >= d) {
              sx += tx;
              sy += ty;
            }
          }
          const double X = carry_x + sx, Y = carry_y + sy;
          carry_x = __shfl_sync(kFullMask, X, 31);
          carry_y = __shfl_sync(kFullMask, Y, 31);
          if (act) {
            double sh, chh;
            sincos(th_next * 0.5, &sh, &chh);
            double* o = rs.path + ((size_t)b * (S + 1) + i) * 3;
            o[0] = X;
            o[1] = Y;
            o[2] = atan2(2.0 * (chh * sh), chh * chh - sh * sh);  // tf2 setRPY -> getYaw
          }
        }
      }
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(kThreads) smpc_eval_kernel(DevParams prm, DevBatch bt, const double* xin, DevEvalOut eo) {
  using L = Layout<NB>;
  constexpr int P = 2 * NB;
  __shared__ double smem[kWarpsPerCta * L::kTotal];
  const int lane = threadIdx.x & 31;
  double* ws = smem + (threadIdx.x >> 5) * L::kTotal;
  LaneConst<NB> lc;
  lane_setup<NB>(lane, prm.bl, bt.dt, lc);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < bt.B; b += n_warps) {
    Prob pb;
    load_problem(bt, b, pb);
    double aa_target[kMaxChunks];
    agent_angle_setup(prm, bt, pb, lane, aa_target);
    __syncwarp();
    if (lane < P) ws[L::kCand + lane] = __ldg(xin + (size_t)b * P + lane);
    __syncwarp();
    const unsigned fl = (bt.S > 32) ? evaluate<NB, true>(prm, bt, pb, aa_target, lc, ws + L::kCand, lane, ws + L::kBuf0)
                                    : evaluate<NB, false>(prm, bt, pb, aa_target, lc, ws + L::kCand, lane, ws + L::kBuf0);
    if (lane == 0) {
      if (eo.cost) eo.cost[b] = ws[L::kBuf0];
      if (eo.ok) eo.ok[b] = (fl == 0) ? 1 : 0;
    }
    if (eo.grad && lane < P) eo.grad[(size_t)b * P + lane] = ws[L::kBuf0 + 1 + lane];
    if (eo.hess)
      for (int e = lane; e < L::NH; e += 32) eo.hess[(size_t)b * L::NH + e] = ws[L::kBuf0 + 1 + P + e];
    __syncwarp();
  }
}

// Line-search polynomial minimiser exposed for unit tests: rows of (lo, hi, f0, g0, t1, f1, g1, t2, f2, g2);
// t2 <= 0 selects the two-sample (cubic) case.
__global__ void smpc_polymin_kernel(int n, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* r = in + (size_t)i * 10;
  out[i] = (r[7] > 0.0) ? quintic_interp_min(r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[0], r[1])
                        : cubic_interp_min(r[2], r[3], r[4], r[5], r[6], r[0], r[1]);
}

// One CTA per robot: arg-min of cost_final over its n_starts consecutive solves (usable ones only; ties -> lowest index).
__global__ void smpc_argmin_kernel(int n_starts, int n_blocks, const double* __restrict__ cost_final,
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
    if (best_u && lane < 2 * n_blocks) {
      best_u[(size_t)robot * 2 * n_blocks + lane] = (bi >= 0) ? u[(base + bi) * 2 * n_blocks + lane] : NAN;
    }
    if (best_u)
      for (int c = 32 + lane; c < 2 * n_blocks; c += 32)
        best_u[(size_t)robot * 2 * n_blocks + c] = (bi >= 0) ? u[(base + bi) * 2 * n_blocks + c] : NAN;
  }
}

// -------------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------------
template <int NB>
static cudaError_t launch_solve_nb(const DevParams& prm, const DevBatch& bt, const DevResult& rs, int* queue, int n_sm,
                                   cudaStream_t stream) {
  const bool multi = bt.S > 32;
  int ctas_per_sm = 0;
  cudaError_t e = multi ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, smpc_solve_kernel<NB, true>, kThreads, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, smpc_solve_kernel<NB, false>, kThreads, 0);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  // persistent grid: a multiple of the SM count, never more warps than problems
  long long want_ctas = ((long long)bt.B + kWarpsPerCta - 1) / kWarpsPerCta;
  long long grid = (long long)n_sm * ctas_per_sm;
  if (want_ctas < grid) grid = want_ctas;
  if (grid < 1) grid = 1;
  if (multi)
    smpc_solve_kernel<NB, true><<<(unsigned)grid, kThreads, 0, stream>>>(prm, bt, rs, queue);
  else
    smpc_solve_kernel<NB, false><<<(unsigned)grid, kThreads, 0, stream>>>(prm, bt, rs, queue);
  return cudaGetLastError();
}

template <int NB>
static cudaError_t launch_eval_nb(const DevParams& prm, const DevBatch& bt, const double* x, const DevEvalOut& eo, int n_sm,
                                  cudaStream_t stream) {
  long long want_ctas = ((long long)bt.B + kWarpsPerCta - 1) / kWarpsPerCta;
  long long grid = (long long)n_sm * 8;
  if (want_ctas < grid) grid = want_ctas;
  if (grid < 1) grid = 1;
  smpc_eval_kernel<NB><<<(unsigned)grid, kThreads, 0, stream>>>(prm, bt, x, eo);
  return cudaGetLastError();
}

#define SMPC_DISPATCH_NB(FN, ...)                  \
  switch (prm.nb) {                                \
    case 1: return FN<1>(__VA_ARGS__);             \
    case 2: return FN<2>(__VA_ARGS__);             \
    case 3: return FN<3>(__VA_ARGS__);             \
    case 4: return FN<4>(__VA_ARGS__);             \
    case 5: return FN<5>(__VA_ARGS__);             \
    case 6: return FN<6>(__VA_ARGS__);             \
    default: return cudaErrorInvalidValue;         \
  }

cudaError_t launch_solve(const DevParams& prm, const DevBatch& bt, const DevResult& rs, int* queue, int n_sm,
                         cudaStream_t stream) {
  SMPC_DISPATCH_NB(launch_solve_nb, prm, bt, rs, queue, n_sm, stream)
}

cudaError_t launch_eval(const DevParams& prm, const DevBatch& bt, const double* x, const DevEvalOut& eo, int n_sm,
                        cudaStream_t stream) {
  SMPC_DISPATCH_NB(launch_eval_nb, prm, bt, x, eo, n_sm, stream)
}

cudaError_t launch_argmin(int n_robots, int n_starts, int n_blocks, const double* cost_final, const uint8_t* usable,
                          const double* u, int32_t* best_index, double* best_cost, double* best_u, cudaStream_t stream) {
  int threads = 32;
  while (threads < n_starts && threads < 256) threads <<= 1;
  smpc_argmin_kernel<<<n_robots, threads, 0, stream>>>(n_starts, n_blocks, cost_final, usable, u, best_index, best_cost,
                                                       best_u);
  return cudaGetLastError();
}

int max_supported_blocks() { return 6; }

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

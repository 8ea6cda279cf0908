#!/usr/bin/env python
"""bench.py — MPC solves/sec of the batched social-MPC solver on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path (the whole bounded TR-LM solve of every problem) over one batch of
synthetic scenarios. Default workload = BASELINE.json configs[1]: obst_only parameters, 4096 corridor
scenarios, one costmap per problem. Under torchrun every rank solves its own batch (weak scaling, no
collective on the solve path); the time is the max over ranks.

  value      solves/s with the batch already resident in HBM (CUDA events on the launching stream)
  e2e        solves/s through the host-buffer C-ABI call smpc_solve_batch: pinned host inputs, H2D, kernel, D2H
  roofline   FP64 CUDA-core roofline of the solve kernel (the path is FP64-instruction bound, not HBM
             or tensor bound); the HBM side is reported next to it
  cpu_baseline  the CPU oracle (Ceres-algorithm restatement; Ceres itself is not installable here) on all host
             cores over a bounded sample of the same workload
--impl reference times only that CPU restatement (the reference's own code needs Ceres + ROS and cannot build here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nav2_social_mpc_controller_b200 import abi, scenarios as sc  # noqa: E402

METRIC = "mpc_solves_per_sec"
UNIT = "solves/s"

WORKLOADS = {
    # name: (builder, description, unique costmap per problem)
    "obst_only_x4096": (lambda: sc.corridor(B=4096), "BASELINE configs[1]: obst_only params, 4096 corridor scenarios, "
                        "S=28, P=6, one 80x80 costmap per problem", True),
    "soc_work_obst_x65536_A20": (lambda: sc.crowd(B=65536, A=20), "BASELINE configs[2]: soc_work_obst params, 65536 "
                                 "crowd scenarios, 20 agents, S=28, P=6, 256 shared costmaps", False),
    "soc_work_obst_x16384_A3": (lambda: sc.crowd(B=16384, A=3, config_id=6), "soc_work_obst params, 16384 crowd "
                                "scenarios with the reference's 3 agents, S=28, P=6", False),
    "obst_only_x65536": (lambda: sc.corridor(B=65536, unique_maps=False, config_id=22), "obst_only params, 65536 corridor "
                         "scenarios, 256 shared costmaps (throughput-mode check)", False),
    "soc_work_obst_x65536_A3": (lambda: sc.crowd(B=65536, A=3, config_id=23), "soc_work_obst params, 65536 crowd "
                                "scenarios with the reference's 3 agents", False),
    "multistart_256x1024": (lambda: sc.multistart(256, 1024), "BASELINE configs[3]: 1024 perturbed starts x 256 robots, "
                            "A=3, per-robot arg-min", False),
    "crowd_x16384_A50": (lambda: sc.crowd(B=16384, A=50, config_id=5), "BASELINE configs[4] slice: 16384 problems, "
                         "50 agents, S=28, P=6 (a 1.1 GB slice of one GPU's shard of the 10^6 sweep)", False),
}
# Workloads whose device batch is larger than the generated one: the unique scenarios are tiled on the device up to
# this many problems IN TOTAL over all ranks (strong scaling: each rank takes total / world). 10^6 problems at A = 50
# are 70 GB of agent trajectories: resident in one B200's HBM, never materialised on the host.
TILED_TOTAL = {"crowd_x1M_A50": 1_000_000}
WORKLOADS["crowd_x1M_A50"] = (lambda: sc.crowd(B=16384, A=50, config_id=5), "BASELINE configs[4] (unicycle): 10^6 "
                              "problems, 50 agents, S=28, P=6 = 16384 unique crowd scenarios tiled on the device "
                              "(70 GB of agent trajectories resident in HBM); e2e and CPU arms use the unique 16384",
                              False)
# bounded CPU samples: roughly 10-30 s of single-core oracle work each (a solve costs ~7 ms without people,
# ~15 ms at A = 3, ~30 ms at A = 20, ~75 ms at A = 50)
CPU_SAMPLE = {"obst_only_x4096": 3072, "obst_only_x65536": 3072, "soc_work_obst_x16384_A3": 1536,
              "soc_work_obst_x65536_A3": 1536, "multistart_256x1024": 1536, "soc_work_obst_x65536_A20": 512,
              "crowd_x16384_A50": 256, "crowd_x1M_A50": 256}


def flops_per_solve(S, P, A_eff, m, n_jac, n_cost, iters):
    """SURVEY §8d algorithmic FLOPs: F_solve = n_J F_jac + n_c F_cost + K F_lin. The per-agent term is 410 instead of
    SURVEY's a-priori 810: the social pair function is odd, so the minimal algorithm needs ONE pair interaction
    (~400 FLOPs with its 2x4 Jacobian) per agent and step, not two."""
    f_jac = S * (224 + 29 * P + 410 * A_eff) + 2 * m * (P * (P + 1) / 2 + P)
    f_cost = S * (120 + 170 * A_eff)
    f_lin = P ** 3 / 3 + 2 * P ** 2 + 4 * P
    return n_jac * f_jac + n_cost * f_cost + iters * f_lin


def n_residuals(batch):
    ch, bl, nb, nbd = batch.dims
    per_step = 8 if batch.arrays["has_people"].any() else 5
    return per_step * batch.n_steps + max(nbd - 1, 0)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe): one long-running
    `nvidia-smi -lms 50` process, every line time-stamped; summary() keeps the samples inside the marked windows."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.windows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append((time.perf_counter(), parts))
        except Exception:
            pass

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        inside = [p for (t, p) in self.samples if any(a <= t <= b for a, b in self.windows)]
        used = inside if inside else [p for _, p in self.samples]
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(x[0]) for x in used if x[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(x[3 + i].lower().startswith("active") for x in used)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(used[0][1]), "reasons": reasons,
                "samples_in_timed_regions": len(inside), "samples_total": len(self.samples)}


def load_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the solve kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/r01_traffic.json); None when no capture exists."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload)
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(batch, workload, threads, kind_note=True):
    from tests import oracle_lib
    o = oracle_lib.load()
    n = min(CPU_SAMPLE.get(workload, 256), batch.n_problems)
    sub = batch.slice(0, n)
    t0 = time.perf_counter()
    out = o.solve_batch(sub, n_threads=threads, want=("u", "cost_final", "usable", "iterations", "termination"))
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} problems of {workload}, {threads} host threads, {dt:.2f} s wall; "
                      "oracle = Ceres-algorithm restatement (Ceres/ROS not installable here)"}, out, sub


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm on the box's host cores (oracle port; rank 0 only)."""
    if rank != 0:
        return
    from tests import oracle_lib
    o = oracle_lib.load()
    builder, desc, _ = WORKLOADS[args.workload]
    batch = builder()
    threads = os.cpu_count() or 1
    n = min(CPU_SAMPLE.get(args.workload, 256), batch.n_problems)
    sub = batch.slice(0, n)
    want = ("u", "cost_final", "usable")
    for _ in range(args.warmup):
        o.solve_batch(sub.slice(0, min(n, 4 * threads)), n_threads=threads, want=want)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        o.solve_batch(sub, n_threads=threads, want=want)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * args.steps / total
    sample = (f"each step = first {n} problems of {args.workload} on {threads} host threads "
              "(CPU oracle: Ceres-algorithm restatement with per-residual Jet<4> re-rollouts)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "problems_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="obst_only_x4096", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-calls", type=int, default=1000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from nav2_social_mpc_controller_b200.optimizer import Optimizer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsmpc has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    builder, desc, unique_map = WORKLOADS[args.workload]
    batch = builder()
    if world > 1:  # weak scaling: every rank gets its own scenarios of the same shape
        rng = np.random.default_rng(1000 + rank)
        perm = rng.permutation(batch.n_problems)
        for k, v in batch.arrays.items():
            if v is not None and k not in ("costmaps", "costmap_origin"):
                batch.arrays[k] = np.ascontiguousarray(v[perm])
    B, S, A = batch.n_problems, batch.n_steps, batch.n_agents
    B_unique = B
    tiled = args.workload in TILED_TOTAL
    if tiled:
        B = TILED_TOTAL[args.workload] // world
    ch, bl, nb, nbd = batch.dims
    P = 2 * nb

    opt = Optimizer(local_rank)
    opt.initialize(batch.params)

    # ---- device-resident arm ------------------------------------------------------------------------
    host_pinned = {k: (torch.from_numpy(v).pin_memory() if v is not None else None) for k, v in batch.arrays.items()}
    dev_arrays = {k: (t.to(dev, non_blocking=True) if t is not None else None) for k, t in host_pinned.items()}
    if tiled:  # problem b of the device batch = unique scenario (rank offset + b) % B_unique; costmap = b % M as before
        assert batch.arrays.get("costmap_index") is not None or B_unique % batch.n_costmaps == 0
        from nav2_social_mpc_controller_b200.sharding import tiled_source_index
        src = torch.from_numpy(tiled_source_index(TILED_TOTAL[args.workload], B_unique, world, rank)).to(dev)
        for k, t in list(dev_arrays.items()):
            if t is None or k in ("costmaps", "costmap_origin"):
                continue
            big = torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            for lo in range(0, B, B_unique):  # gather piecewise: index_select on the whole batch would double the 70 GB
                n = min(B_unique, B - lo)
                torch.index_select(t, 0, src[lo:lo + n], out=big[lo:lo + n])
            dev_arrays[k] = big
        del src
    shapes = abi.result_shapes(B, S, nb)
    want = ("u", "cmds", "cost_initial", "cost_final", "iterations", "termination", "usable", "n_evals")
    tdt = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8}
    dev_out = {k: torch.zeros(shapes[k][0], dtype=tdt[shapes[k][1]], device=dev) for k in want}
    dstruct = abi.make_batch_struct(dev_arrays, B, S, A, batch.n_costmaps, batch.size_x, batch.size_y,
                                    batch.resolution, batch.dt)
    stream = torch.cuda.Stream(device=dev)  # non-default stream: the kernel and the events share it
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    def step_device():
        opt.solve_batch_device(dstruct, dev_out, stream=stream.cuda_stream)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    launches0 = opt.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_region0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            flush.fill_(s & 0xFF)  # L2 flush between timed iterations (outside the event pair)
            ev[s][0].record(stream)
            step_device()
            ev[s][1].record(stream)
    torch.cuda.synchronize()
    sampler.mark(t_region0, time.perf_counter())
    launches = opt.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    kernel_ms = opt.last_kernel_ms()

    # ---- end-to-end arm: host-buffer C-ABI call, pinned host memory, H2D + kernel + D2H every step ---------
    host_np = {k: (t.numpy() if t is not None else None) for k, t in host_pinned.items()}
    hshapes = abi.result_shapes(B_unique, S, nb)
    host_out_t = {k: torch.zeros(hshapes[k][0], dtype=tdt[hshapes[k][1]]).pin_memory() for k in want}
    host_out = {k: t.numpy() for k, t in host_out_t.items()}
    hbatch = sc.Batch(params=batch.params, n_problems=B_unique, n_steps=S, n_agents=A, n_costmaps=batch.n_costmaps,
                      size_x=batch.size_x, size_y=batch.size_y, resolution=batch.resolution, dt=batch.dt, arrays=host_np)
    for _ in range(args.warmup):
        opt.solve_batch(hbatch, out=host_out)
    if world > 1:
        dist.barrier()
    e2e_t = []
    t_region0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.solve_batch(hbatch, out=host_out)
        e2e_t.append(time.perf_counter() - t0)
    e2e_total = float(sum(e2e_t))
    sampler.mark(t_region0, time.perf_counter())
    time.sleep(0.06)
    sampler.stop()
    sampler.join(timeout=2)
    h2d = int(sum(v.nbytes for v in host_np.values() if v is not None))
    d2h = int(sum(v.nbytes for v in host_out.values()))

    # parity spot check of the timed run against the e2e run (same inputs -> identical results)
    dev_u = dev_out["u"][:B_unique].cpu().numpy()
    if tiled and (rank * B) % B_unique:
        dev_u = None  # this rank's device batch starts at another unique scenario
    if dev_u is not None and not np.array_equal(dev_u, host_out["u"]):
        raise SystemExit("device-resident and host-buffer solves disagree")

    # ---- reduce over ranks (max time), whole-job value ---------------------------------------------------
    if world > 1:
        t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0]), float(t[1])
    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_value = world * B_unique * args.steps / e2e_total

    line = None
    if rank == 0:
        n_evals = dev_out["n_evals"].cpu().numpy().astype(np.float64)
        iters = dev_out["iterations"].cpu().numpy().astype(np.float64)
        term = dev_out["termination"].cpu().numpy()
        A_eff = A if batch.arrays["has_people"].any() else 0
        m = n_residuals(batch)
        flops = float(sum(flops_per_solve(S, P, A_eff, m, n_evals[:, 0], n_evals[:, 1], iters)))
        peaks, peak_kind = load_peaks()
        fp64_peak = opt.measure_fp64_peak()
        k_ms = total_ms / args.steps if world > 1 else float(np.mean(step_ms))
        achieved_tflops = flops / (k_ms * 1e-3) / 1e12
        alg_bytes = B * batch.algorithmic_bytes_per_problem(unique_map)
        if not unique_map:
            alg_bytes += batch.n_costmaps * batch.size_x * batch.size_y
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if tiled else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "problems_per_gpu": B, "n_steps": S,
                       "e2e_problems_per_gpu": B_unique,
                       "n_params": P, "n_agents": A, "l2": "flushed between timed steps (256 MiB write)",
                       "timing": "CUDA events per step on the launching stream, max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_total / args.steps,
                    "how": "smpc_solve_batch(host pinned buffers): H2D + solve kernel + D2H + sync, wall clock"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp64_peak if fp64_peak > 0 else None,
                         "traffic": (load_traffic(args.workload) or {}).get("bytes"),
                         "traffic_detail": load_traffic(args.workload),
                         "kernel": f"smpc_solve_kernel<{nb}>", "kernel_ms": k_ms,
                         "peak_source": "DFMA microbenchmark measured in this run (smpc_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "algorithmic_flops_per_launch": flops,
                         "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                 "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_kind}},
            "solver": {"mean_iterations": float(iters.mean()), "mean_evaluations": float(n_evals[:, 0].mean()),
                       "termination_histogram": {abi.TERMINATION_NAMES[int(k)]: int(v) for k, v in
                                                 zip(*np.unique(term, return_counts=True))},
                       "usable_fraction": float(dev_out["usable"].float().mean())},
        }
        # p50 / p99 single-solve latency through the host-buffer C-ABI (B = 1, incl. H2D / D2H)
        if args.latency_calls > 0:
            one = hbatch.slice(0, 1)
            one_out = {k: np.zeros((1,) + shapes[k][0][1:], dtype=shapes[k][1]) for k in want}
            for _ in range(20):
                opt.solve_batch(one, out=one_out)
            lat = []
            for _ in range(args.latency_calls):
                t0 = time.perf_counter()
                opt.solve_batch(one, out=one_out)
                lat.append(time.perf_counter() - t0)
            lat = np.sort(np.array(lat)) * 1e3
            line["latency_ms"] = {"p50": float(lat[len(lat) // 2]), "p99": float(lat[int(len(lat) * 0.99)]),
                                  "calls": args.latency_calls, "what": "smpc_solve_batch, B=1, host buffers"}
        if not args.no_cpu_baseline and world == 1:
            cb, ref_out, sub = cpu_baseline(hbatch, args.workload, os.cpu_count() or 1)
            n = sub.n_problems
            us = ref_out["usable"].astype(bool)
            du = np.abs(host_out["u"][:n] - ref_out["u"]).reshape(n, -1).max(axis=1)
            dc = np.abs(host_out["cost_final"][:n] - ref_out["cost_final"]) / np.maximum(np.abs(ref_out["cost_final"]), 1e-300)
            ok = (~us & (host_out["usable"][:n] == 0)) | (us & (du <= 1e-6) & (dc <= 1e-8))
            line["cpu_baseline"] = cb
            line["parity_vs_oracle"] = {"problems": int(n), "within_1e-6_u_and_1e-8_cost": float(ok.mean()),
                                        "same_termination": float((host_out["termination"][:n] == ref_out["termination"]).mean()),
                                        "same_iteration_count": float((host_out["iterations"][:n] == ref_out["iterations"]).mean()),
                                        "max_du": float(du[us].max()) if us.any() else None,
                                        "max_rel_dcost": float(dc[us].max()) if us.any() else None}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    opt.close()


if __name__ == "__main__":
    main()

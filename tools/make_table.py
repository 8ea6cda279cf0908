#!/usr/bin/env python
"""Markdown rows of BASELINE.md §4 from the bench lines of one evidence run. usage: make_table.py gpurun_out/finalN"""
import glob
import json
import os
import sys

src = sys.argv[1]
order = ["obst_only_x4096", "obst_only_x65536", "soc_work_obst_x16384_A3", "soc_work_obst_x65536_A3",
         "soc_work_obst_x65536_A20", "multistart_256x1024", "crowd_x16384_A50", "crowd_x1M_A50"]
rows = {}
for f in glob.glob(os.path.join(src, "bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    rows[d["config"]["workload"]] = d
print("| workload (`bench.py --workload`) | GPU solves/s (resident) | ms / batch | e2e solves/s | CPU restatement solves/s "
      "(threads) | e2e ÷ CPU | FP64 frac | evaluations (with JᵀJ + gradient-only) / iterations | parity (prefix) |")
print("|---|---|---|---|---|---|---|---|---|")
for w in order:
    if w not in rows:
        continue
    d = rows[w]
    cb, pv = d.get("cpu_baseline", {}), d.get("parity_vs_oracle", {})
    n = pv.get("problems", 0)
    ok = round(pv.get("within_1e-6_u_and_1e-8_cost", 0) * n)
    sv = d["solver"]
    print(f"| `{w}` | {d['value'] / 1e6:.3f} M | {d['ms_per_step']:.2f} | {d['e2e']['value'] / 1e6:.3f} M | "
          f"{cb.get('value', 0) / 1e3:.2f} k ({cb.get('cores')}) | {d['e2e']['value'] / max(cb.get('value', 1), 1):.0f}× | "
          f"{d['roofline']['frac']:.3f} | {sv['mean_evaluations_with_JtJ']:.1f} + {sv['mean_evaluations_gradient_only']:.1f} / "
          f"{sv['mean_iterations']:.1f} | {ok}/{n} |")

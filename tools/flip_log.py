#!/usr/bin/env python
"""Per-problem flip log: where and why a GPU solve leaves the oracle's iterate path.

Both solvers record one row per trial point (smpc_result.trace / smpc_oracle_solve_evals: iteration, phase, step size,
differentiated cost, plain cost, aux, decision code, radius). For every problem of a workload prefix whose result is
outside the north-star tolerance (controls 1e-6, final cost 1e-8) or whose iteration count / termination differs, this
tool aligns the two traces, finds the FIRST row that differs and classifies the divergence:

  armijo      the same line-search sample passes the sufficient-decrease test on one side only
              margin = |cost - (cost(x) + 1e-4 g0 t)| / cost(x)
  accept      the same candidate is accepted (relative decrease > 1e-3) on one side only;  margin = |rho - 1e-3|
  param_tol / fn_tol / gradient / iteration-cap
              one side terminates on this row, the other goes on;  margin = the terminating side's distance to its
              threshold (aux of that row), relative
  step        same decisions so far, but the next step size differs by more than 1e-6 relative (the line-search
              polynomial minimiser picked another root / end point);  margin = relative difference of the two steps
  value       same point (step size equal to 1e-9), costs differ by more than 1e-9 relative: a branch inside a residual
              flipped (sgn(theta) of the social force, closest-agent choice) or round-off was amplified earlier

Run on the GPU box:  python tools/flip_log.py --workload crowd_A20 --n 512 --out profiles/r02_flip_log_A20.json
Needs the CPU oracle (test infrastructure) — this is a diagnostic, not a product path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nav2_social_mpc_controller_b200 import scenarios as sc  # noqa: E402
from nav2_social_mpc_controller_b200.optimizer import Optimizer  # noqa: E402
from tests import oracle_lib  # noqa: E402

WORKLOADS = {
    "corridor": lambda n, kw: sc.corridor(B=n, **kw),
    "crowd_A3": lambda n, kw: sc.crowd(B=n, A=3, config_id=6, **kw),
    "crowd_A20": lambda n, kw: sc.crowd(B=n, A=20, **kw),
    "crowd_A50": lambda n, kw: sc.crowd(B=n, A=50, config_id=5, **kw),
    "blocks18": lambda n, kw: sc.crowd(B=n, A=3, config_id=31, control_horizon=18, parameter_block_length=1, **kw),
}
TERM = {0: "gradient", 1: "param_tol", 2: "fn_tol", 3: "radius", 4: "iteration-cap", 5: "invalid-steps", 6: "eval-failure"}


def oracle_rows(oracle, batch, b, max_rows=400):
    rows = np.zeros((max_rows, 8))
    st = batch.struct()
    f = oracle.lib.smpc_oracle_solve_evals
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    f.restype = C.c_int
    n = f(C.byref(batch.params), C.byref(st), b, rows.ctypes.data, max_rows)
    return rows[:n]


def rel(a, b):
    return abs(a - b) / max(abs(a), abs(b), 1e-300)


def classify(g, o):
    """g, o: trace rows [n][8] of the GPU and of the oracle. Returns dict(row, kind, margin, ...)."""
    n = min(len(g), len(o))
    x_cost = None
    for k in range(n):
        gr, orow = g[k], o[k]
        gc, oc = int(gr[6]), int(orow[6])
        same_point = gr[1] == orow[1] and rel(gr[2], orow[2]) <= 1e-6
        info = dict(row=k, iteration=int(orow[0]), phase=int(orow[1]), t_gpu=gr[2], t_oracle=orow[2],
                    cost_gpu=gr[3], cost_oracle=orow[3], aux_gpu=gr[5], aux_oracle=orow[5], code_gpu=gc, code_oracle=oc)
        if gr[1] != orow[1]:
            return dict(info, kind="phase", margin=None)
        if not same_point:
            return dict(info, kind="step", margin=rel(gr[2], orow[2]))
        both = not (np.isnan(gr[3]) or np.isnan(orow[3]))
        if both and rel(gr[3], orow[3]) > 1e-9 and gc == oc:
            return dict(info, kind="value", margin=rel(gr[3], orow[3]))
        if gc != oc:
            diff = gc ^ oc
            if diff & 1:  # Armijo pass on one side only: the rejecting side recorded cost - rhs
                m = gr[5] if not (gc & 1) else orow[5]
                return dict(info, kind="armijo", margin=abs(m) / max(abs(x_cost or orow[3]), 1e-300))
            if diff & 8 or (gc >> 4) != (oc >> 4):  # one side terminated here (or for another reason)
                tg, to = (gc >> 4) if gc & 8 else None, (oc >> 4) if oc & 8 else None
                side = gr if gc & 8 else orow
                t = tg if gc & 8 else to
                m = abs(side[5]) if t in (1, 2) and not np.isnan(side[5]) else None
                if t == 2 and m is not None:
                    m = m / max(abs(x_cost or 1.0), 1e-300)
                return dict(info, kind=TERM.get(t, "termination"), margin=m, term_gpu=tg, term_oracle=to)
            if diff & 4:
                rho = gr[5] if not np.isnan(gr[5]) else orow[5]
                return dict(info, kind="accept", margin=abs(rho - 1e-3))
            return dict(info, kind="code", margin=None)
        if gc & 4 or int(gr[1]) == 1:  # accepted (or iteration zero): this row's differentiated cost is the new cost(x)
            x_cost = orow[3]
    if len(g) != len(o):
        return dict(row=n, kind="length", margin=None, rows_gpu=len(g), rows_oracle=len(o))
    return dict(row=-1, kind="none", margin=None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="crowd_A20", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--ceres-compat", type=int, default=200)
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    batch = WORKLOADS[args.workload](args.n, dict(ceres_compat=args.ceres_compat))
    oracle = oracle_lib.load()
    opt = Optimizer(0)
    opt.initialize(batch.params)
    if args.group:
        opt.set_group(args.group)
    got = opt.solve_batch(batch, trace_rows=400)
    opt.close()
    ref = oracle.solve_batch(batch, n_threads=os.cpu_count() or 1)
    n = batch.n_problems
    us = ref["usable"].astype(bool)
    du = np.abs(got["u"] - ref["u"]).reshape(n, -1).max(axis=1)
    dc = np.abs(got["cost_final"] - ref["cost_final"]) / np.maximum(np.abs(ref["cost_final"]), 1e-300)
    ok = (~us & (got["usable"] == 0)) | (us & (got["usable"] == 1) & (du <= 1e-6) & (dc <= 1e-8))
    same_path = (got["iterations"] == ref["iterations"]) & (got["termination"] == ref["termination"])
    suspects = np.nonzero(~ok | ~same_path)[0]
    entries = []
    for b in suspects:
        g = got["trace"][b]
        g = g[~np.isnan(g[:, 0])]
        o = oracle_rows(oracle, batch, int(b))
        c = classify(g, o)
        c.update(problem=int(b), within_tolerance=bool(ok[b]), du=float(du[b]), dcost=float(dc[b]),
                 iterations_gpu=int(got["iterations"][b]), iterations_oracle=int(ref["iterations"][b]),
                 termination_gpu=int(got["termination"][b]), termination_oracle=int(ref["termination"][b]))
        entries.append({k: (None if isinstance(v, float) and np.isnan(v) else v) for k, v in c.items()})
    kinds = {}
    for e in entries:
        kinds.setdefault(e["kind"], []).append(e["margin"])
    summary = dict(workload=args.workload, problems=n, ceres_compat=args.ceres_compat,
                   within_tolerance=int(ok.sum()), same_iterations_and_termination=int(same_path.sum()),
                   out_of_tolerance=int((~ok).sum()), max_du=float(du[us].max()) if us.any() else None,
                   max_rel_dcost=float(dc[us].max()) if us.any() else None,
                   divergences={k: dict(count=len(v), max_margin=max([m for m in v if m is not None], default=None))
                                for k, v in kinds.items()})
    doc = dict(summary=summary, entries=entries)
    text = json.dumps(doc, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    print(json.dumps(summary))
    for e in entries[:40]:
        print(f"  problem {e['problem']:5d} row {e['row']:3d} it {e.get('iteration')} {e['kind']:13s} margin {e['margin']} "
              f"du {e['du']:.2e} dcost {e['dcost']:.2e} iters {e['iterations_gpu']}/{e['iterations_oracle']}")


if __name__ == "__main__":
    main()

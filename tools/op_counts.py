#!/usr/bin/env python
"""FLOPs of ONE differentiated evaluation (cost + Jacobian): what the reference-shaped algorithm executes in its Jet
arithmetic (oracle built with operation counting: per-residual Jet<4> passes, every functor re-rolling-out steps 0..i,
SURVEY §8d convention) next to the minimal single-rollout analytic model that bench.py's roofline uses. CPU only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nav2_social_mpc_controller_b200 import abi, scenarios as sc  # noqa: E402


def load():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "liboracle_count.so"], check=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_count.so"))
    lib.smpc_oracle_count_jet_flops.restype = C.c_longlong
    lib.smpc_oracle_count_jet_flops.argtypes = [C.POINTER(abi.SmpcParams), C.POINTER(abi.SmpcBatch), C.c_int, C.c_void_p]
    return lib


def count(lib, batch, b=0):
    st = batch.struct()
    x = np.ascontiguousarray(batch.arrays["u0"][b], dtype=np.float64).ravel()
    return int(lib.smpc_oracle_count_jet_flops(C.byref(batch.params), C.byref(st), b, x.ctypes.data))


def model_f_jac(S, P, A_eff, m):
    """bench.py flops_per_solve's F_jac (DESIGN.md §3 Roofline)."""
    return S * (224 + 29 * P + 410 * A_eff) + 2 * m * (P * (P + 1) / 2 + P)


CASES = [
    ("README example, 3 agents", lambda: sc.single("readme", n_people=3)),
    ("obst_only (people-free)", lambda: sc.corridor(B=1)),
    ("soc_work_obst, 3 agents", lambda: sc.single("soc_work_obst", n_people=3)),
    ("soc_work_obst, 20 agents", lambda: sc.crowd(B=1, A=20, n_maps=1)),
    ("soc_work_obst, 50 agents", lambda: sc.crowd(B=1, A=50, n_maps=1, config_id=5)),
    ("params/params.yaml, 3 agents", lambda: sc.single("params_yaml", n_people=3)),
]

if __name__ == "__main__":
    lib = load()
    print("| problem | S / P / A | reference-shaped Jet FLOPs per evaluation | minimal analytic model (F_jac) | ratio |")
    print("|---|---|---|---|---|")
    for name, make in CASES:
        b = make()
        S, P, A = b.n_steps, 2 * b.n_blocks, b.n_agents
        people = bool(b.arrays["has_people"].any())
        ch, bl, nb, nbd = b.dims
        m = (8 if people else 5) * S + max(nbd - 1, 0)
        ref = count(lib, b)
        mod = model_f_jac(S, P, A if people else 0, m)
        print(f"| {name} | {S} / {P} / {A if people else 0} | {ref / 1e3:.0f} k | {mod / 1e3:.1f} k | {ref / mod:.1f}× |")

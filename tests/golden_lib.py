"""Loader of the committed golden vectors (tests/golden/solve_cases.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np

from nav2_social_mpc_controller_b200 import abi, scenarios as sc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_cases.npz")


def load():
    """{case name: (Batch, outputs dict, eval dict)}"""
    z = np.load(PATH)
    names = sorted({k.split("/")[0] for k in z.files})
    cases = {}
    for n in names:
        raw = z[f"{n}/params"].tobytes()
        assert len(raw) == abi.C.sizeof(abi.SmpcParams), "smpc_params layout changed: regenerate tests/golden"
        params = abi.SmpcParams.from_buffer_copy(raw)
        B, S, A, M, sx, sy = (int(v) for v in z[f"{n}/meta"])
        res, dt = (float(v) for v in z[f"{n}/scalars"])
        arrays = {k: None for k in abi.BATCH_FIELDS}
        for k in z.files:
            if k.startswith(f"{n}/in/"):
                arrays[k.split("/")[2]] = np.ascontiguousarray(z[k])
        batch = sc.Batch(params=params, n_problems=B, n_steps=S, n_agents=A, n_costmaps=M, size_x=sx, size_y=sy,
                         resolution=res, dt=dt, arrays=arrays)
        outs = {k.split("/")[2]: z[k] for k in z.files if k.startswith(f"{n}/out/")}
        ev = {k.split("/")[2]: z[k] for k in z.files if k.startswith(f"{n}/eval/")}
        cases[n] = (batch, outs, ev)
    return cases

"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from nav2_social_mpc_controller_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        P = C.POINTER
        dp, ip = P(C.c_double), P(C.c_int)
        lib.smpc_oracle_num_residuals.argtypes = [P(abi.SmpcParams), P(abi.SmpcBatch), C.c_int]
        lib.smpc_oracle_num_residuals.restype = C.c_int
        lib.smpc_oracle_evaluate.argtypes = [P(abi.SmpcParams), P(abi.SmpcBatch), C.c_int, dp, dp, dp, dp, dp]
        lib.smpc_oracle_evaluate.restype = C.c_int
        lib.smpc_oracle_layout.argtypes = [P(abi.SmpcParams), P(abi.SmpcBatch), C.c_int, ip, ip]
        lib.smpc_oracle_layout.restype = C.c_int
        lib.smpc_oracle_solve_batch.argtypes = [P(abi.SmpcParams), P(abi.SmpcBatch), P(abi.SmpcResult), C.c_int,
                                                C.c_int, C.c_int]
        lib.smpc_oracle_solve_batch.restype = C.c_int
        lib.smpc_oracle_solve_trace.argtypes = [P(abi.SmpcParams), P(abi.SmpcBatch), C.c_int, dp, dp, C.c_int, ip, dp,
                                                dp]
        lib.smpc_oracle_solve_trace.restype = C.c_int
        lib.smpc_oracle_poly_min.argtypes = [dp, C.c_int, C.c_double, C.c_double]
        lib.smpc_oracle_poly_min.restype = C.c_double
        lib.smpc_oracle_poly_roots.argtypes = [dp, C.c_int, dp]
        lib.smpc_oracle_poly_roots.restype = C.c_int
        lib.smpc_oracle_bicubic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp]
        lib.smpc_oracle_bicubic.restype = C.c_double
        lib.smpc_oracle_yaw_roundtrip.argtypes = [C.c_double]
        lib.smpc_oracle_yaw_roundtrip.restype = C.c_double

    @staticmethod
    def _dp(a):
        return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))

    def num_residuals(self, batch, b=0):
        batch = batch.expanded()  # scenario-sharing batches: the oracle reads one row per problem
        st = batch.struct()
        return self.lib.smpc_oracle_num_residuals(C.byref(batch.params), C.byref(st), b)

    def layout(self, batch, b=0):
        m = self.num_residuals(batch, b)
        kinds = np.zeros(m, dtype=np.int32)
        steps = np.zeros(m, dtype=np.int32)
        batch = batch.expanded()  # scenario-sharing batches: the oracle reads one row per problem
        st = batch.struct()
        ip = C.POINTER(C.c_int)
        self.lib.smpc_oracle_layout(C.byref(batch.params), C.byref(st), b, kinds.ctypes.data_as(ip),
                                    steps.ctypes.data_as(ip))
        return kinds, steps

    def evaluate(self, batch, b, x, want_jac=True):
        """Returns dict(ok, cost, residuals[m], grad[P], jac[m][P])."""
        m = self.num_residuals(batch, b)
        each = batch.arrays.get("n_steps_each")
        S_b = batch.n_steps if each is None else int(each[b])  # problem b's own horizon -> its own block count
        dof = 3 if int(batch.params.omni_solve) else 2
        P = dof * abi.problem_dims(batch.params.control_horizon, batch.params.parameter_block_length, S_b)[2]
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel()[:P])
        cost = C.c_double(0.0)
        res = np.zeros(m)
        grad = np.zeros(P) if want_jac else None
        jac = np.zeros((m, P)) if want_jac else None
        batch = batch.expanded()  # scenario-sharing batches: the oracle reads one row per problem
        st = batch.struct()
        ok = self.lib.smpc_oracle_evaluate(C.byref(batch.params), C.byref(st), b, self._dp(x), C.byref(cost),
                                           self._dp(res), self._dp(grad), self._dp(jac))
        return dict(ok=ok == 1, cost=cost.value, residuals=res, grad=grad, jac=jac)

    def solve_batch(self, batch, first=0, count=None, n_threads=1, want=("u", "cmds", "path", "cost_initial",
                                                                          "cost_final", "iterations",
                                                                          "termination", "usable", "n_evals")):
        count = batch.n_problems - first if count is None else count
        shapes = abi.result_shapes(batch.n_problems, batch.n_steps, batch.n_blocks, 3 if int(batch.params.omni_solve) else 2)
        out = {k: np.zeros(shapes[k][0], dtype=shapes[k][1]) for k in want}
        rs = abi.make_result_struct(out)
        batch = batch.expanded()  # scenario-sharing batches: the oracle reads one row per problem
        st = batch.struct()
        rc = self.lib.smpc_oracle_solve_batch(C.byref(batch.params), C.byref(st), C.byref(rs), first, count,
                                              n_threads)
        if rc != 0:
            raise RuntimeError(f"smpc_oracle_solve_batch failed: {rc}")
        return out

    def solve_trace(self, batch, b=0, max_rows=256):
        P = (3 if int(batch.params.omni_solve) else 2) * batch.n_blocks
        x = np.zeros(P)
        trace = np.zeros((max_rows, 10))
        term = C.c_int(0)
        ci, cf = C.c_double(0), C.c_double(0)
        batch = batch.expanded()  # scenario-sharing batches: the oracle reads one row per problem
        st = batch.struct()
        rows = self.lib.smpc_oracle_solve_trace(C.byref(batch.params), C.byref(st), b, self._dp(x), self._dp(trace),
                                                max_rows, C.byref(term), C.byref(ci), C.byref(cf))
        return dict(x=x, trace=trace[:rows], termination=term.value, cost_initial=ci.value, cost_final=cf.value)

    def poly_min(self, samples, x_min, x_max):
        s = np.ascontiguousarray(samples, dtype=np.float64)
        return self.lib.smpc_oracle_poly_min(self._dp(s), s.shape[0], x_min, x_max)

    def poly_roots(self, coeffs):
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        out = np.zeros(len(c) + 2)
        n = self.lib.smpc_oracle_poly_roots(self._dp(c), len(c), self._dp(out))
        return out[:n]

    def bicubic(self, cmap, r, c):
        cmap = np.ascontiguousarray(cmap, dtype=np.uint8)
        dr, dc = C.c_double(0), C.c_double(0)
        f = self.lib.smpc_oracle_bicubic(cmap.ctypes.data, cmap.shape[1], cmap.shape[0], r, c, C.byref(dr),
                                         C.byref(dc))
        return f, dr.value, dc.value


def load(rebuild_if_missing=True) -> Oracle:
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if rebuild_if_missing:
            build()
        _LIB = Oracle(C.CDLL(path))
    return _LIB

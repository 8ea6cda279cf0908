"""Host-side mirror of the reference optimizer interface for the solve path, over libsmpc.so.

Reference interface mirrored (include/nav2_social_mpc_controller/optimizer.hpp):
  struct OptimizerParams (:59-101)          -> OptimizerParams (defaults of src/optimizer.cpp:26-84, from_yaml)
  Optimizer::initialize(params) (:152)      -> Optimizer.initialize(params)
  ceres::Solve inside Optimizer::optimize   -> Optimizer.solve_batch / solve_batch_device (level-1 C-ABI)
All compute happens in the CUDA library; this file only marshals buffers. No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, abi


class OptimizerParams:
    """Same fields and defaults as the reference OptimizerParams (+ trajectorizer.* the optimizer reads)."""

    def __init__(self, **overrides):
        self.c = abi.SmpcParams()
        _lib.lib().smpc_params_default(C.byref(self.c))
        for k, v in overrides.items():
            self.set(k, v)

    def set(self, key, value):
        if key in ("linear_solver_type", "base_frame") and isinstance(value, str):
            value = value.encode()
        setattr(self.c, key, value)

    def __getattr__(self, key):
        if key == "c":
            raise AttributeError(key)
        v = getattr(self.c, key)
        return v.decode() if isinstance(v, bytes) else v

    @classmethod
    def from_yaml(cls, path: str, plugin_name: str = "FollowPath") -> "OptimizerParams":
        """OptimizerParams::get (reference src/optimizer.cpp:16-85): reads `<plugin>.optimizer.*`,
        `<plugin>.optimizer.weights.*` and `<plugin>.trajectorizer.*`. Raises on an invalid linear_solver_type
        like the reference's std::runtime_error (:44)."""
        self = cls.__new__(cls)
        self.c = abi.SmpcParams()
        _lib.check(_lib.lib().smpc_params_from_yaml(path.encode(), plugin_name.encode(), C.byref(self.c)))
        return self

    @classmethod
    def from_struct(cls, c: abi.SmpcParams) -> "OptimizerParams":
        self = cls.__new__(cls)
        self.c = c
        return self


class Optimizer:
    """Batched drop-in for the solve inside nav2_social_mpc_controller::Optimizer. One instance per GPU."""

    def __init__(self, device: int = 0):
        self.device = device
        self._h = None
        self.params = None

    def initialize(self, params) -> None:
        """Optimizer::initialize (reference src/optimizer.cpp:98-132)."""
        if isinstance(params, abi.SmpcParams):
            params = OptimizerParams.from_struct(params)
        self.close()
        h = C.c_void_p()
        _lib.check(_lib.lib().smpc_create(C.byref(params.c), self.device, C.byref(h)))
        self._h = h
        self.params = params

    def close(self):
        if self._h is not None:
            _lib.lib().smpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _need(self):
        if self._h is None:
            raise RuntimeError("Optimizer.initialize(params) has not been called")

    def dims(self, n_steps: int):
        ch, bl, nb, nbd = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().smpc_problem_dims(C.byref(self.params.c), n_steps, C.byref(ch), C.byref(bl), C.byref(nb),
                                                C.byref(nbd)))
        return ch.value, bl.value, nb.value, nbd.value

    # ---- level-1 solve, host buffers (numpy or pinned torch CPU tensors) --------------------------------
    def solve_batch(self, batch, out: dict | None = None,
                    want=("u", "cmds", "cost_initial", "cost_final", "iterations", "termination", "usable",
                          "n_evals")) -> dict:
        """H2D copy of the batch, one solve launch, D2H copy of the results (all inside the C call)."""
        self._need()
        nb = self.dims(batch.n_steps)[2]
        if out is None:
            shapes = abi.result_shapes(batch.n_problems, batch.n_steps, nb)
            out = {k: np.zeros(shapes[k][0], dtype=shapes[k][1]) for k in want}
        st = batch.struct()
        rs = abi.make_result_struct(out)
        _lib.check(_lib.lib().smpc_solve_batch(self._h, C.byref(st), C.byref(rs)))
        return out

    # ---- level-1 solve, device buffers (torch CUDA tensors), asynchronous on `stream` -------------------
    def solve_batch_device(self, batch_struct: abi.SmpcBatch, out: dict, stream: int | None = None) -> None:
        self._need()
        rs = abi.make_result_struct(out)
        _lib.check(_lib.lib().smpc_solve_batch_device(self._h, C.byref(batch_struct), C.byref(rs),
                                                      C.c_void_p(stream) if stream else None))

    def eval_batch(self, batch, x: np.ndarray) -> dict:
        """cost, J^T r, J^T J at block values x [B][NB][2] (host buffers)."""
        self._need()
        nb = self.dims(batch.n_steps)[2]
        P = 2 * nb
        B = batch.n_problems
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(B, P)
        out = dict(cost=np.zeros(B), grad=np.zeros((B, P)), hess=np.zeros((B, P * (P + 1) // 2)),
                   ok=np.zeros(B, dtype=np.uint8))
        eo = abi.SmpcEvalOut()
        for k in ("cost", "grad", "hess", "ok"):
            setattr(eo, k, out[k].ctypes.data)
        st = batch.struct()
        _lib.check(_lib.lib().smpc_eval_batch(self._h, C.byref(st), x.ctypes.data, C.byref(eo)))
        return out

    def multistart_argmin_device(self, n_robots, n_starts, n_blocks, cost_final, usable, u, best_index, best_cost,
                                 best_u, stream: int | None = None):
        self._need()
        p = abi._ptr
        _lib.check(_lib.lib().smpc_multistart_argmin_device(
            self._h, n_robots, n_starts, n_blocks, p(cost_final), p(usable), p(u), p(best_index), p(best_cost),
            p(best_u), C.c_void_p(stream) if stream else None))

    def set_group(self, lanes_per_problem: int) -> None:
        """Lanes per problem (0 = auto, 4, 8, 16, 32): throughput vs latency mapping of the solve kernel."""
        self._need()
        _lib.check(_lib.lib().smpc_set_group(self._h, lanes_per_problem))

    def last_kernel_ms(self) -> float:
        self._need()
        return float(_lib.lib().smpc_last_kernel_ms(self._h))

    def debug_polymin(self, rows: np.ndarray) -> np.ndarray:
        """Line-search polynomial minimiser on rows of (lo, hi, f0, g0, t1, f1, g1, t2, f2, g2) (unit tests)."""
        self._need()
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 10)
        out = np.zeros(rows.shape[0])
        _lib.check(_lib.lib().smpc_debug_polymin(self._h, rows.shape[0], rows.ctypes.data, out.ctypes.data))
        return out

    def measure_fp64_peak(self) -> float:
        """TFLOP/s of a DFMA-saturating microbenchmark on this GPU (roofline denominator)."""
        self._need()
        v = C.c_double(0.0)
        _lib.check(_lib.lib().smpc_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        self._need()
        return int(_lib.lib().smpc_launch_count(self._h))


def hess_to_dense(h_packed: np.ndarray, P: int) -> np.ndarray:
    """Row-major lower-triangle packing of include/smpc.h -> dense symmetric [.., P, P]."""
    out = np.zeros(h_packed.shape[:-1] + (P, P))
    e = 0
    for a in range(P):
        for b in range(a + 1):
            out[..., a, b] = h_packed[..., e]
            out[..., b, a] = h_packed[..., e]
            e += 1
    return out

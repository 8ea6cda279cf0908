"""Operation counting in the oracle's Jet algebra: the reference-shaped evaluator re-rolls-out steps
0..i inside every functor, so its FLOPs per evaluation grow quadratically with the horizon; the regular oracle library
carries no counting code."""
import ctypes as C
import os

from nav2_social_mpc_controller_b200 import abi, scenarios as sc
from tools import op_counts


def test_reference_shaped_flops_grow_quadratically_with_the_horizon():
    lib = op_counts.load()
    short = sc.single("readme", n_people=3)          # S = 13
    long_ = sc.single("soc_work_obst", n_people=3)   # S = 28, same P = 6 and agent count
    n_short, n_long = op_counts.count(lib, short), op_counts.count(lib, long_)
    assert short.n_steps == 13 and long_.n_steps == 28 and n_short > 0
    ratio = n_long / n_short
    assert (28 / 13) ** 1.7 < ratio < (28 / 13) ** 2.2, ratio            # O(S^2) rollout work dominates
    # against the minimal single-rollout model (linear in S) the reference shape costs an order of magnitude more
    m = 8 * 28 + 2
    assert n_long > 5 * op_counts.model_f_jac(28, 6, 3, m)


def test_regular_oracle_has_no_counting(oracle):
    fn = oracle.lib.smpc_oracle_count_jet_flops
    fn.restype = C.c_longlong
    fn.argtypes = [C.POINTER(abi.SmpcParams), C.POINTER(abi.SmpcBatch), C.c_int, C.c_void_p]
    b = sc.single("readme", n_people=3)
    st = b.struct()
    x = b.arrays["u0"][0].ravel()
    assert fn(C.byref(b.params), C.byref(st), 0, x.ctypes.data) == -1

"""nav2_social_mpc_controller_b200 — B200-native batched solver for the hot path of
PIC4SeR/nav2_social_mpc_controller (the Ceres MPC solve inside Optimizer::optimize).

The product is libsmpc.so (CUDA sm_100a kernels + C++ host library behind the C-ABI of
include/smpc.h). This package is the thin Python host layer over that ABI; importing it
does not load CUDA. There is no CPU fallback: every solve entry raises if the library or
a GPU is missing.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
__version__ = "0.1.0"

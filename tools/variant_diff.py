#!/usr/bin/env python
"""Do two launch mappings of the solve kernel (same lanes per problem, different warps per CTA / register budget)
return bit-identical results? usage: variant_diff.py A B  (crowd scenarios, A agents, B problems)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav2_social_mpc_controller_b200 import scenarios as sc  # noqa: E402
from nav2_social_mpc_controller_b200.optimizer import Optimizer  # noqa: E402

A, B = int(sys.argv[1]), int(sys.argv[2])
batch = sc.crowd(B=B, A=A, config_id=5) if A else sc.corridor(B=B)
outs = {}
for warps, minb in (("4", "2"), ("16", "0"), ("4", "3")):
    os.environ.update(SMPC_WARPS=warps, SMPC_MINB=minb, SMPC_CHUNKS="1", SMPC_GROUP="32")
    opt = Optimizer(0)
    opt.initialize(batch.params)
    outs[(warps, minb)] = opt.solve_batch(batch)
    opt.close()
ref = outs[("4", "2")]
for k, o in outs.items():
    du = np.abs(o["u"] - ref["u"]).max(axis=(1, 2))
    print(f"A={A} B={B} W={k[0]} MB={k[1]}: identical u on {(du == 0).mean():.4f} of problems, max |du| {du.max():.3e}, "
          f"iterations differ on {(o['iterations'] != ref['iterations']).mean():.4f}")

import numpy as np, sys
sys.path.insert(0,'.')
from nav2_social_mpc_controller_b200 import scenarios as sc
from nav2_social_mpc_controller_b200.optimizer import Optimizer
from tests import oracle_lib
o=oracle_lib.load()
rng = np.random.default_rng(17)
B = 96
base = sc.crowd(B=B, A=3, config_id=6, n_valid=2)
n_each = rng.integers(1, base.n_steps + 1, size=B)
n_each[:4] = [base.n_steps, 1, 2, 7]
batch = sc.with_horizons(base, n_each)
opt=Optimizer(0); opt.initialize(batch.params); opt.set_group(32)
got=opt.solve_batch(batch, want=("u","cmds","path","cost_initial","cost_final","iterations","termination","usable","n_evals"))
ref=o.solve_batch(batch,n_threads=8)
for b in range(B):
    S_b=int(n_each[b]); ch,bl,nb_b,_=sc.abi.problem_dims(batch.params.control_horizon,batch.params.parameter_block_length,S_b)
    du=np.abs(got["u"][b,:nb_b]-ref["u"][b,:nb_b]).max(); dc=abs(got["cost_final"][b]-ref["cost_final"][b])/max(abs(ref["cost_final"][b]),1e-300)
    if du>1e-6 or dc>1e-8 or got["usable"][b]!=ref["usable"][b]:
        print(b,'S',S_b,'nb',nb_b,'du %.2e dc %.2e'%(du,dc),'it',got["iterations"][b],ref["iterations"][b],'term',got["termination"][b],ref["termination"][b],'ci %.6e %.6e'%(got["cost_initial"][b],ref["cost_initial"][b]))

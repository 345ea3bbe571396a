import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets
from oracle import binding as orc
ASYNC = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, rel_change_tol=0.0, max_num_iters=10**9)
def rel(a,b): return np.linalg.norm(a-b)/np.linalg.norm(b)
pb = datasets.make_synthetic_problem(4000, 30000, 2, seed=1)
oteam = orc.OracleTeam(pb, **ASYNC)
team, agents = gpu.make_team(pb, **ASYNC)
team.set_schedule(1)
for tick in range(4):
    # gradient the next tick will use, three ways
    for rid in range(2):
        X = agents[rid].getX()
        f, rg, _, _ = agents[rid].edgeGrad(None)
        fo, ego, rgo = oteam.eval(rid, oteam.get_x(rid))
        f2, eg2, rg2 = agents[rid].eval(X)
        print(f"tick {tick} robot {rid}: X vs oracle {rel(X, oteam.get_x(rid)):.2e} | edge rgrad vs oracle {rel(rg, rgo):.2e} | ell rgrad vs oracle {rel(rg2, rgo):.2e} | f {f:.6f} {fo:.6f}")
    team.run(1, stop_on_terminate=False)
    oteam.run_parallel(1, threads=2)
for rid in range(2):
    print("final", rid, rel(agents[rid].getX(), oteam.get_x(rid)))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets
from oracle import binding as orc
pb = datasets.load_g2o_problem("smallGrid3D", 2)
kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, restart_interval=7, rel_change_tol=1e-9, max_num_iters=10000)
oteam = orc.OracleTeam(pb, **kw); team, agents = gpu.make_team(pb, **kw)
for it in range(40):
    oteam.run(1, stop_on_terminate=False); team.run(1, stop_on_terminate=False)
    errs = [np.linalg.norm(agents[r].getX()-oteam.get_x(r))/np.linalg.norm(oteam.get_x(r)) for r in range(2)]
    print(it, ["%.2e" % e for e in errs], "cost %.6f" % oteam.global_cost())

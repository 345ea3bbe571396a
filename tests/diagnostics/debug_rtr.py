import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets
from oracle import binding as orc
pb = datasets.load_g2o_problem("sphere2500", 5)
kw = dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2)
oteam = orc.OracleTeam(pb, **kw); team, agents = gpu.make_team(pb, **kw)
for it in range(262):
    ro = oteam.run(1, stop_on_terminate=True); rg = team.run(1, stop_on_terminate=True)
    sel = it % 5
    if it % 20 == 0 or it > 245:
        errs = max(np.linalg.norm(agents[r].getX()-oteam.get_x(r))/np.linalg.norm(oteam.get_x(r)) for r in range(5))
        so = oteam.status(sel); sg = agents[sel].getStatus()
        oo = oteam.opt_result(sel); og = agents[sel].localOptResult()
        print(it, "maxerr %.2e" % errs, "relchange o/g %.6f %.6f" % (so.relative_change, sg.relative_change), "tcg", oo.tcg_iters, og.tcg_iters, "term", ro.terminated, rg.terminated)
    if ro.terminated and rg.terminated: break

"""ms per step of ONE 12 500-pose agent under the synchronous (M=1) and the parallel (M=2) kernels (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpgo_ros_b200 import agent as gpu, datasets
pb = datasets.make_synthetic_problem(12500, 125000, 1, seed=0)
kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, rel_change_tol=0.0, max_num_iters=10**9)
team, agents = gpu.make_team(pb, **kw)
for sched in (0, 1, 0, 1):
    team.set_schedule(sched)
    team.run(2, stop_on_terminate=False)
    res = team.run(10, stop_on_terminate=False)
    print("schedule", sched, "ms/step", res.device_ms / 10, "cost", team.global_cost())

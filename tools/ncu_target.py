"""Target for ncu captures: config-2 team, a few warm-up launches, then ONE K-iteration launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpgo_ros_b200 import agent as gpu, datasets
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
method = sys.argv[2] if len(sys.argv) > 2 else "rgd"
pb = datasets.load_g2o_problem("sphere2500", 8)
if method == "rgd":
    kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50, rel_change_tol=0.0, max_num_iters=10**9)
else:
    kw = dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.0, max_num_iters=10**9)
team, agents = gpu.make_team(pb, **kw)
team.run(20, stop_on_terminate=False)
team.run(20, stop_on_terminate=False)
res = team.run(K, stop_on_terminate=False)
print("iters", res.iterations, "us/iter", res.device_ms * 1e3 / K)

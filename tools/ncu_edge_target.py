"""Target for the ncu capture of k_edge_grad: one large agent (n = 8000, ~80 k edges), a few calls of the kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpgo_ros_b200 import agent as gpu, datasets
ASYNC = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=0, acceleration=0, rel_change_tol=0.0, max_num_iters=10 ** 9)
pb = datasets.make_synthetic_problem(16000, 160000, 2, seed=1)
team, agents = gpu.make_team(pb, **ASYNC)
for _ in range(4):
    f, rg, kns, ens = agents[0].edgeGrad(None, flush_l2=True)
print("k_edge_grad n =", pb.n[0], "edges =", len(pb.robot_measurements(0)), "us", kns * 1e-3)

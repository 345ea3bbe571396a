"""Quick perf probe (GPU): us per global RBCD iteration of the persistent kernel vs grid size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets

def main():
    name, robots = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("sphere2500", 8)
    pb = datasets.load_g2o_problem(name, robots)
    cfgs = {
        "rgd_accel_precond": dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50, rel_change_tol=0.0, max_num_iters=10**9),
        "rgd_accel_noprecond": dict(r=5, method=1, rgd_stepsize=1e-3, rgd_use_preconditioner=0, acceleration=1, restart_interval=50, rel_change_tol=0.0, max_num_iters=10**9),
        "rtr": dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.0, max_num_iters=10**9),
    }
    for cname, kw in cfgs.items():
        for grid in (148, 96, 64, 32, 16):
            team, agents = gpu.make_team(pb, **kw)
            team.set_grid(grid)
            team.exchange_all()
            team.run(50, stop_on_terminate=False)
            K = 1000 if kw["method"] == 1 else 100
            res = team.run(K, stop_on_terminate=False)
            print(f"{name}/{robots} {cname:22s} grid={grid:4d}  {res.device_ms*1e3/K:9.2f} us/iter  ({K} iters, {res.kernel_launches} launches)", flush=True)
            team.close()
            for a in agents: a.close()

if __name__ == "__main__":
    main()

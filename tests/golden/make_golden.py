"""Generate tests/golden/golden.json from the CPU oracle (PARITY UNPINNED: the reference ships no
optimiser golden vectors; these pin the oracle against regressions and give the GPU
tests a fixture that does not need the oracle at run time).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dpgo_ros_b200 import datasets  # noqa: E402
from oracle import binding as orc  # noqa: E402

CASES = {
    # BASELINE config 1: smallGrid3D, 2 agents, r=5, RGD, L2
    "config1_smallGrid3D_2_rgd": dict(dataset="smallGrid3D", robots=2, max_run=400,
                                      params=dict(r=5, method=1, rgd_stepsize=0.1, rgd_use_preconditioner=1,
                                                  acceleration=0, rel_change_tol=0.01, max_num_iters=1000)),
    # BASELINE config 2: sphere2500, 8 agents, RGD + Nesterov
    "config2_sphere2500_8_rgd_nesterov": dict(dataset="sphere2500", robots=8, max_run=2000,
                                              params=dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1,
                                                          acceleration=1, restart_interval=50, rel_change_tol=0.1,
                                                          max_num_iters=1000)),
    # README demo: sphere2500, 5 agents, RTR 3x50, tol 0.2 (README.md:32,44)
    "readme_sphere2500_5_rtr": dict(dataset="sphere2500", robots=5, max_run=1000,
                                    params=dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2)),
    # smallGrid3D, 2 agents, RTR to a tight tolerance (SE-Sync optimum 1025.398)
    "smallGrid3D_2_rtr_tight": dict(dataset="smallGrid3D", robots=2, max_run=3000,
                                    params=dict(r=5, method=0, gradnorm_tol=1e-3, rel_change_tol=1e-5,
                                                max_num_iters=2000)),
    # GNC-TLS schedule on smallGrid3D (3 weight updates, resets)
    "smallGrid3D_2_gnc": dict(dataset="smallGrid3D", robots=2, max_run=200,
                              params=dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2, cost_type=5, gnc_barc=3.0,
                                          gnc_mu_step=2.0, gnc_init_mu=1e-5, robust_opt_num_weight_updates=3,
                                          robust_opt_num_resets=3, robust_opt_inner_iters=10, max_num_iters=38)),
    # BASELINE config 5 at a size the oracle finishes in seconds: synthetic lattice graph (seed 0), 8 agents, the
    # asynchronous mode as parallel ticks (schedule 1), RGD 0.2 + preconditioner (launch/asapp_demo.launch:7-8)
    "config5_synthetic2000_8_parallel": dict(dataset="synthetic:2000:20000", robots=8, max_run=40, schedule=1,
                                             params=dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1,
                                                         acceleration=0, rel_change_tol=0.0, max_num_iters=10 ** 9)),
    # README demo from the Chordal initialisation (launch/dpgo_demo.launch:9): sphere2500, 5 agents, RTR, tol 0.2
    "readme_sphere2500_5_chordal_rtr": dict(dataset="sphere2500", robots=5, max_run=1000, init="chordal",
                                            params=dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2)),
}


def load_problem(c):
    if c["dataset"].startswith("synthetic:"):
        _, n, m = c["dataset"].split(":")
        pb = datasets.make_synthetic_problem(int(n), int(m), c["robots"], seed=0)
    else:
        pb = datasets.load_g2o_problem(c["dataset"], c["robots"])
    if c.get("init") == "chordal":
        o = orc.OracleTeam(pb, r=c["params"]["r"], initialize=False)
        pb = datasets.with_local_initialization(pb, lambda rid: o.initialize_chordal(rid))
    return pb


def fingerprint(X):
    idx = np.linspace(0, X.size - 1, 16).astype(int)
    return {"fro": float(np.linalg.norm(X)), "sum": float(X.sum()), "samples": [float(v) for v in X.ravel(order="F")[idx]]}


def main():
    out = {}
    for name, c in CASES.items():
        pb = load_problem(c)
        team = orc.OracleTeam(pb, **c["params"])
        res = team.run_parallel(c["max_run"], threads=4) if c.get("schedule") else team.run(c["max_run"], threads=4)
        out[name] = {"dataset": c["dataset"], "robots": c["robots"], "params": c["params"], "max_run": c["max_run"],
                     "schedule": c.get("schedule", 0), "init": c.get("init", "odometry"),
                     "iterations": res.iterations, "terminated": bool(res.terminated),
                     "weight_updates": res.weight_updates, "final_cost_2f": team.global_cost(),
                     "X": [fingerprint(team.get_x(r)) for r in range(c["robots"])]}
        print(name, res.iterations, res.terminated, out[name]["final_cost_2f"])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

"""BASELINE config 5 at the named size, ONE of its 8 ranks on one GPU: synthetic 100k-pose / 1M-edge SE(3) graph
(datasets.make_synthetic_problem, seed 0), 8 agents of 12 500 poses, RGD step 0.2 with the dense preconditioner
(launch/asapp_demo.launch:7-8).  Robot `--robot` is built as a stand-alone agent, its neighbours' public poses come
from the initial guess, and iterate(true) is timed: at this size the step streams the 20 GB preconditioner once,
so it is the HBM-bound regime SURVEY 8d names for the roofline statement.

    python tools/bench_config5.py [--poses 100000 --edges 1000000 --robots 8 --robot 0 --iters 10]
"""
import argparse, json, os, sys, time
os.environ["DPGO_B200_SYM_PRECOND"] = "1"   # allocate the symmetric pass's buffers with the agent (variant 1 below)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets, capi

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=100000)
ap.add_argument("--edges", type=int, default=1000000)
ap.add_argument("--robots", type=int, default=8)
ap.add_argument("--robot", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--r", type=int, default=5)
args = ap.parse_args()

t0 = time.time()
pb = datasets.make_synthetic_problem(args.poses, args.edges, args.robots, seed=0)
t_gen = time.time() - t0
rid = args.robot
kw = dict(r=args.r, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, rel_change_tol=0.0,
          max_num_iters=10 ** 9, num_robots=args.robots)
P = gpu.make_params(**kw)
yl = datasets.fixed_lifting_matrix(P.r)
eye = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)
ag = gpu.PGOAgent(rid, P, 0)
m = pb.robot_measurements(rid)
ag.addMeasurements(m)
ag.setLiftingMatrix(yl)
ag.initialize(pb.T_init[rid])
ag.initializeInGlobalFrame(eye)
# neighbours' public poses from the (global-frame) initial guess, lifted the same way
sh = m.r1 != m.r2
need = {}
for e in np.nonzero(sh)[0]:
    o, f = (int(m.r2[e]), int(m.p2[e])) if int(m.r1[e]) == rid else (int(m.r1[e]), int(m.p1[e]))
    need.setdefault(o, set()).add(f)
for o, frames in need.items():
    fr = np.array(sorted(frames), dtype=np.int32)
    poses = np.einsum("ak,nkc->nca", yl, pb.T_init[o][fr])   # [cnt][4][r]: r x 4 column-major per pose
    ag.updateNeighborPoses(o, fr, np.ascontiguousarray(poses), False)
n = pb.n[rid]
edges = len(m)
t0 = time.time()
ag.iterate(True)          # first call builds Q, the dense inverse and everything else
t_setup = time.time() - t0
L = capi.lib()
variants = {}
for name, env in (("edge-record gradient + one triangle of Pinv (opt-in)", {"DPGO_B200_SYM_PRECOND": "1"}),
                  ("edge-record gradient + full Pinv pass inside the persistent kernel (default)", {}),
                  ("everything inside the persistent kernel (round 1)", {"DPGO_B200_NO_EDGE_GRAD": "1"})):
    for k in ("DPGO_B200_SYM_PRECOND", "DPGO_B200_NO_EDGE_GRAD"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ag.iterate(True)
    tv = []
    for _ in range(args.iters):
        t0 = time.perf_counter()
        ag.iterate(True)
        tv.append(time.perf_counter() - t0)
    variants[name] = float(np.median(tv)) * 1e3
for k in ("DPGO_B200_SYM_PRECOND", "DPGO_B200_NO_EDGE_GRAD"):
    os.environ.pop(k, None)
ts = []
for _ in range(args.iters):
    t0 = time.perf_counter()
    ag.iterate(True)
    ts.append(time.perf_counter() - t0)
opt = ag.localOptResult()
ms = float(np.median(ts)) * 1e3
grad_us = None
try:
    g = [ag.edgeGrad(None, flush_l2=True)[2] for _ in range(7)]
    gw = [ag.edgeGrad(None, flush_l2=False)[2] for _ in range(7)]
    grad_us = {"cold_l2": float(np.median(g)) * 1e-3, "warm_l2": float(np.median(gw)) * 1e-3}
except Exception as e:  # noqa: BLE001
    grad_us = {"error": str(e)[:200]}
n4 = 4 * n
npad = (n4 + 31) // 32 * 32
b_precond = npad * npad * 8 + 2 * n * args.r * 4 * 8
b_grad = edges * 128 + 2 * n * args.r * 4 * 8 + sum(len(v) for v in need.values()) * args.r * 4 * 8
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
achieved = (b_precond + b_grad) / (ms * 1e-3) / 1e9
print(json.dumps({"workload": f"synthetic {args.poses} poses / {args.edges} edges / {args.robots} agents, robot {rid} "
                              f"(n={n}, {edges} edges, {len(need)} neighbours), r={args.r}, RGD 0.2 + dense preconditioner",
                  "ms_per_iterate": ms, "ms_per_iterate_by_variant": variants, "k_edge_grad_us": grad_us, "setup_s": t_setup, "generate_s": t_gen, "f_init": opt.f_init,
                  "gradnorm_init": opt.gradnorm_init, "relative_change": opt.relative_change,
                  "algorithmic_bytes": b_precond + b_grad, "achieved_GBps": achieved, "peak_GBps": peak,
                  "frac": achieved / peak}))

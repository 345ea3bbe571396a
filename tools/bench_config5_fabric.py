"""BASELINE config 5 at the named size over the multi-GPU fabric: synthetic 100k-pose / 1M-edge SE(3) graph, 8 agents
of 12 500 poses, one agent per GPU, asynchronous mode (RGD step 0.2 + dense preconditioner, launch/asapp_demo.launch:7-8)
as parallel ticks.  Run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_config5_fabric.py [--ticks 20]

Every tick each GPU streams its 20 GB preconditioner once; the public poses cross NVLink inside the kernels.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dpgo_ros_b200 import datasets, dist as ddist

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=100000)
ap.add_argument("--edges", type=int, default=1000000)
ap.add_argument("--robots", type=int, default=8)
ap.add_argument("--ticks", type=int, default=20)
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
pb = datasets.make_synthetic_problem(args.poses, args.edges, args.robots, seed=0)
kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, rel_change_tol=0.0,
          max_num_iters=10 ** 9)
t0 = time.time()
rt = ddist.GpuRankTeam(pb, rank, world, local, fabric=True, schedule=1, **kw)
rt.run(1, False)           # first tick builds Q, the dense inverse (50 016^2) and the wiring
torch.cuda.synchronize()
dist.barrier()
setup = time.time() - t0


def cost():
    Xs = {rid: ag.getX() for rid, ag in rt.agents.items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, Xs)
    if rank != 0:
        return None
    allX = {}
    for g in gathered:
        allX.update(g)
    m = pb.meas
    c = 0.0
    for lo in range(0, len(m), 100000):   # vectorised global cost 2f, in slices
        sl = slice(lo, min(lo + 100000, len(m)))
        Xi = np.stack([allX[int(a)][:, 4 * int(p):4 * int(p) + 4] for a, p in zip(m.r1[sl], m.p1[sl])])
        Xj = np.stack([allX[int(a)][:, 4 * int(p):4 * int(p) + 4] for a, p in zip(m.r2[sl], m.p2[sl])])
        rot = np.einsum("nak,nkc->nac", Xi[:, :, :3], m.R[sl]) - Xj[:, :, :3]
        tr = Xj[:, :, 3] - Xi[:, :, 3] - np.einsum("nak,nk->na", Xi[:, :, :3], m.t[sl])
        c += float(np.sum(m.weight[sl] * (m.kappa[sl] * np.sum(rot * rot, axis=(1, 2)) + m.tau[sl] * np.sum(tr * tr, axis=1))))
    return c


c0 = cost()
torch.cuda.synchronize()
dist.barrier()
done, _, _, ms = rt.run(args.ticks, False)
torch.cuda.synchronize()
t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
c1 = cost()
if rank == 0:
    ms_tick = float(t.item()) / args.ticks
    n = pb.n[0]
    npad = (4 * n + 31) // 32 * 32
    bytes_per_gpu = npad * npad * 8
    peak = 6650.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    per_gpu = bytes_per_gpu * (args.robots // world) / (ms_tick * 1e-3) / 1e9
    print(json.dumps({"workload": f"synthetic {args.poses} poses / {args.edges} edges / {args.robots} agents over {world} GPUs, "
                                  "asynchronous mode as parallel ticks, RGD 0.2 + dense preconditioner, fabric",
                      "ticks": args.ticks, "ms_per_tick": ms_tick, "robot_updates_per_s": args.robots / (ms_tick * 1e-3),
                      "setup_s": setup, "cost_2f_before": c0, "cost_2f_after": c1,
                      "preconditioner_GBps_per_gpu": per_gpu, "peak_GBps": peak, "frac_per_gpu": per_gpu / peak,
                      "aggregate_TBps": per_gpu * world / 1e3}))
dist.barrier()
dist.destroy_process_group()

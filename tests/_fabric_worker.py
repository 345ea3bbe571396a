"""Worker of tests/test_gpu_fabric_multi.py: one rank of a real multi-GPU fabric run (torchrun, NCCL for the
bootstrap only).  Rank 0 compares the gathered iterates with a single-team run of the same schedule."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpgo_ros_b200 import agent as gpu  # noqa: E402
from dpgo_ros_b200 import datasets  # noqa: E402
from dpgo_ros_b200 import dist as ddist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    mode = sys.argv[1]
    pb = datasets.load_g2o_problem("sphere2500", 8)
    if mode == "shm":
        return shm_mode(pb, rank, world, local % torch.cuda.device_count())
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if mode == "sync":
        kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50,
                  rel_change_tol=0.1)
        schedule, iters, stop = 0, 2000, True
    else:
        kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, rel_change_tol=0.0,
                  max_num_iters=10 ** 9)
        schedule, iters, stop = 1, 200, False
    rt = ddist.GpuRankTeam(pb, rank, world, local, fabric=True, schedule=schedule, **kw)
    done, term, wu, ms = rt.run(iters, stop)
    Xs = {rid: ag.getX() for rid, ag in rt.agents.items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, (done, term, Xs))
    if rank == 0:
        assert len({(g[0], g[1]) for g in gathered}) == 1, "ranks disagree on the iteration count"
        team, agents = gpu.make_team(pb, device=local, **kw)
        team.set_schedule(schedule)
        res = team.run(iters, stop_on_terminate=stop)
        assert res.iterations == done and bool(res.terminated) == term, (res.iterations, done)
        allX = {}
        for g in gathered:
            allX.update(g[2])
        for a in agents:
            assert np.array_equal(a.getX(), allX[a.id]), f"robot {a.id}: fabric iterate differs from the single-team run"
        print(f"fabric {mode} ok: {done} iterations on {world} GPUs, bit-identical to one team "
              f"({ms * 1e3 / max(done, 1):.1f} us/iteration)")
    dist.barrier()
    dist.destroy_process_group()


def shm_mode(pb, rank, world, device):
    """Per-robot C ABI with host buffers, robots spread over processes, poses through shared memory
    (dpgo_b200_sync_driver_run_shm).  Works with several processes on one GPU too."""
    dist.init_process_group("gloo")
    kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=7,
              rel_change_tol=0.0, max_num_iters=10 ** 9)
    ht = ddist.ShmHostTeam(pb, rank, world, device, tag=f"test{os.environ.get('MASTER_PORT', '0')}", **kw)
    ht.run(25)
    ht.run(36)   # a second call continues the schedule (restart iterations land inside the lookahead chains)
    Xs = {a.id: a.getX() for a in ht.agents}
    its = {a.id: a.iteration_number() for a in ht.agents}
    gathered = [None] * world
    dist.all_gather_object(gathered, (Xs, its))
    ht.close()
    if rank == 0:
        team, agents = gpu.make_team(pb, device=device, **kw)
        team.run(61, stop_on_terminate=False)
        allX, allit = {}, {}
        for g in gathered:
            allX.update(g[0])
            allit.update(g[1])
        for a in agents:
            err = np.linalg.norm(a.getX() - allX[a.id]) / np.linalg.norm(allX[a.id])
            assert err < 1e-12, (a.id, err)
            assert allit[a.id] == 61
        print(f"fabric shm ok: 61 iterations over {world} processes match the single-team run")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""torchrun --nproc-per-node 2 tools/probe_arm_shm.py : the host-buffer e2e arm (dpgo_b200_sync_driver_run_shm) with ONE
robot per GPU -- the deployment armed launches are for -- with and without them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from dpgo_ros_b200 import datasets, dist as dd
import bench

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
os.environ["DPGO_B200_DRIVER_PROFILE"] = "1"
k = 0
for name in ("smallGrid3D", "sphere2500"):
    pb = datasets.load_g2o_problem(name, world)
    for tag in ("armed", "not armed", "armed", "not armed"):
        k += 1
        if tag == "armed":
            os.environ.pop("DPGO_B200_NO_ARM", None)
        else:
            os.environ["DPGO_B200_NO_ARM"] = "1"
        ht = dd.ShmHostTeam(pb, rank, world, local, tag=f"probe{k}", **bench.CONFIG2)
        ht.run(40)
        dist.barrier()
        sec = ht.run(2000)
        t = torch.tensor([sec], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name} / {world} robots, one per GPU, {tag}: {float(t.item()) / 2000 * 1e6:.1f} us per global iteration", flush=True)
        ht.close()
dist.barrier()
dist.destroy_process_group()

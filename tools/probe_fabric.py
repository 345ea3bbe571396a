"""torchrun --nproc-per-node N tools/probe_fabric.py : the synchronous schedule over the fabric, us per step for the
protocol variants of DPGO_B200_FAB_VARIANT, with the phase profile of one CTA (DPGO_B200_FAB_PROF)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dpgo_ros_b200 import datasets, dist as dd
import bench

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
pb = datasets.load_g2o_problem("sphere2500", 8)
rt = dd.GpuRankTeam(pb, rank, world, local, fabric=True, **bench.CONFIG2)
rt.run(200, False)
for variant in (0, 1, 2, 3, 0):
    os.environ["DPGO_B200_FAB_VARIANT"] = str(variant)
    with open(f"gpurun_out/fabprof_rank{rank}.txt", "a") as f:
        f.write(f"== variant {variant}\n")
    rt.run(2000, False)
    torch.cuda.synchronize(); dist.barrier()
    out = []
    for steps in (20, 4000):
        _, _, _, ms = rt.run(steps, False)
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t.item()) * 1e3 / steps)
    if rank == 0:
        print(f"variant {variant}: {out[0]:.2f} us/step at 20 steps, {out[1]:.2f} us/step at 4000 steps", flush=True)
dist.barrier()
dist.destroy_process_group()

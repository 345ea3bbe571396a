"""Worker of tests/test_dist_cpu.py: world_size-2 gloo run of the multi-GPU exchange plan with mock endpoints."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpgo_ros_b200 import datasets  # noqa: E402
from dpgo_ros_b200 import dist as ddist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    name, robots, accel = sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3]))
    pb = datasets.load_g2o_problem(name, robots)
    r = 5
    nbrs, counts = {}, {}
    for rid in range(robots):
        m = pb.robot_measurements(rid)
        sh = m.r1 != m.r2
        other = np.where(m.r1[sh] == rid, m.r2[sh], m.r1[sh])
        mine = np.where(m.r1[sh] == rid, m.p1[sh], m.p2[sh])
        nbrs[rid] = sorted({int(x) for x in other})
        for b in nbrs[rid]:
            counts[(rid, b)] = len({int(f) for f, o in zip(mine, other) if o == b})
    plan = ddist.build_plan(nbrs, robots, world, accel)
    local = ddist.robots_of_rank(robots, world, rank)
    assert sorted(sum((ddist.robots_of_rank(robots, world, k) for k in range(world)), [])) == list(range(robots))

    def fill(robot, nbr, aux, step):
        n = counts[(robot, nbr)] * 4 * r
        return torch.arange(n, dtype=torch.float64) + 1000.0 * robot + 100.0 * nbr + 10.0 * aux + 1e6 * step

    outbox, inbox, marked = {}, {}, set()
    for a in local:
        for b in nbrs[a]:
            for aux in ((False, True) if accel else (False,)):
                outbox[(a, b, aux)] = fill(a, b, aux, 0)
                inbox[(a, b, aux)] = torch.zeros(counts[(b, a)] * 4 * r, dtype=torch.float64)
    total = 0
    for step in range(3):
        for k in outbox:
            outbox[k] = fill(k[0], k[1], k[2], step)
        sel = step % robots
        senders = [x for x in range(robots) if x != sel] if step < 2 else [sel]
        total += ddist.exchange(plan, rank, senders, lambda a, b, x: outbox[(a, b, x)], lambda a, b, x: inbox[(a, b, x)],
                                lambda a, b, x: marked.add((a, b, x, step)), dist)
        for (a, b, aux), buf in inbox.items():
            if ddist.rank_of_robot(robots, world, b) == rank:
                continue  # co-located neighbour: not part of the cross-rank plan
            if b in senders:
                assert torch.equal(buf, fill(b, a, aux, step)), (rank, a, b, aux, step)
                assert (a, b, aux, step) in marked
    dist.barrier()
    print(f"rank {rank} ok: {len(plan)} transfers in plan, {total} bytes received")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

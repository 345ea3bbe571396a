"""world_size-2 gloo test of the multi-GPU public-pose exchange plan (host logic of dpgo_ros_b200/dist.py)."""
import os
import socket
import subprocess
import sys

import pytest

from dpgo_ros_b200 import dist as ddist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_robot_to_rank_partition():
    for world in (1, 2, 3, 4, 8):
        seen = []
        for rk in range(world):
            seen += ddist.robots_of_rank(8, world, rk)
        assert seen == list(range(8))
        for rid in range(8):
            assert rid in ddist.robots_of_rank(8, world, ddist.rank_of_robot(8, world, rid))
    assert ddist.robots_of_rank(8, 2, 1) == [4, 5, 6, 7]


def test_plan_only_holds_cross_rank_pairs():
    nbrs = {a: [b for b in (a - 1, a + 1) if 0 <= b < 8] for a in range(8)}  # the sphere2500/8 chain
    plan = ddist.build_plan(nbrs, 8, 2, accelerated=True)
    assert {(t.src_robot, t.dst_robot) for t in plan} == {(3, 4), (4, 3)}
    assert len(plan) == 4  # regular + auxiliary, both directions
    assert len(ddist.build_plan(nbrs, 8, 8, accelerated=False)) == 14
    assert ddist.build_plan(nbrs, 8, 1, accelerated=True) == []


@pytest.mark.parametrize("name,robots,accel", [("smallGrid3D", 4, 1), ("sphere2500", 8, 1), ("smallGrid3D", 3, 0)])
def test_exchange_over_gloo_world2(name, robots, accel):
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), name,
                                       str(robots), str(accel)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {rank} ok" in out

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


def test_fabric_wiring_routes_every_cross_rank_pair():
    """Every rank exports, for each local robot b and neighbour a, where a's poses live in b's inbox; the routes of
    rank(a) must point exactly there (dpgo_b200_team_fabric_route arguments)."""
    nbrs = {a: [b for b in (a - 1, a + 1) if 0 <= b < 8] for a in range(8)}
    world = 4
    offsets = []
    for rk in range(world):
        tab = {}
        for b in ddist.robots_of_rank(8, world, rk):
            for a in nbrs[b]:
                tab[(b, a)] = (1000 * b + 10 * a, 1000 * b + 10 * a + 5)
        offsets.append(tab)
    for rk in range(world):
        routes = ddist.fabric_wiring(offsets, 8, world, rk, nbrs)
        local = ddist.robots_of_rank(8, world, rk)
        want = {(a, b) for a in local for b in nbrs[a] if b not in local}
        assert {(r[0], r[1]) for r in routes} == want
        for a, b, peer, off_reg, off_aux in routes:
            assert peer == ddist.rank_of_robot(8, world, b)
            assert (off_reg, off_aux) == (1000 * b + 10 * a, 1000 * b + 10 * a + 5)
    with pytest.raises(KeyError):
        ddist.fabric_wiring([{}, {}, {}, {}], 8, world, 0, nbrs)


def test_owned_weight_updates_follow_the_lower_id_rule():
    """publishMeasurementWeights: the lower-ID robot owns a shared edge's weight (src/PGOAgentROS.cpp:732, 1340)."""
    import numpy as np
    # robot 1 (rank 0 of 2, robots 0-3 local) shares edges with robots 0, 2 (local) and 5 (remote)
    r1 = np.array([1, 0, 1, 5], dtype=np.int32)
    p1 = np.array([3, 7, 4, 2], dtype=np.int32)
    r2 = np.array([2, 1, 5, 1], dtype=np.int32)
    p2 = np.array([0, 1, 9, 8], dtype=np.int32)
    w = np.array([0.5, 0.25, 0.75, 0.125])
    fx = np.zeros(4, dtype=np.uint8)
    out = ddist.owned_weight_updates({1: (r1, p1, r2, p2, w, fx)}, 8, 2, 0)
    assert set(out) == {1}
    assert sorted(out[1]) == [(5, 1, 4, 5, 9, 0.75, 0), (5, 5, 2, 1, 8, 0.125, 0)]
    # the higher-ID end never sends
    assert ddist.owned_weight_updates({5: (r1[2:], p1[2:], r2[2:], p2[2:], w[2:], fx[2:])}, 8, 2, 1) == {}

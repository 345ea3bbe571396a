"""Edge cases of the hot path through the C ABI: every supported relaxation rank, a robot without neighbours,
ragged teams (robots of very different sizes, one with a single shared edge), unsupported configurations."""
import numpy as np
import pytest

from dpgo_ros_b200 import agent as gpu
from dpgo_ros_b200 import datasets
from dpgo_ros_b200.capi import DpgoError
from oracle import binding as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("r", [3, 4, 6, 7, 8])
@pytest.mark.parametrize("method", [0, 1])
def test_every_relaxation_rank(r, method):
    """r = 3 ... 8 each has its own kernel instantiation (RGD and RTR); r = 5 is covered everywhere else."""
    pb = datasets.load_g2o_problem("tinyGrid3D", 2) if r == 3 else datasets.make_synthetic_problem(240, 1200, 3, seed=r)
    kw = (dict(method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=9)
          if method == 1 else dict(method=0, gradnorm_tol=1e-3, acceleration=0))
    kw.update(r=r, rel_change_tol=0.0, max_num_iters=10 ** 6)
    oteam = orc.OracleTeam(pb, **kw)
    team, agents = gpu.make_team(pb, **kw)
    iters = 12 if method == 1 else 6
    res = team.run(iters, stop_on_terminate=False)
    oteam.run(iters, stop_on_terminate=False)
    assert res.iterations == iters
    for rid in range(pb.num_robots):
        assert rel(agents[rid].getX(), oteam.get_x(rid)) < (1e-9 if method == 1 else 1e-6), (r, rid)


def test_single_robot_has_no_exchange():
    """One robot owns the whole graph: no neighbours, no inbox, no publication lists."""
    pb = datasets.load_g2o_problem("smallGrid3D", 1)
    kw = dict(r=5, method=0, gradnorm_tol=1e-2, acceleration=0, rel_change_tol=1e-3, max_num_iters=200)
    oteam = orc.OracleTeam(pb, **kw)
    team, agents = gpu.make_team(pb, **kw)
    assert agents[0].getNeighbors() == []
    res = team.run(200, stop_on_terminate=True)
    ores = oteam.run(200, stop_on_terminate=True)
    assert res.iterations == ores.iterations and res.terminated
    assert abs(team.global_cost() - oteam.global_cost()) < 1e-6 * oteam.global_cost()
    assert abs(team.global_cost() - 1025.398) < 1e-2        # SE-Sync's optimum for smallGrid3D


def test_ragged_team():
    """Robots of 9 ... 600 poses in one team (the chunk tables, slab sizes and grids differ per robot)."""
    pb = datasets.make_synthetic_problem(1000, 6000, 5, seed=11)
    # re-partition unevenly: robot sizes 9, 41, 150, 200, 600
    meas = pb.meas
    start = np.array([0, 9, 50, 200, 400, 1000])
    gid = np.concatenate([[0], np.cumsum(pb.n)])
    g1 = gid[meas.r1] + meas.p1
    g2 = gid[meas.r2] + meas.p2
    rob = lambda g: np.searchsorted(start, g, side="right") - 1
    m2 = meas.take(np.arange(len(meas)))
    m2.r1, m2.r2 = rob(g1).astype(np.int32), rob(g2).astype(np.int32)
    m2.p1, m2.p2 = (g1 - start[m2.r1]).astype(np.int32), (g2 - start[m2.r2]).astype(np.int32)
    m2.fixed = ((m2.r1 == m2.r2) & (m2.p1 + 1 == m2.p2)).astype(np.uint8)
    Tall = np.concatenate(pb.T_init)
    pb2 = datasets.Problem(name="ragged", num_robots=5, meas=m2, n=[int(start[i + 1] - start[i]) for i in range(5)],
                           T_init=[Tall[start[i]:start[i + 1]] for i in range(5)])
    kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=20,
              rel_change_tol=0.0, max_num_iters=10 ** 6)
    oteam = orc.OracleTeam(pb2, **kw)
    team, agents = gpu.make_team(pb2, **kw)
    team.run(35, stop_on_terminate=False)
    oteam.run(35, stop_on_terminate=False)
    for rid in range(5):
        assert rel(agents[rid].getX(), oteam.get_x(rid)) < 1e-9, rid
    assert team.global_cost() < 0.6 * orc.OracleTeam(pb2, **kw).global_cost()


def test_unsupported_configurations_are_refused():
    pb = datasets.load_g2o_problem("tinyGrid3D", 2)
    for bad in (dict(r=9), dict(r=2), dict(cost_type=2), dict(d=2)):
        with pytest.raises(DpgoError):
            gpu.make_team(pb, **dict(dict(r=5), **bad))
    # a team must hold one problem: same r, same robot count
    _, a5 = gpu.make_team(pb, colocate=False, r=5)
    _, a6 = gpu.make_team(pb, colocate=False, r=6)
    team = gpu.Team(0)
    team.add(a5[0])
    with pytest.raises(DpgoError):
        team.add(a6[1])


@pytest.mark.parametrize("name,robots", [("smallGrid3D", 2), ("sphere2500", 8), ("tinyGrid3D", 1)])
def test_chordal_initialization_matches_oracle(name, robots):
    """local_initialization_method Chordal (src/PGOAgentROSNode.cpp:106-112) on the device: dense inverse of the
    anchored block Laplacian, two products, polar factors -- against oracle Agent::initializeChordal."""
    pb = datasets.load_g2o_problem(name, robots)
    o = orc.OracleTeam(pb, r=5, initialize=False)
    P = gpu.make_params(r=5, num_robots=robots)
    locals_gpu = {}
    for rid in range(robots):
        ag = gpu.PGOAgent(rid, P, 0)
        ag.addMeasurements(pb.robot_measurements(rid))
        T = ag.initializeChordal()
        To = o.initialize_chordal(rid)
        assert T.shape == To.shape
        assert np.abs(T - To).max() < 1e-9 * max(1.0, np.abs(To).max()), rid
        locals_gpu[rid] = T
        ag.close()
    # the whole pipeline from that guess: same iteration count and iterate as the oracle from ITS Chordal guess
    if robots > 1:
        pg = datasets.with_local_initialization(pb, lambda rid: locals_gpu[rid])
        po = datasets.with_local_initialization(pb, lambda rid: o.local_trajectory(rid))
        kw = dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2, max_num_iters=300)
        team, agents = gpu.make_team(pg, **kw)
        oteam = orc.OracleTeam(po, **kw)
        res = team.run(300, stop_on_terminate=True)
        ores = oteam.run(300, stop_on_terminate=True)
        assert res.iterations == ores.iterations and res.terminated
        for rid in range(robots):
            assert rel(agents[rid].getX(), oteam.get_x(rid)) < 1e-6, rid


def test_recover_rewinds_the_iteration_number():
    """RECOVER (src/PGOAgentROS.cpp:1191-1209) writes mIterationNumber: the restart schedule of the acceleration follows
    the new numbering, and steps speculated under the old one are discarded."""
    pb = datasets.load_g2o_problem("sphere2500", 8)
    kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=7,
              rel_change_tol=0.0, max_num_iters=10 ** 6)
    oteam = orc.OracleTeam(pb, **kw)
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    gpu.exchange_host(agents, accel=True)

    def steps(first, count):
        for it in range(first, first + count):
            sel = it % 8
            for a in agents:
                if a.id != sel:
                    a.iterate(False)
                    oteam.iterate(a.id, False)
            gpu.exchange_host(agents, accel=True, only=[a.id for a in agents if a.id != sel])
            oteam.exchange_all()
            agents[sel].iterate(True)
            oteam.iterate(sel, True)
            gpu.exchange_host(agents, accel=True, only=[sel])
            oteam.exchange_all()

    steps(0, 11)
    for a in agents:
        a.setIterationNumber(3)
        oteam.set_iteration_number(a.id, 3)
    steps(3, 10)
    for a in agents:
        assert a.iteration_number() == 13
        assert rel(a.getX(), oteam.get_x(a.id)) < 1e-9, a.id


# ----------------------------------------------------------------------------------------------------------------------
# data matrices assembled on the device (dpgo_ros_b200/csrc/assemble.cu)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,robots", [("tinyGrid3D", 1), ("smallGrid3D", 2), ("sphere2500", 8), ("tunnels", 8)])
def test_device_assembled_q_matches_oracle(name, robots):
    """Per-edge accumulation kernel (k_assemble_values) against PoseGraph::denseQ() of the oracle, both copies the
    kernel writes (block-CSR for the dense inverse, ELL + overflow for the hot phases): <= 1e-13 relative."""
    pb = datasets.load_tunnels_problem() if name == "tunnels" else datasets.load_g2o_problem(name, robots)
    kw = dict(r=5, method=0)
    oteam = orc.OracleTeam(pb, **kw)
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    for a in agents:
        Qo, _ = oteam.dense_q(a.id)
        Qc, Qe = a.denseQ()
        assert rel(Qc, Qo) < 1e-13 and rel(Qe, Qo) < 1e-13, (a.id, rel(Qc, Qo), rel(Qe, Qo))
        assert np.array_equal(Qc, Qe)
    for a in agents:
        a.close()


def test_weight_change_reassembles_on_the_device():
    """setMeasurementWeight + clearDataMatrices (src/PGOAgentROS.cpp:1341-1351) and a plain setMeasurementWeight (C-ABI
    user, no clearDataMatrices): Q, the gradient and the preconditioner follow the new weight."""
    pb = datasets.load_g2o_problem("smallGrid3D", 2)
    kw = dict(r=5, method=0)
    oteam = orc.OracleTeam(pb, **kw)
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    gpu.exchange_host(agents, accel=False)
    a = agents[0]
    r1, p1, r2, p2, w, fx = a.sharedLoopClosures()
    e = 3
    key = (int(r1[e]), int(p1[e]), int(r2[e]), int(p2[e]))
    assert a.setMeasurementWeight(*key, 0.25, False)          # no clearDataMatrices on purpose
    oteam.set_measurement_weight(0, *key, 0.25, False)
    X = a.getX()
    f, eg, rg = a.eval(X)
    fo, ego, rgo = oteam.eval(0, X)
    assert abs(f - fo) <= 1e-12 * abs(fo) and rel(eg, ego) < 1e-12
    Qo, _ = oteam.dense_q(0)
    assert rel(a.denseQ()[0], Qo) < 1e-13
    V = np.asfortranarray(np.random.default_rng(0).standard_normal(X.shape))
    assert rel(a.precond(X, V), oteam.precond(0, X, V)) < 1e-8
    for a in agents:
        a.close()


def test_residuals_are_batched_and_cached():
    """computeMeasurementResidual for every loop closure (the TERMINATE handler, src/PGOAgentROS.cpp:1044-1057): ONE
    kernel launch serves them all; values match the oracle; a pose update invalidates the cache."""
    from dpgo_ros_b200 import capi
    pb = datasets.load_tunnels_problem()
    kw = dict(r=5, method=0, cost_type=5, gnc_barc=3.0)
    oteam = orc.OracleTeam(pb, **kw)
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    gpu.exchange_host(agents, accel=False)
    oteam.exchange_all()
    a = agents[7]
    r1, p1, r2, p2, w, fx = a.sharedLoopClosures()
    L = capi.lib()
    a.computeMeasurementResidual(int(r1[0]), int(p1[0]), int(r2[0]), int(p2[0]))   # everything built, cache filled
    before = L.dpgo_b200_kernel_launch_count()
    for e in range(len(w)):
        key = (int(r1[e]), int(p1[e]), int(r2[e]), int(p2[e]))
        got, want = a.computeMeasurementResidual(*key), oteam.compute_measurement_residual(7, *key)
        assert got is not None and abs(got - want) <= 1e-10 * max(1.0, abs(want)), (e, got, want)
    assert L.dpgo_b200_kernel_launch_count() == before, "cached residuals must not launch"
    a.iterate(True)
    oteam.iterate(7, True)
    key = (int(r1[5]), int(p1[5]), int(r2[5]), int(p2[5]))
    got, want = a.computeMeasurementResidual(*key), oteam.compute_measurement_residual(7, *key)
    assert abs(got - want) <= 1e-7 * max(1.0, abs(want))
    for a in agents:
        a.close()


def test_neighbour_inactive_from_the_start():
    """A robot that is inactive from INITIALIZE onwards (SET_ACTIVE_ROBOTS with a subset, src/PGOAgentROS.cpp:377-400)
    never sends its poses: its inbox slots stay empty, and iterate(true) must still optimise -- against the oracle, which
    skips inactive neighbours before the pose look-up."""
    pb = datasets.load_g2o_problem("sphere2500", 4)
    kw = dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2)
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    oteam = orc.OracleTeam(pb, **kw)
    live = [a for a in agents if a.id != 2]
    for a in live:
        a.setRobotActive(2, False)
        oteam.set_robot_active(a.id, 2, False)
    gpu.exchange_host(live, accel=False)          # robot 2 publishes nothing
    oteam.exchange_all()
    for it in range(6):
        sel = live[it % 3]
        assert sel.iterate(True) is True
        oteam.iterate(sel.id, True)
        gpu.exchange_host(live, accel=False, only=[sel.id])
        oteam.exchange_all()
    for a in live:
        assert rel(a.getX(), oteam.get_x(a.id)) < 1e-7, a.id
    # and without the deactivation the same call reports that the solve was skipped
    _, fresh = gpu.make_team(pb, colocate=False, **kw)
    gpu.exchange_host([b for b in fresh if b.id != 2], accel=False)
    assert fresh[1].iterate(True) is False and fresh[1].iteration_number() == 1
    for a in agents + fresh:
        a.close()


# ----------------------------------------------------------------------------------------------------------------------
# the dense SPD inverse behind the preconditioner (dpgo_ros_b200/csrc/dense_inverse.cu: recursive blocked
# Cholesky / triangular inverse / W^T W on the FP64 tensor path, 64 x 64 register-resident leaves)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [32, 64, 96, 100, 160, 448, 1248, 2016])
def test_spd_inverse_matches_lapack(n):
    """Random SPD matrices with condition number ~1e6 (the preconditioner's Q + 0.1 I is ill-conditioned too):
    ||A P - I|| at the level numpy's own inverse reaches, symmetric output, every recursion shape: a single leaf, a
    32-row leaf behind a 64-row one, sizes that are not multiples of the tile, both tile sizes of the product kernel."""
    rng = np.random.default_rng(n)
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (U * np.logspace(0, 6, n)) @ U.T
    A = 0.5 * (A + A.T)
    P, ms = gpu.spd_inverse(A)
    assert ms > 0.0
    assert np.array_equal(P, P.T)
    ref = np.linalg.inv(A)
    res_gpu = np.linalg.norm(A @ P - np.eye(n)) / np.sqrt(n)
    res_ref = np.linalg.norm(A @ ref - np.eye(n)) / np.sqrt(n)
    assert res_gpu < 10 * res_ref + 1e-12, (res_gpu, res_ref)
    assert rel(P, ref) < 1e-9, rel(P, ref)


def test_spd_inverse_of_the_preconditioner_matrix():
    """Q + 0.1 I of a sphere2500 robot (n = 312 poses, the bench workload): the inverse the preconditioner streams."""
    pb = datasets.load_g2o_problem("sphere2500", 8)
    oteam = orc.OracleTeam(pb, r=5, method=1)
    Q, _ = oteam.dense_q(0)
    A = Q + 0.1 * np.eye(Q.shape[0])
    P, _ = gpu.spd_inverse(A)
    assert rel(P, np.linalg.inv(A)) < 1e-10
    assert np.linalg.norm(A @ P - np.eye(A.shape[0])) / np.sqrt(A.shape[0]) < 1e-10


def test_spd_inverse_reports_an_indefinite_matrix():
    A = np.diag(np.r_[np.full(70, 2.0), -1.0, np.full(25, 2.0)])
    with pytest.raises(DpgoError):
        gpu.spd_inverse(A)

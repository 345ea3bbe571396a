"""Multi-GPU fabric (persistent kernel per rank, peer-memory publication, flag barriers) checked on ONE GPU:
`world` ranks of this process run side by side with small grids (dist.LocalFabric) and must reproduce the
single-team run of the same schedule -- same iterates, same iteration count, same GNC weights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RGD = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50,
           rel_change_tol=0.1)


def _single(problem, grid, iters, stop, **kw):
    from dpgo_ros_b200 import agent as gpu
    team, agents = gpu.make_team(problem, **kw)
    team.set_grid(grid)
    team.exchange_all()
    res = team.run(iters, stop_on_terminate=stop)
    X = {a.id: a.getX() for a in agents}
    w = {a.id: a.lcWeights() for a in agents}
    out = (res.iterations, bool(res.terminated), res.weight_updates, X, w)
    team.close()
    for a in agents:
        a.close()
    return out


def _fabric(problem, world, grid, iters, stop, **kw):
    from dpgo_ros_b200 import dist as ddist
    fab = ddist.LocalFabric(problem, world, grid=grid, **kw)
    done, term, wu, ms = fab.run(iters, stop_on_terminate=stop)
    ag = fab.all_agents()
    X = {rid: a.getX() for rid, a in ag.items()}
    w = {rid: a.lcWeights() for rid, a in ag.items()}
    fab.close()
    return done, term, wu, X, w


def _compare(a, b, tol):
    assert a[0] == b[0], f"iterations differ: {a[0]} vs {b[0]}"
    assert a[1] == b[1] and a[2] == b[2]
    for rid in a[3]:
        err = np.linalg.norm(a[3][rid] - b[3][rid]) / np.linalg.norm(b[3][rid])
        assert err <= tol, f"robot {rid}: fabric iterate differs from the single-team run by {err:.3e}"
        np.testing.assert_allclose(a[4][rid], b[4][rid], rtol=0, atol=1e-12)


@pytest.mark.parametrize("world", [2, 4])
def test_fabric_rgd_nesterov_matches_single_team(world):
    from dpgo_ros_b200 import datasets
    pb = datasets.load_g2o_problem("smallGrid3D", 4)
    ref = _single(pb, 24, 120, False, **RGD)
    got = _fabric(pb, world, 24, 120, False, **RGD)
    _compare(got, ref, 0.0)   # same arithmetic, same reduction order: bit-identical


def test_fabric_plain_rbcd_matches_single_team():
    from dpgo_ros_b200 import datasets
    pb = datasets.load_g2o_problem("smallGrid3D", 3)
    kw = dict(RGD, acceleration=0)
    _compare(_fabric(pb, 3, 24, 90, False, **kw), _single(pb, 24, 90, False, **kw), 0.0)


def test_fabric_rtr_matches_single_team():
    from dpgo_ros_b200 import datasets
    pb = datasets.load_g2o_problem("smallGrid3D", 4)
    kw = dict(r=5, method=0, rtr_iterations=3, rtr_tcg_iterations=50, gradnorm_tol=0.5, acceleration=1,
              restart_interval=30, rel_change_tol=0.05)
    _compare(_fabric(pb, 2, 24, 40, True, **kw), _single(pb, 24, 40, True, **kw), 0.0)


def test_fabric_sphere8_to_convergence(sphere8_problem):
    """BASELINE config 2 over 4 ranks: identical iteration-to-convergence count and final iterate."""
    ref = _single(sphere8_problem, 32, 2000, True, **RGD)
    got = _fabric(sphere8_problem, 4, 32, 2000, True, **RGD)
    assert ref[1] and ref[0] < 2000
    _compare(got, ref, 0.0)


def test_fabric_gnc_weight_updates_cross_ranks():
    from dpgo_ros_b200 import datasets
    pb = datasets.load_tunnels_problem()
    kw = dict(r=5, method=0, rtr_iterations=3, rtr_tcg_iterations=50, gradnorm_tol=0.5, acceleration=1,
              restart_interval=30, cost_type=5, gnc_barc=3.0, gnc_init_mu=1e-5, gnc_mu_step=2.0,
              robust_opt_num_weight_updates=2, robust_opt_num_resets=1, robust_opt_inner_iters=8,
              rel_change_tol=0.2, max_num_iters=60)
    ref = _single(pb, 16, 40, True, **kw)
    got = _fabric(pb, 4, 16, 40, True, **kw)
    assert ref[2] == 2
    _compare(got, ref, 0.0)


def test_fabric_missing_peer_times_out_cleanly():
    """A rank whose peer never launches must give up (no hang), report an error, and stay usable."""
    from dpgo_ros_b200 import datasets, dist as ddist
    from dpgo_ros_b200.capi import DpgoError
    pb = datasets.load_g2o_problem("smallGrid3D", 2)
    fab = ddist.LocalFabric(pb, 2, grid=8, **RGD)
    fab.teams[0].fabric_set_timeout(0.3)
    with pytest.raises(DpgoError):
        fab.teams[0].fabric_run(10, False)
    fab.close()

"""The DPGO:: C++ shim (include/DPGO/*.h): the source-compatible surface `PGOAgentROS : public PGOAgent`
(include/dpgo_ros/PGOAgentROS.h:121) builds against, forwarding to the C ABI.  CPU: the headers compile with
-Wall -Wextra, the host-side pieces behave, and the harness refuses to run without a GPU (no CPU fallback).
GPU: a ROS-free stand-in for PGOAgentROS (tests/cpp/shim_harness.cpp) replays the synchronous protocol through
the shim and must land on the oracle's iterates."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")
LIBDIR = os.path.join(ROOT, "dpgo_ros_b200")


def _compile(src, out, link):
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", src), "-o", out]
    if link:
        cmd += ["-L" + LIBDIR, "-ldpgo_b200", "-Wl,-rpath," + LIBDIR]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    return out


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    from dpgo_ros_b200 import capi
    capi.lib()  # the shared library must exist
    return _compile("shim_harness.cpp", str(tmp_path_factory.mktemp("shim") / "shim_harness"), link=True)


def test_shim_host_side(tmp_path):
    exe = _compile("shim_host_checks.cpp", str(tmp_path / "shim_host_checks"), link=False)
    res = subprocess.run([exe, DATA], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "shim host checks ok" in res.stdout


def test_wire_formats_and_iteration_log(tmp_path):
    """ROS-free mirror of msg/*.msg + src/utils.cpp codecs + the CSV log columns (include/dpgo_ros_wire/wire.h); the
    checks repeat the reference's own unit test (tests/testUtils.cpp:16-70)."""
    exe = _compile("wire_checks.cpp", str(tmp_path / "wire_checks"), link=False)
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "wire checks ok" in res.stdout


def test_cmake_package_finds_the_shim(tmp_path):
    """`find_package(DPGO REQUIRED)` + `target_link_libraries(... DPGO)` -- the two lines dpgo_ros's own build uses
    (CMakeLists.txt:6, 151-154) -- resolve against cmake/DPGOConfig.cmake."""
    import shutil
    if shutil.which("cmake") is None:
        pytest.skip("cmake not available")
    from dpgo_ros_b200 import capi
    capi.lib()
    (tmp_path / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.10)\nproject(shimcheck CXX)\nfind_package(DPGO REQUIRED)\n"
        f"add_executable(shim_harness {os.path.join(ROOT, 'tests', 'cpp', 'shim_harness.cpp')})\n"
        "target_link_libraries(shim_harness DPGO)\n")
    b = tmp_path / "build"
    r1 = subprocess.run(["cmake", "-S", str(tmp_path), "-B", str(b), "-DDPGO_DIR=" + os.path.join(ROOT, "cmake")],
                        capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    r2 = subprocess.run(["cmake", "--build", str(b)], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert (b / "shim_harness").exists()


def test_shim_harness_builds_and_refuses_to_run_without_gpu(harness, tmp_path):
    from dpgo_ros_b200 import capi
    if capi.lib().dpgo_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    res = subprocess.run([harness, os.path.join(DATA, "smallGrid3D.g2o"), "2", "5", "rgd", str(tmp_path / "o")],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 3
    assert "no usable CUDA device" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,robots,mode,iters", [("sphere2500", 8, "rgd", 48), ("smallGrid3D", 2, "rtr", 12)])
def test_wrapper_standin_on_the_shim_matches_oracle(harness, tmp_path, name, robots, mode, iters):
    """(RGD + Nesterov on smallGrid3D amplifies rounding differences ~30x every 4 iterations, so the accelerated
    run uses the contractive config-2 workload.)"""
    from dpgo_ros_b200 import datasets
    from oracle import binding as orc
    prefix = str(tmp_path / "run")
    res = subprocess.run([harness, os.path.join(DATA, name + ".g2o"), str(robots), str(iters), mode, prefix],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    if name == "smallGrid3D":
        assert "robot 0: 62 poses, 61 odometry, 71 private LC, 30 shared LC, state 2" in out
        assert "robot 1: 63 poses, 62 odometry, 73 private LC, 30 shared LC, state 2" in out
    else:
        assert "robot 7: 316 poses, 315 odometry, 266 private LC, 51 shared LC, state 2" in out
    pb = datasets.load_g2o_problem(name, robots)
    if mode == "rgd":
        kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50,
                  rel_change_tol=0.1, max_num_iters=100000)
    else:
        kw = dict(r=5, method=0, gradnorm_tol=0.5, acceleration=0, rel_change_tol=0.1, max_num_iters=100000)
    oteam = orc.OracleTeam(pb, **kw)
    ores = oteam.run(iters, stop_on_terminate=True)
    line = [l for l in out.splitlines() if l.startswith("iterations")][0].split()
    assert int(line[1]) == ores.iterations, (line, ores.iterations)
    for rid in range(robots):
        X = np.fromfile(f"{prefix}_X{rid}.bin").reshape((5, -1), order="F")
        Xo = oteam.get_x(rid)
        assert X.shape == Xo.shape
        err = np.linalg.norm(X - Xo) / np.linalg.norm(Xo)
        assert err < (1e-9 if mode == "rgd" else 1e-6), (rid, err)
        # global-frame read-out (a11): d x (d+1) n trajectory, rotations in SO(3), robot 0's first pose at the origin
        T = np.fromfile(f"{prefix}_T{rid}.bin").reshape((3, -1), order="F")
        assert T.shape == (3, 4 * pb.n[rid])
        R = T.reshape(3, -1, 4, order="F")[:, :, :3] if False else np.stack([T[:, 4 * i:4 * i + 3] for i in range(pb.n[rid])])
        assert np.allclose(np.einsum("nij,nkj->nik", R, R), np.eye(3), atol=1e-9)
        assert np.allclose(np.linalg.det(R), 1.0, atol=1e-9)
        # rounding agrees with Y_a^T X_i computed from the oracle's iterate
        Ya, pa = oteam.get_x(0)[:, 0:3], oteam.get_x(0)[:, 3]
        want_t = Ya.T @ (Xo[:, 3::4] - pa[:, None])
        assert np.allclose(T[:, 3::4], want_t, atol=1e-6 * max(1.0, np.abs(want_t).max()))
    assert "robot 0:" in out and "trajectory 1 first_t" in out


@pytest.mark.gpu
def test_asynchronous_mode_through_the_shim(harness, tmp_path):
    """asynchronous = true (src/PGOAgentROSNode.cpp:80-93): the shim owns one optimisation thread per agent (Poisson
    clock), the wrapper stand-in only publishes when mPublishAsynchronousRequested is raised (:119-127).  Not
    deterministic, so the check is what SURVEY 8d asks of config 5: everybody iterated and the cost came down."""
    import sys
    sys.path.insert(0, ROOT)
    from dpgo_ros_b200 import datasets
    from dpgo_ros_b200.dist import _global_cost
    prefix = str(tmp_path / "async")
    res = subprocess.run([harness, os.path.join(DATA, "sphere2500.g2o"), "4", "600", "async", prefix],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    pb = datasets.load_g2o_problem("sphere2500", 4)
    X = {rid: np.fromfile(f"{prefix}_X{rid}.bin").reshape((5, -1), order="F") for rid in range(4)}
    yl = datasets.fixed_lifting_matrix(5)
    X0 = {rid: np.concatenate([yl @ pb.T_init[rid][i] for i in range(pb.n[rid])], axis=1) for rid in range(4)}
    c0, c1 = _global_cost(pb, X0, 5), _global_cost(pb, X, 5)
    its = [int(l.split("iteration")[1].split()[0]) for l in res.stdout.splitlines() if l.startswith("robot") and "iteration" in l]
    assert len(its) == 4 and min(its) > 20, res.stdout
    assert "async:" in res.stdout and int(res.stdout.split("async:")[1].split()[0]) > 20
    assert c1 < 0.1 * c0, (c0, c1)


@pytest.mark.gpu
def test_cross_robot_initialization_through_the_shim(harness, tmp_path):
    """multirobot_initialization (src/PGOAgentROSNode.cpp:120): only robot 0 is placed in the global frame; the others
    start from their local odometry chain and initialise themselves inside updateNeighborPoses from the shared loop
    closures with an initialised neighbour (robust transform averaging).  Same result as the Python harness'
    alignment (datasets.robust_frame_alignment) fed to the oracle."""
    from dpgo_ros_b200 import datasets
    from oracle import binding as orc
    prefix = str(tmp_path / "multi")
    res = subprocess.run([harness, os.path.join(DATA, "smallGrid3D.g2o"), "2", "12", "rtr_multi", prefix],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "multi-robot initialisation: all 2 robots initialized" in res.stdout
    pb = datasets.load_g2o_problem("smallGrid3D", 2)

    def local(rid):
        R, t = datasets.odometry_chain(pb.robot_measurements(rid), rid, pb.n[rid])
        return np.concatenate([R, t[:, :, None]], axis=2)
    pbl = datasets.with_local_initialization(pb, local)
    kw = dict(r=5, method=0, gradnorm_tol=0.5, acceleration=0, rel_change_tol=0.1, max_num_iters=100000)
    oteam = orc.OracleTeam(pbl, **kw)
    ores = oteam.run(12, stop_on_terminate=True)
    line = [l for l in res.stdout.splitlines() if l.startswith("iterations")][0].split()
    assert int(line[1]) == ores.iterations
    for rid in range(2):
        X = np.fromfile(f"{prefix}_X{rid}.bin").reshape((5, -1), order="F")
        Xo = oteam.get_x(rid)
        assert np.linalg.norm(X - Xo) / np.linalg.norm(Xo) < 1e-6, rid

"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- see dpgo_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NATIVE = False


class OrcParams(C.Structure):
    _fields_ = [
        ("d", C.c_int), ("r", C.c_int), ("num_robots", C.c_int),
        ("method", C.c_int),
        ("rgd_stepsize", C.c_double),
        ("rgd_use_preconditioner", C.c_int),
        ("rtr_iterations", C.c_int), ("rtr_tcg_iterations", C.c_int),
        ("rtr_initial_radius", C.c_double), ("gradnorm_tol", C.c_double),
        ("acceleration", C.c_int), ("restart_interval", C.c_int),
        ("cost_type", C.c_int),
        ("gnc_barc", C.c_double), ("gnc_mu_step", C.c_double), ("gnc_init_mu", C.c_double),
        ("robust_opt_num_weight_updates", C.c_int), ("robust_opt_num_resets", C.c_int),
        ("robust_opt_inner_iters", C.c_int),
        ("robust_opt_min_convergence_ratio", C.c_double),
        ("max_num_iters", C.c_int),
        ("rel_change_tol", C.c_double),
        ("precond_lambda", C.c_double),
    ]


class OrcRunResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("terminated", C.c_int), ("weight_updates", C.c_int),
                ("wall_seconds", C.c_double)]


class OrcOptResult(C.Structure):
    _fields_ = [("success", C.c_int), ("f_init", C.c_double), ("f_opt", C.c_double),
                ("gradnorm_init", C.c_double), ("gradnorm_opt", C.c_double), ("relative_change", C.c_double),
                ("rtr_outer_iters", C.c_int), ("tcg_iters", C.c_int), ("rtr_rejections", C.c_int)]


class OrcStatus(C.Structure):
    _fields_ = [("agent_id", C.c_int), ("state", C.c_int), ("instance_number", C.c_int),
                ("iteration_number", C.c_int), ("ready_to_terminate", C.c_int), ("relative_change", C.c_double)]


DEFAULTS = dict(
    d=3, r=5, num_robots=1, method=0, rgd_stepsize=1e-3, rgd_use_preconditioner=1, rtr_iterations=3,
    rtr_tcg_iterations=50, rtr_initial_radius=100.0, gradnorm_tol=1e-2, acceleration=0, restart_interval=50,
    cost_type=0, gnc_barc=5.0, gnc_mu_step=2.0, gnc_init_mu=1e-5, robust_opt_num_weight_updates=4,
    robust_opt_num_resets=0, robust_opt_inner_iters=30, robust_opt_min_convergence_ratio=0.0, max_num_iters=1000,
    rel_change_tol=0.1, precond_lambda=0.1)


def make_params(**kw) -> OrcParams:
    vals = dict(DEFAULTS)
    vals.update(kw)
    return OrcParams(**vals)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("dpgo_oracle.cpp", "oracle_capi.cpp", "dpgo_oracle.hpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def build_native() -> str | None:
    """liboracle_native.so: the same sources with the reference's own flags (CMakeLists.txt:9: -O3 -march=native), built
    ON THE MACHINE THAT RUNS IT -- what bench.py's CPU arms time (liboracle.so is pinned to x86-64-v3 so that one
    binary runs both in the build container and on the GPU box).  None when it cannot be built here."""
    so = os.path.join(_HERE, "liboracle_native.so")
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                              timeout=300)
        C.CDLL(so)
        return so
    except Exception:  # noqa: BLE001
        return None


def use_native() -> bool:
    """Switch this process to liboracle_native.so (before the first oracle call).  Returns whether it worked."""
    global _LIB, _NATIVE
    if _LIB is not None:
        return _NATIVE
    so = build_native()
    if so is None:
        return False
    os.environ["DPGO_ORACLE_LIB"] = so
    _NATIVE = True
    return True


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.environ.get("DPGO_ORACLE_LIB") or build()
    try:
        L = C.CDLL(so)
    except OSError:
        so = build(force=True)
        L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    L.orc_team_create.restype = C.c_void_p
    L.orc_team_create.argtypes = [C.POINTER(OrcParams)]
    L.orc_team_destroy.argtypes = [C.c_void_p]
    L.orc_add_measurements.argtypes = [C.c_void_p, C.c_int, C.c_int, ip, ip, ip, ip, dp, dp, dp, dp, dp,
                                       C.POINTER(C.c_ubyte)]
    L.orc_num_poses.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_lifting_matrix.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_initialize.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_initialize_in_global_frame.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_exchange_all.argtypes = [C.c_void_p]
    L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(OrcRunResult)]
    L.orc_initialize_chordal.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_iteration_number.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_set_robot_active.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_get_local_trajectory.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.orc_run_parallel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(OrcRunResult)]
    L.orc_agent_iterate.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_get_x.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    L.orc_get_opt_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(OrcOptResult)]
    L.orc_get_status.argtypes = [C.c_void_p, C.c_int, C.POINTER(OrcStatus)]
    L.orc_global_cost.restype = C.c_double
    L.orc_global_cost.argtypes = [C.c_void_p]
    L.orc_weight_update_count.argtypes = [C.c_void_p, C.c_int]
    L.orc_robust_mu.restype = C.c_double
    L.orc_robust_mu.argtypes = [C.c_void_p, C.c_int]
    L.orc_get_lc_weights.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    L.orc_eval.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp]
    L.orc_hess.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
    L.orc_precond.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
    L.orc_dense_q.argtypes = [C.c_void_p, C.c_int, dp, dp]
    L.orc_set_measurement_weight.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double, C.c_int]
    L.orc_compute_measurement_residual.argtypes = [C.c_void_p] + [C.c_int] * 5 + [dp]
    L.orc_manifold_project.argtypes = [C.c_int, C.c_int, dp, dp]
    L.orc_tangent_project.argtypes = [C.c_int, C.c_int, dp, dp, dp]
    L.orc_retract.argtypes = [C.c_int, C.c_int, dp, dp, dp]
    L.orc_robust_weight.restype = C.c_double
    L.orc_robust_weight.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    _LIB = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _chk(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"oracle: {what} failed (rc={rc})")


class OracleTeam:
    """All agents of one problem, driven in-process with the wrapper's synchronous schedule."""

    def __init__(self, problem, ylift: np.ndarray | None = None, initialize: bool = True, **params):
        from dpgo_ros_b200 import datasets  # loaders only (host-side numpy, no CUDA)

        self.L = lib()
        self.problem = problem
        params = dict(params)
        params["num_robots"] = problem.num_robots
        self.params = make_params(**params)
        self.r = self.params.r
        self.h = self.L.orc_team_create(C.byref(self.params))
        if not self.h:
            raise RuntimeError("oracle: team_create failed")
        self.n = []
        for rid in range(problem.num_robots):
            m = problem.robot_measurements(rid)
            self.add_measurements(rid, m)
            self.n.append(self.L.orc_num_poses(self.h, rid))
        if initialize:
            yl = ylift if ylift is not None else datasets.fixed_lifting_matrix(self.r)
            eye = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)
            for rid in range(problem.num_robots):
                self.set_lifting_matrix(rid, yl)
                self.initialize(rid, problem.T_init[rid])
                self.initialize_in_global_frame(rid, eye)
            self.exchange_all()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.orc_team_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_measurements(self, rid, m):
        r1, p1, r2, p2 = (np.ascontiguousarray(x, dtype=np.int32) for x in (m.r1, m.p1, m.r2, m.p2))
        R, t, ka, ta, w = _f64(m.R), _f64(m.t), _f64(m.kappa), _f64(m.tau), _f64(m.weight)
        fx = np.ascontiguousarray(m.fixed, dtype=np.uint8)
        _chk(self.L.orc_add_measurements(self.h, rid, len(m), _ip(r1), _ip(p1), _ip(r2), _ip(p2), _dp(R), _dp(t),
                                         _dp(ka), _dp(ta), _dp(w), fx.ctypes.data_as(C.POINTER(C.c_ubyte))),
             "add_measurements")

    def set_lifting_matrix(self, rid, ylift):
        y = np.asfortranarray(ylift, dtype=np.float64)
        _chk(self.L.orc_set_lifting_matrix(self.h, rid, _dp(y)), "set_lifting_matrix")

    def initialize(self, rid, T_local=None):
        if T_local is None:
            _chk(self.L.orc_initialize(self.h, rid, None), "initialize")
        else:
            T = _f64(T_local)
            _chk(self.L.orc_initialize(self.h, rid, _dp(T)), "initialize")

    def set_iteration_number(self, rid, it):
        _chk(self.L.orc_set_iteration_number(self.h, rid, int(it)), "set_iteration_number")

    def set_robot_active(self, rid, robot, active):
        """setRobotActive on agent `rid` (src/PGOAgentROS.cpp:382 ... :1582)."""
        _chk(self.L.orc_set_robot_active(self.h, rid, int(robot), 1 if active else 0), "set_robot_active")

    def initialize_chordal(self, rid):
        """Chordal local initialisation (Agent::initializeChordal); returns the local trajectory [n, 3, 4]."""
        _chk(self.L.orc_initialize_chordal(self.h, rid), "initialize_chordal")
        return self.local_trajectory(rid)

    def local_trajectory(self, rid) -> np.ndarray:
        out = np.zeros((self.n[rid], 3, 4))
        _chk(self.L.orc_get_local_trajectory(self.h, rid, _dp(out)), "local_trajectory")
        return out

    def initialize_in_global_frame(self, rid, Tw):
        T = _f64(Tw)
        _chk(self.L.orc_initialize_in_global_frame(self.h, rid, _dp(T)), "initialize_in_global_frame")

    def exchange_all(self):
        _chk(self.L.orc_exchange_all(self.h), "exchange_all")

    def run(self, max_iters: int, threads: int = 1, stop_on_terminate: bool = True) -> OrcRunResult:
        out = OrcRunResult()
        _chk(self.L.orc_run(self.h, max_iters, threads, int(stop_on_terminate), C.byref(out)), "run")
        return out

    def run_parallel(self, ticks: int, threads: int = 1) -> OrcRunResult:
        """The asynchronous mode as its equal-rate / unit-delay schedule (Team::runParallel)."""
        out = OrcRunResult()
        _chk(self.L.orc_run_parallel(self.h, ticks, threads, C.byref(out)), "run_parallel")
        return out

    def iterate(self, rid: int, do_opt: bool):
        _chk(self.L.orc_agent_iterate(self.h, rid, int(do_opt)), "iterate")

    def get_x(self, rid: int, which: int = 0) -> np.ndarray:
        """r x 4n (Fortran order). which: 0 X, 1 Y (aux), 2 V."""
        out = np.zeros((self.r, 4 * self.n[rid]), order="F")
        _chk(self.L.orc_get_x(self.h, rid, which, _dp(out)), "get_x")
        return out

    def opt_result(self, rid: int) -> OrcOptResult:
        o = OrcOptResult()
        self.L.orc_get_opt_result(self.h, rid, C.byref(o))
        return o

    def status(self, rid: int) -> OrcStatus:
        s = OrcStatus()
        self.L.orc_get_status(self.h, rid, C.byref(s))
        return s

    def global_cost(self) -> float:
        return float(self.L.orc_global_cost(self.h))

    def lc_weights(self, rid: int) -> np.ndarray:
        buf = np.zeros(1 << 16)
        k = self.L.orc_get_lc_weights(self.h, rid, _dp(buf), buf.size)
        return buf[:k].copy()

    def eval(self, rid: int, X: np.ndarray):
        X = np.asfortranarray(X, dtype=np.float64)
        f = C.c_double()
        eg = np.zeros_like(X, order="F")
        rg = np.zeros_like(X, order="F")
        _chk(self.L.orc_eval(self.h, rid, _dp(X), C.byref(f), _dp(eg), _dp(rg)), "eval")
        return f.value, eg, rg

    def hess(self, rid: int, X: np.ndarray, V: np.ndarray) -> np.ndarray:
        X = np.asfortranarray(X, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        out = np.zeros_like(X, order="F")
        _chk(self.L.orc_hess(self.h, rid, _dp(X), _dp(V), _dp(out)), "hess")
        return out

    def precond(self, rid: int, X: np.ndarray, V: np.ndarray) -> np.ndarray:
        X = np.asfortranarray(X, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        out = np.zeros_like(X, order="F")
        _chk(self.L.orc_precond(self.h, rid, _dp(X), _dp(V), _dp(out)), "precond")
        return out

    def set_measurement_weight(self, rid, r1, p1, r2, p2, w, fixed=False):
        """setMeasurementWeight (src/PGOAgentROS.cpp:1341) followed by clearDataMatrices (:1351)."""
        _chk(self.L.orc_set_measurement_weight(self.h, rid, r1, p1, r2, p2, float(w), int(fixed)), "set_measurement_weight")

    def compute_measurement_residual(self, rid, r1, p1, r2, p2):
        """computeMeasurementResidual (src/PGOAgentROS.cpp:1049); None when a pose is unavailable."""
        res = C.c_double()
        rc = self.L.orc_compute_measurement_residual(self.h, rid, r1, p1, r2, p2, C.byref(res))
        return res.value if rc == 0 else None

    def dense_q(self, rid: int):
        n = self.n[rid]
        Q = np.zeros((4 * n, 4 * n), order="F")
        G = np.zeros((self.r, 4 * n), order="F")
        _chk(self.L.orc_dense_q(self.h, rid, _dp(Q), _dp(G)), "dense_q")
        return Q, G


def manifold_project(M: np.ndarray) -> np.ndarray:
    M = np.asfortranarray(M, dtype=np.float64)
    out = np.zeros_like(M, order="F")
    _chk(lib().orc_manifold_project(M.shape[0], M.shape[1] // 4, _dp(M), _dp(out)), "manifold_project")
    return out


def tangent_project(X: np.ndarray, Z: np.ndarray) -> np.ndarray:
    X = np.asfortranarray(X, dtype=np.float64)
    Z = np.asfortranarray(Z, dtype=np.float64)
    out = np.zeros_like(X, order="F")
    _chk(lib().orc_tangent_project(X.shape[0], X.shape[1] // 4, _dp(X), _dp(Z), _dp(out)), "tangent_project")
    return out


def retract(X: np.ndarray, xi: np.ndarray) -> np.ndarray:
    X = np.asfortranarray(X, dtype=np.float64)
    xi = np.asfortranarray(xi, dtype=np.float64)
    out = np.zeros_like(X, order="F")
    _chk(lib().orc_retract(X.shape[0], X.shape[1] // 4, _dp(X), _dp(xi), _dp(out)), "retract")
    return out


def robust_weight(cost_type: int, barc: float, mu: float, residual: float) -> float:
    return float(lib().orc_robust_weight(cost_type, barc, mu, residual))

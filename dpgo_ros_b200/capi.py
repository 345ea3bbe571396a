"""ctypes binding of the C ABI in include/dpgo_b200.h (libdpgo_b200.so).

The library is the product: hand-written sm_100a kernels behind plain C entry
points.  This module only loads it and declares signatures; it fails loudly if
the shared object is missing (no Python / CPU fallback exists).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdpgo_b200.so")
_LIB = None


class Params(C.Structure):
    """dpgo_b200_params (include/dpgo_b200.h)."""
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


class OptResult(C.Structure):
    _fields_ = [("success", C.c_int), ("f_init", C.c_double), ("f_opt", C.c_double),
                ("gradnorm_init", C.c_double), ("gradnorm_opt", C.c_double), ("relative_change", C.c_double),
                ("rtr_outer_iters", C.c_int), ("tcg_iters", C.c_int), ("rtr_rejections", C.c_int)]


class Status(C.Structure):
    _fields_ = [("agent_id", C.c_int), ("state", C.c_int), ("instance_number", C.c_int),
                ("iteration_number", C.c_int), ("ready_to_terminate", C.c_int), ("relative_change", C.c_double)]


class RunResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("terminated", C.c_int), ("weight_updates", C.c_int),
                ("device_ms", C.c_float), ("kernel_launches", C.c_int), ("stop_reason", C.c_int)]


# defaults of launch/PGOAgent.launch:9-50 (the values the node actually runs with)
DEFAULTS = dict(
    d=3, r=5, num_robots=1, method=0, rgd_stepsize=1e-3, rgd_use_preconditioner=1, rtr_iterations=3,
    rtr_tcg_iterations=50, rtr_initial_radius=100.0, gradnorm_tol=1e-2, acceleration=0, restart_interval=50,
    cost_type=0, gnc_barc=5.0, gnc_mu_step=2.0, gnc_init_mu=1e-5, robust_opt_num_weight_updates=4,
    robust_opt_num_resets=0, robust_opt_inner_iters=30, robust_opt_min_convergence_ratio=0.0, max_num_iters=1000,
    rel_change_tol=0.1, precond_lambda=0.1)

# every symbol include/dpgo_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "dpgo_b200_version", "dpgo_b200_last_error", "dpgo_b200_device_count", "dpgo_b200_kernel_launch_count",
    "dpgo_b200_agent_create", "dpgo_b200_agent_destroy", "dpgo_b200_reset", "dpgo_b200_add_measurements",
    "dpgo_b200_num_poses", "dpgo_b200_iteration_number", "dpgo_b200_num_neighbors", "dpgo_b200_get_neighbors",
    "dpgo_b200_measurement_counts", "dpgo_b200_set_lifting_matrix", "dpgo_b200_get_lifting_matrix",
    "dpgo_b200_initialize", "dpgo_b200_initialize_in_global_frame", "dpgo_b200_iterate",
    "dpgo_b200_get_opt_result", "dpgo_b200_get_status", "dpgo_b200_set_neighbor_status",
    "dpgo_b200_should_terminate", "dpgo_b200_should_update_measurement_weights", "dpgo_b200_get_x",
    "dpgo_b200_set_x", "dpgo_b200_num_shared_poses", "dpgo_b200_get_shared_pose_dict",
    "dpgo_b200_update_neighbor_poses", "dpgo_b200_outbox_device_ptr", "dpgo_b200_inbox_device_ptr",
    "dpgo_b200_mark_inbox_updated", "dpgo_b200_update_measurement_weights", "dpgo_b200_set_measurement_weight",
    "dpgo_b200_compute_measurement_residual", "dpgo_b200_robust_weight", "dpgo_b200_clear_data_matrices",
    "dpgo_b200_get_lc_weights", "dpgo_b200_weight_update_count", "dpgo_b200_eval", "dpgo_b200_hess",
    "dpgo_b200_precond", "dpgo_b200_manifold_project", "dpgo_b200_tangent_project", "dpgo_b200_retract",
    "dpgo_b200_team_create", "dpgo_b200_team_destroy", "dpgo_b200_team_add_agent",
    "dpgo_b200_team_exchange_all", "dpgo_b200_team_run", "dpgo_b200_team_global_cost", "dpgo_b200_team_set_grid",
    "dpgo_b200_sync_driver_run", "dpgo_b200_team_step",
    "dpgo_b200_team_fabric_init", "dpgo_b200_team_fabric_window", "dpgo_b200_team_fabric_import",
    "dpgo_b200_team_fabric_route", "dpgo_b200_team_fabric_run", "dpgo_b200_team_fabric_set_timeout",
    "dpgo_b200_team_fabric_close", "dpgo_b200_team_gnc_compute_weights", "dpgo_b200_team_gnc_finish_update",
    "dpgo_b200_get_shared_loop_closures", "dpgo_b200_team_set_schedule", "dpgo_b200_get_opt_result_lazy",
    "dpgo_b200_get_pose", "dpgo_b200_sync_driver_shm_bytes", "dpgo_b200_sync_driver_run_shm",
    "dpgo_b200_initialize_chordal", "dpgo_b200_get_local_trajectory", "dpgo_b200_set_iteration_number", "dpgo_b200_set_robot_active",
    "dpgo_b200_debug_dense_q", "dpgo_b200_debug_api_profile", "dpgo_b200_debug_edge_grad",
    "dpgo_b200_debug_spd_inverse",
]


def make_params(**kw) -> Params:
    vals = dict(DEFAULTS)
    vals.update(kw)
    return Params(**vals)


def build(force: bool = False) -> str:
    """Compile the CUDA extension for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j8", "-s"] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libdpgo_b200.so was not produced")
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the RBCD path)")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.dpgo_b200_version.restype = C.c_char_p
    L.dpgo_b200_last_error.restype = C.c_char_p
    L.dpgo_b200_kernel_launch_count.restype = C.c_longlong
    L.dpgo_b200_agent_create.argtypes = [C.c_int, C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.dpgo_b200_agent_destroy.argtypes = [vp]
    L.dpgo_b200_reset.argtypes = [vp]
    L.dpgo_b200_add_measurements.argtypes = [vp, C.c_int, ip, ip, ip, ip, dp, dp, dp, dp, dp, C.POINTER(C.c_ubyte)]
    for name in ("num_poses", "iteration_number", "num_neighbors", "should_terminate",
                 "should_update_measurement_weights", "update_measurement_weights", "clear_data_matrices",
                 "weight_update_count"):
        getattr(L, "dpgo_b200_" + name).argtypes = [vp]
    L.dpgo_b200_get_neighbors.argtypes = [vp, ip, C.c_int]
    L.dpgo_b200_set_iteration_number.argtypes = [vp, C.c_int]
    L.dpgo_b200_set_robot_active.argtypes = [vp, C.c_int, C.c_int]
    L.dpgo_b200_measurement_counts.argtypes = [vp, ip, ip, ip]
    L.dpgo_b200_set_lifting_matrix.argtypes = [vp, dp]
    L.dpgo_b200_get_lifting_matrix.argtypes = [vp, dp]
    L.dpgo_b200_initialize.argtypes = [vp, dp]
    L.dpgo_b200_initialize_in_global_frame.argtypes = [vp, dp]
    L.dpgo_b200_initialize_chordal.argtypes = [vp]
    L.dpgo_b200_get_local_trajectory.argtypes = [vp, dp]
    L.dpgo_b200_iterate.argtypes = [vp, C.c_int]
    L.dpgo_b200_get_opt_result.argtypes = [vp, C.POINTER(OptResult)]
    L.dpgo_b200_get_status.argtypes = [vp, C.POINTER(Status)]
    L.dpgo_b200_set_neighbor_status.argtypes = [vp, C.POINTER(Status)]
    L.dpgo_b200_get_x.argtypes = [vp, C.c_int, dp]
    L.dpgo_b200_get_pose.argtypes = [vp, C.c_int, C.c_int, dp]
    L.dpgo_b200_get_opt_result_lazy.argtypes = [vp, C.POINTER(OptResult)]
    L.dpgo_b200_set_x.argtypes = [vp, dp]
    L.dpgo_b200_num_shared_poses.argtypes = [vp, C.c_int]
    L.dpgo_b200_get_shared_pose_dict.argtypes = [vp, C.c_int, C.c_int, ip, dp, C.c_int, ip]
    L.dpgo_b200_update_neighbor_poses.argtypes = [vp, C.c_int, C.c_int, ip, dp, C.c_int]
    L.dpgo_b200_outbox_device_ptr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.dpgo_b200_inbox_device_ptr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.dpgo_b200_mark_inbox_updated.argtypes = [vp, C.c_int, C.c_int]
    L.dpgo_b200_set_measurement_weight.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    L.dpgo_b200_compute_measurement_residual.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.dpgo_b200_robust_weight.restype = C.c_double
    L.dpgo_b200_robust_weight.argtypes = [vp, C.c_double]
    L.dpgo_b200_get_lc_weights.argtypes = [vp, dp, C.c_int]
    L.dpgo_b200_get_shared_loop_closures.argtypes = [vp, ip, ip, ip, ip, dp, C.POINTER(C.c_ubyte), C.c_int]
    L.dpgo_b200_debug_dense_q.argtypes = [vp, dp, dp]
    L.dpgo_b200_debug_edge_grad.argtypes = [vp, dp, C.c_int, dp, dp, dp, dp]
    L.dpgo_b200_debug_spd_inverse.argtypes = [C.c_int, C.c_int, dp, dp, dp]
    L.dpgo_b200_debug_api_profile.argtypes = [C.c_double, C.c_double, C.c_char_p, C.c_int, C.c_int]
    L.dpgo_b200_eval.argtypes = [vp, dp, dp, dp, dp]
    L.dpgo_b200_hess.argtypes = [vp, dp, dp, dp]
    L.dpgo_b200_precond.argtypes = [vp, dp, dp, dp]
    L.dpgo_b200_manifold_project.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp]
    L.dpgo_b200_tangent_project.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.dpgo_b200_retract.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.dpgo_b200_team_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.dpgo_b200_team_destroy.argtypes = [vp]
    L.dpgo_b200_team_add_agent.argtypes = [vp, vp]
    L.dpgo_b200_team_exchange_all.argtypes = [vp]
    L.dpgo_b200_team_run.argtypes = [vp, C.c_int, C.c_int, C.POINTER(RunResult)]
    L.dpgo_b200_team_global_cost.restype = C.c_double
    L.dpgo_b200_team_global_cost.argtypes = [vp, ip]
    L.dpgo_b200_team_set_grid.argtypes = [vp, C.c_int]
    L.dpgo_b200_team_set_schedule.argtypes = [vp, C.c_int]
    L.dpgo_b200_team_step.argtypes = [vp, C.c_int, C.c_int]
    L.dpgo_b200_team_fabric_init.argtypes = [vp, C.c_int, C.c_int]
    L.dpgo_b200_team_fabric_window.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), vp]
    L.dpgo_b200_team_fabric_import.argtypes = [vp, C.c_int, vp, vp]
    L.dpgo_b200_team_fabric_route.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t]
    L.dpgo_b200_team_fabric_run.argtypes = [vp, C.c_int, C.c_int, C.POINTER(RunResult)]
    L.dpgo_b200_team_fabric_set_timeout.argtypes = [vp, C.c_double]
    L.dpgo_b200_team_fabric_close.argtypes = [vp]
    L.dpgo_b200_team_gnc_compute_weights.argtypes = [vp]
    L.dpgo_b200_team_gnc_finish_update.argtypes = [vp]
    L.dpgo_b200_sync_driver_shm_bytes.restype = C.c_size_t
    L.dpgo_b200_sync_driver_shm_bytes.argtypes = [C.c_int, C.c_int]
    L.dpgo_b200_sync_driver_run_shm.argtypes = [C.POINTER(vp), ip, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                                                C.c_int, dp, ip]
    L.dpgo_b200_sync_driver_run.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, dp, C.POINTER(C.c_longlong), ip]
    _LIB = L
    return L


class DpgoError(RuntimeError):
    def __init__(self, code: int, what: str):
        msg = lib().dpgo_b200_last_error()
        super().__init__(f"{what} failed (code {code}): {msg.decode() if msg else ''}")
        self.code = code


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise DpgoError(rc, what)

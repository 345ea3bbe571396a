"""Host-side mirror of the DPGO::PGOAgent call surface over the C ABI.

Method names follow the reference API that `PGOAgentROS` drives (SURVEY
App. A; call sites in src/PGOAgentROS.cpp): addMeasurement, setLiftingMatrix,
initialize, initializeInGlobalFrame, iterate, getSharedPoseDictWithNeighbor,
updateNeighborPoses, updateMeasurementWeights, shouldTerminate, ... plus the
north-star aliases getX / setNeighborPoses.  All arithmetic happens in
libdpgo_b200.so (sm_100a kernels); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import capi
from .capi import OptResult, Params, RunResult, Status, check, make_params


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class PGOAgent:
    """One robot's RBCD agent on one CUDA device (DPGO::PGOAgent(ID, params), src/PGOAgentROS.cpp:26)."""

    def __init__(self, agent_id: int, params: Params, device: int = 0):
        self.L = capi.lib()
        self.id = agent_id
        self.params = params
        self.r = params.r
        self.device = device
        h = C.c_void_p()
        check(self.L.dpgo_b200_agent_create(agent_id, C.byref(params), device, C.byref(h)), "agent_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.dpgo_b200_agent_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pose graph ------------------------------------------------------------------
    def addMeasurements(self, m) -> None:
        """addMeasurement for a whole SoA (src/PGOAgentROS.cpp:277,1307)."""
        r1, p1, r2, p2 = (np.ascontiguousarray(x, dtype=np.int32) for x in (m.r1, m.p1, m.r2, m.p2))
        R, t, ka, ta, w = _f64(m.R), _f64(m.t), _f64(m.kappa), _f64(m.tau), _f64(m.weight)
        fx = np.ascontiguousarray(m.fixed, dtype=np.uint8)
        check(self.L.dpgo_b200_add_measurements(self.h, len(m), _ip(r1), _ip(p1), _ip(r2), _ip(p2), _dp(R), _dp(t),
                                                _dp(ka), _dp(ta), _dp(w), fx.ctypes.data_as(C.POINTER(C.c_ubyte))),
              "addMeasurement")

    def num_poses(self) -> int:
        return self.L.dpgo_b200_num_poses(self.h)

    def iteration_number(self) -> int:
        return self.L.dpgo_b200_iteration_number(self.h)

    def getNeighbors(self) -> List[int]:
        k = self.L.dpgo_b200_num_neighbors(self.h)
        buf = np.zeros(max(k, 1), dtype=np.int32)
        check(self.L.dpgo_b200_get_neighbors(self.h, _ip(buf), k), "getNeighbors")
        return [int(x) for x in buf[:k]]

    def measurementCounts(self) -> Tuple[int, int, int]:
        o, p, s = C.c_int(), C.c_int(), C.c_int()
        check(self.L.dpgo_b200_measurement_counts(self.h, C.byref(o), C.byref(p), C.byref(s)), "measurementCounts")
        return o.value, p.value, s.value

    # ---- lifecycle ---------------------------------------------------------------------
    def setLiftingMatrix(self, ylift: np.ndarray) -> None:
        y = np.asfortranarray(ylift, dtype=np.float64)
        check(self.L.dpgo_b200_set_lifting_matrix(self.h, _dp(y)), "setLiftingMatrix")

    def getLiftingMatrix(self) -> np.ndarray:
        y = np.zeros((self.r, 3), order="F")
        check(self.L.dpgo_b200_get_lifting_matrix(self.h, _dp(y)), "getLiftingMatrix")
        return y

    def initialize(self, T_local: Optional[np.ndarray] = None) -> None:
        if T_local is None:
            check(self.L.dpgo_b200_initialize(self.h, None), "initialize")
        else:
            T = _f64(T_local)
            check(self.L.dpgo_b200_initialize(self.h, _dp(T)), "initialize")

    def initializeInGlobalFrame(self, T_world_robot: np.ndarray) -> None:
        T = _f64(T_world_robot)
        check(self.L.dpgo_b200_initialize_in_global_frame(self.h, _dp(T)), "initializeInGlobalFrame")

    def reset(self) -> None:
        check(self.L.dpgo_b200_reset(self.h), "reset")

    # ---- hot path ------------------------------------------------------------------------
    def iterate(self, doOptimization: bool = True) -> bool:
        """PGOAgent::iterate: True when the local solve ran (or none was asked for); False when iterate(true) had to
        skip it because a neighbour's poses have not arrived (the iteration is counted all the same)."""
        rc = self.L.dpgo_b200_iterate(self.h, int(doOptimization))
        if rc == -4:  # DPGO_B200_ERR_MISSING
            return False
        check(rc, "iterate")
        return True

    def setIterationNumber(self, iteration: int) -> None:
        """What the RECOVER handler does to mIterationNumber (src/PGOAgentROS.cpp:1196)."""
        check(self.L.dpgo_b200_set_iteration_number(self.h, int(iteration)), "setIterationNumber")

    def setRobotActive(self, robot: int, active: bool) -> None:
        """setRobotActive (src/PGOAgentROS.cpp:382 ... :1582): a deactivated robot no longer counts in shouldTerminate."""
        check(self.L.dpgo_b200_set_robot_active(self.h, int(robot), 1 if active else 0), "setRobotActive")

    def initializeChordal(self) -> np.ndarray:
        """Chordal local initialisation on the device; returns the local trajectory [n, 3, 4]."""
        check(self.L.dpgo_b200_initialize_chordal(self.h), "initializeChordal")
        return self.localTrajectory()

    def localTrajectory(self) -> np.ndarray:
        out = np.zeros((self.num_poses(), 3, 4))
        check(self.L.dpgo_b200_get_local_trajectory(self.h, _dp(out)), "localTrajectory")
        return out

    def getX(self, which: int = 0) -> np.ndarray:
        out = np.zeros((self.r, 4 * self.num_poses()), order="F")
        check(self.L.dpgo_b200_get_x(self.h, which, _dp(out)), "getX")
        return out

    def setX(self, X: np.ndarray) -> None:
        X = np.asfortranarray(X, dtype=np.float64)
        check(self.L.dpgo_b200_set_x(self.h, _dp(X)), "setX")

    def localOptResult(self) -> OptResult:
        o = OptResult()
        check(self.L.dpgo_b200_get_opt_result(self.h, C.byref(o)), "getOptResult")
        return o

    def getStatus(self) -> Status:
        s = Status()
        check(self.L.dpgo_b200_get_status(self.h, C.byref(s)), "getStatus")
        return s

    def setNeighborStatus(self, s: Status) -> None:
        check(self.L.dpgo_b200_set_neighbor_status(self.h, C.byref(s)), "setNeighborStatus")

    def shouldTerminate(self) -> bool:
        rc = self.L.dpgo_b200_should_terminate(self.h)
        if rc < 0:
            check(rc, "shouldTerminate")
        return bool(rc)

    def shouldUpdateMeasurementWeights(self) -> bool:
        rc = self.L.dpgo_b200_should_update_measurement_weights(self.h)
        if rc < 0:
            check(rc, "shouldUpdateMeasurementWeights")
        return bool(rc)

    # ---- public poses with host buffers (a9) ---------------------------------------------------
    def getSharedPoseDictWithNeighbor(self, neighbor: int, aux: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        cap = max(1, self.L.dpgo_b200_num_shared_poses(self.h, neighbor))
        frames = np.zeros(cap, dtype=np.int32)
        poses = np.zeros((cap, 4, self.r))  # each pose r x 4 column-major == [4][r] C-order
        cnt = C.c_int()
        check(self.L.dpgo_b200_get_shared_pose_dict(self.h, neighbor, int(aux), _ip(frames), _dp(poses), cap,
                                                    C.byref(cnt)), "getSharedPoseDictWithNeighbor")
        return frames[:cnt.value], poses[:cnt.value]

    def getAuxSharedPoseDictWithNeighbor(self, neighbor: int):
        return self.getSharedPoseDictWithNeighbor(neighbor, aux=True)

    def updateNeighborPoses(self, neighbor: int, frames: np.ndarray, poses: np.ndarray, aux: bool = False) -> None:
        frames = np.ascontiguousarray(frames, dtype=np.int32)
        poses = _f64(poses)
        check(self.L.dpgo_b200_update_neighbor_poses(self.h, neighbor, int(aux), _ip(frames), _dp(poses),
                                                     len(frames)), "updateNeighborPoses")

    def updateAuxNeighborPoses(self, neighbor: int, frames, poses) -> None:
        self.updateNeighborPoses(neighbor, frames, poses, aux=True)

    setNeighborPoses = updateNeighborPoses  # north-star alias

    def outboxDevicePtr(self, neighbor: int, aux: bool = False) -> Tuple[int, int]:
        p, nbytes = C.c_void_p(), C.c_size_t()
        check(self.L.dpgo_b200_outbox_device_ptr(self.h, neighbor, int(aux), C.byref(p), C.byref(nbytes)),
              "outboxDevicePtr")
        return int(p.value or 0), int(nbytes.value)

    def inboxDevicePtr(self, neighbor: int, aux: bool = False) -> Tuple[int, int]:
        p, nbytes = C.c_void_p(), C.c_size_t()
        check(self.L.dpgo_b200_inbox_device_ptr(self.h, neighbor, int(aux), C.byref(p), C.byref(nbytes)),
              "inboxDevicePtr")
        return int(p.value or 0), int(nbytes.value)

    def markInboxUpdated(self, neighbor: int, aux: bool = False) -> None:
        check(self.L.dpgo_b200_mark_inbox_updated(self.h, neighbor, int(aux)), "markInboxUpdated")

    # ---- GNC ---------------------------------------------------------------------------------
    def updateMeasurementWeights(self) -> None:
        check(self.L.dpgo_b200_update_measurement_weights(self.h), "updateMeasurementWeights")

    def setMeasurementWeight(self, r1, p1, r2, p2, w, fixed=False) -> bool:
        return self.L.dpgo_b200_set_measurement_weight(self.h, r1, p1, r2, p2, float(w), int(fixed)) == 0

    def computeMeasurementResidual(self, r1, p1, r2, p2) -> Optional[float]:
        res = C.c_double()
        rc = self.L.dpgo_b200_compute_measurement_residual(self.h, r1, p1, r2, p2, C.byref(res))
        return res.value if rc == 0 else None

    def robustWeight(self, residual: float) -> float:
        return float(self.L.dpgo_b200_robust_weight(self.h, float(residual)))

    def clearDataMatrices(self) -> None:
        check(self.L.dpgo_b200_clear_data_matrices(self.h), "clearDataMatrices")

    def lcWeights(self) -> np.ndarray:
        buf = np.zeros(1 << 16)
        k = self.L.dpgo_b200_get_lc_weights(self.h, _dp(buf), buf.size)
        return buf[:k].copy()

    def sharedLoopClosures(self):
        """(r1, p1, r2, p2, weight, fixed) arrays of every shared loop closure (publishMeasurementWeights, :721-754)."""
        k = self.L.dpgo_b200_get_shared_loop_closures(self.h, None, None, None, None, None, None, 0)
        if k < 0:
            check(k, "sharedLoopClosures")
        ids = [np.zeros(k, dtype=np.int32) for _ in range(4)]
        w = np.zeros(k)
        fx = np.zeros(k, dtype=np.uint8)
        self.L.dpgo_b200_get_shared_loop_closures(self.h, _ip(ids[0]), _ip(ids[1]), _ip(ids[2]), _ip(ids[3]), _dp(w),
                                                  fx.ctypes.data_as(C.POINTER(C.c_ubyte)), k)
        return ids[0], ids[1], ids[2], ids[3], w, fx

    def weightUpdateCount(self) -> int:
        return self.L.dpgo_b200_weight_update_count(self.h)

    # ---- parity hooks -----------------------------------------------------------------------------
    def denseQ(self):
        """Q as the device kernel assembled it: (from the block-CSR copy, from the ELL + overflow copy), 4n x 4n."""
        n4 = 4 * self.num_poses()
        a = np.zeros((n4, n4), order="F")
        b = np.zeros((n4, n4), order="F")
        check(self.L.dpgo_b200_debug_dense_q(self.h, _dp(a), _dp(b)), "denseQ")
        return a, b

    def edgeGrad(self, X: Optional[np.ndarray] = None, flush_l2: bool = False):
        """k_edge_grad on its own (LARGE agents): (f, rgrad, kernel ns by globaltimer, ns by CUDA events)."""
        n4 = 4 * self.num_poses()
        rg = np.zeros((self.r, n4), order="F")
        f, kns, ens = C.c_double(), C.c_double(), C.c_double()
        Xp = None
        if X is not None:
            X = np.asfortranarray(X, dtype=np.float64)
            Xp = _dp(X)
        check(self.L.dpgo_b200_debug_edge_grad(self.h, Xp, int(flush_l2), C.byref(f), _dp(rg), C.byref(kns), C.byref(ens)),
              "edgeGrad")
        return f.value, rg, kns.value, ens.value

    def eval(self, X: np.ndarray):
        X = np.asfortranarray(X, dtype=np.float64)
        f = C.c_double()
        eg = np.zeros_like(X, order="F")
        rg = np.zeros_like(X, order="F")
        check(self.L.dpgo_b200_eval(self.h, _dp(X), C.byref(f), _dp(eg), _dp(rg)), "eval")
        return f.value, eg, rg

    def hess(self, X: np.ndarray, V: np.ndarray) -> np.ndarray:
        X = np.asfortranarray(X, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        out = np.zeros_like(X, order="F")
        check(self.L.dpgo_b200_hess(self.h, _dp(X), _dp(V), _dp(out)), "hess")
        return out

    def precond(self, X: np.ndarray, V: np.ndarray) -> np.ndarray:
        X = np.asfortranarray(X, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        out = np.zeros_like(X, order="F")
        check(self.L.dpgo_b200_precond(self.h, _dp(X), _dp(V), _dp(out)), "precond")
        return out


class Team:
    """Co-located agents on one device; the persistent kernel runs the whole
    synchronous schedule on the device (dpgo_b200_team_run)."""

    def __init__(self, device: int = 0):
        self.L = capi.lib()
        h = C.c_void_p()
        check(self.L.dpgo_b200_team_create(device, C.byref(h)), "team_create")
        self.h = h
        self.agents: List[PGOAgent] = []

    def close(self):
        if getattr(self, "h", None):
            self.L.dpgo_b200_team_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, agent: PGOAgent) -> None:
        check(self.L.dpgo_b200_team_add_agent(self.h, agent.h), "team_add_agent")
        self.agents.append(agent)

    def exchange_all(self) -> None:
        check(self.L.dpgo_b200_team_exchange_all(self.h), "team_exchange_all")

    def run(self, max_iters: int, stop_on_terminate: bool = True) -> RunResult:
        out = RunResult()
        check(self.L.dpgo_b200_team_run(self.h, max_iters, int(stop_on_terminate), C.byref(out)), "team_run")
        return out

    def global_cost(self) -> float:
        st = C.c_int()
        c = self.L.dpgo_b200_team_global_cost(self.h, C.byref(st))
        check(st.value, "team_global_cost")
        return float(c)

    def step(self, selected_robot: int, mode: int = 0) -> None:
        """One global iteration for the local agents of a partial team (multi-GPU); see dpgo_b200_team_step."""
        check(self.L.dpgo_b200_team_step(self.h, selected_robot, mode), "team_step")

    def set_schedule(self, schedule: int) -> None:
        """0: synchronous RoundRobin token; 1: parallel ticks (the asynchronous mode, see dpgo_b200_team_set_schedule)."""
        check(self.L.dpgo_b200_team_set_schedule(self.h, schedule), "team_set_schedule")

    def set_grid(self, num_ctas: int) -> None:
        check(self.L.dpgo_b200_team_set_grid(self.h, num_ctas), "team_set_grid")

    # ---- multi-GPU fabric (include/dpgo_b200.h: "multi-GPU fabric") ---------------------------------
    def fabric_init(self, world: int, rank: int) -> None:
        check(self.L.dpgo_b200_team_fabric_init(self.h, world, rank), "team_fabric_init")

    def fabric_window(self, want_handle: bool = True) -> Tuple[int, int, bytes]:
        """(base pointer, bytes, 64-byte CUDA IPC handle) of this rank's window."""
        base, nbytes = C.c_void_p(), C.c_size_t()
        hbuf = C.create_string_buffer(64) if want_handle else None
        check(self.L.dpgo_b200_team_fabric_window(self.h, C.byref(base), C.byref(nbytes), hbuf), "team_fabric_window")
        return int(base.value or 0), int(nbytes.value), (hbuf.raw if want_handle else b"")

    def fabric_import(self, peer_rank: int, ipc_handle: Optional[bytes] = None, base: Optional[int] = None) -> None:
        hb = C.create_string_buffer(ipc_handle, 64) if ipc_handle else None
        check(self.L.dpgo_b200_team_fabric_import(self.h, peer_rank, hb, C.c_void_p(base) if base else None),
              "team_fabric_import")

    def fabric_route(self, robot: int, neighbor: int, peer_rank: int, off_reg: int, off_aux: int) -> None:
        check(self.L.dpgo_b200_team_fabric_route(self.h, robot, neighbor, peer_rank, off_reg, off_aux),
              "team_fabric_route")

    def fabric_run(self, max_iters: int, stop_on_terminate: bool = True) -> RunResult:
        out = RunResult()
        check(self.L.dpgo_b200_team_fabric_run(self.h, max_iters, int(stop_on_terminate), C.byref(out)),
              "team_fabric_run")
        return out

    def fabric_set_timeout(self, seconds: float) -> None:
        check(self.L.dpgo_b200_team_fabric_set_timeout(self.h, float(seconds)), "team_fabric_set_timeout")

    def fabric_close(self) -> None:
        check(self.L.dpgo_b200_team_fabric_close(self.h), "team_fabric_close")

    def gnc_compute_weights(self) -> None:
        check(self.L.dpgo_b200_team_gnc_compute_weights(self.h), "team_gnc_compute_weights")

    def gnc_finish_update(self) -> None:
        check(self.L.dpgo_b200_team_gnc_finish_update(self.h), "team_gnc_finish_update")


def make_team(problem, ylift: Optional[np.ndarray] = None, device: int = 0, colocate: bool = True, **params):
    """All robots of `problem` as initialised agents (odometry guess lifted by a fixed
    YLift) -- either co-located in one Team or as standalone agents."""
    from . import datasets

    params = dict(params)
    params["num_robots"] = problem.num_robots
    P = make_params(**params)
    yl = ylift if ylift is not None else datasets.fixed_lifting_matrix(P.r)
    eye = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)
    agents = []
    for rid in range(problem.num_robots):
        ag = PGOAgent(rid, P, device)
        ag.addMeasurements(problem.robot_measurements(rid))
        ag.setLiftingMatrix(yl)
        ag.initialize(problem.T_init[rid])
        ag.initializeInGlobalFrame(eye)
        agents.append(ag)
    if not colocate:
        return None, agents
    team = Team(device)
    for ag in agents:
        team.add(ag)
    team.exchange_all()
    return team, agents


def exchange_host(agents: List[PGOAgent], accel: bool, only: Optional[List[int]] = None) -> int:
    """publishPublicPoses / publicPosesCallback through HOST buffers for standalone
    agents (src/PGOAgentROS.cpp:662-690, 1255-1284).  Returns bytes moved D2H+H2D."""
    moved = 0
    by_id: Dict[int, PGOAgent] = {a.id: a for a in agents}
    for a in agents:
        if only is not None and a.id not in only:
            continue
        for nb in a.getNeighbors():
            if nb not in by_id:
                continue
            fr, poses = a.getSharedPoseDictWithNeighbor(nb, False)
            by_id[nb].updateNeighborPoses(a.id, fr, poses, False)
            moved += 2 * poses.nbytes
            if accel:
                fr, poses = a.getSharedPoseDictWithNeighbor(nb, True)
                by_id[nb].updateNeighborPoses(a.id, fr, poses, True)
                moved += 2 * poses.nbytes
    return moved


def sync_driver_run(agents: List[PGOAgent], steps: int, accelerated: bool):
    """Native replay of the wrapper's synchronous call sequence through the per-robot C ABI with host
    buffers, one OS thread per robot (dpgo_b200_sync_driver_run).  Returns (seconds, terminated_at)."""
    L = capi.lib()
    arr = (C.c_void_p * len(agents))(*[a.h for a in agents])
    sec = C.c_double()
    nbytes = C.c_longlong()
    term = C.c_int()
    check(L.dpgo_b200_sync_driver_run(arr, len(agents), steps, int(accelerated), C.byref(sec), C.byref(nbytes),
                                      C.byref(term)), "sync_driver_run")
    return sec.value, term.value


def exchange_payload_bytes(agents: List[PGOAgent], accelerated: bool) -> int:
    """Bytes of public poses every robot publishes in one accelerated step (each pose r x 4 FP64)."""
    total = 0
    for a in agents:
        for nb in a.getNeighbors():
            total += a.L.dpgo_b200_num_shared_poses(a.h, nb) * a.r * 4 * 8 * (2 if accelerated else 1)
    return total


def spd_inverse(A: np.ndarray, device: int = 0) -> Tuple[np.ndarray, float]:
    """A^-1 of a symmetric positive definite matrix through the library's own dense inverse
    (`dpgo_b200_debug_spd_inverse`, dpgo_ros_b200/csrc/dense_inverse.cu -- the kernel set behind the preconditioner,
    where the reference factors Q + lambda I with CHOLMOD).  N is padded to a multiple of 32 with identity, as the
    library pads its own matrices.  Returns (inverse, device milliseconds of the factorisation)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    N = (n + 31) // 32 * 32
    Ap = np.eye(N)
    Ap[:n, :n] = A
    Af = np.asfortranarray(Ap)
    P = np.empty((N, N), dtype=np.float64, order="F")
    ms = C.c_double(0.0)
    check(capi.lib().dpgo_b200_debug_spd_inverse(int(device), N, _dp(Af), _dp(P), C.byref(ms)), "debug_spd_inverse")
    return np.ascontiguousarray(P[:n, :n]), ms.value

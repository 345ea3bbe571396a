"""Dataset loaders and the multi-robot partition rule (host side, numpy only).

These feed *identical* inputs to the CPU oracle and to the CUDA agent.  They
mirror what the reference's dataset publisher does before any arithmetic runs:

* g2o parsing with the SE-Sync precision rule (kappa, tau from the information
  matrix) -- the reference calls ``read_g2o_file`` at
  ``src/PGODatasetPublisherNode.cpp:80``;
* contiguous-block partition of the global pose index into robots and the
  odometry / private / shared classification,
  ``src/PGODatasetPublisherNode.cpp:84-134``;
* the per-robot CSV measurement format of the tunnels dataset
  (``data/tunnels/robot0/measurements.csv:1``), loaded at
  ``src/PGODatasetPublisherNode.cpp:168``.

Measurements are held as a struct-of-arrays (`Measurements`) because that is
the form the C ABI takes (`dpgo_b200_add_measurements`).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Tuple

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA_DIR = os.path.join(REPO_ROOT, "data")


@dataclasses.dataclass
class Measurements:
    """SoA of relative SE(3) measurements (i -> j): T_ij = (R, t), precisions."""

    r1: np.ndarray  # int32 [m]  source robot
    p1: np.ndarray  # int32 [m]  source pose (robot-local index)
    r2: np.ndarray  # int32 [m]
    p2: np.ndarray  # int32 [m]
    R: np.ndarray  # float64 [m,3,3]
    t: np.ndarray  # float64 [m,3]
    kappa: np.ndarray  # float64 [m]
    tau: np.ndarray  # float64 [m]
    weight: np.ndarray  # float64 [m]
    fixed: np.ndarray  # uint8 [m]  fixedWeight (known inlier)

    def __len__(self) -> int:
        return int(self.r1.shape[0])

    def take(self, idx) -> "Measurements":
        idx = np.asarray(idx, dtype=np.int64)
        return Measurements(*(getattr(self, f.name)[idx].copy() for f in dataclasses.fields(self)))

    @staticmethod
    def concat(parts: List["Measurements"]) -> "Measurements":
        return Measurements(
            *(np.concatenate([getattr(p, f.name) for p in parts], axis=0) for f in dataclasses.fields(Measurements))
        )

    @staticmethod
    def empty() -> "Measurements":
        z = np.zeros
        return Measurements(z(0, np.int32), z(0, np.int32), z(0, np.int32), z(0, np.int32), z((0, 3, 3)), z((0, 3)),
                            z(0), z(0), z(0), z(0, np.uint8))


def quat_to_rot(q: np.ndarray) -> np.ndarray:
    """Unit quaternion(s) (x, y, z, w) -> rotation matrix/matrices."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def read_g2o(path: str) -> Tuple[Measurements, int]:
    """Parse a 3-D g2o file (EDGE_SE3:QUAT).  Returns (measurements, num_poses).

    All measurements come back with robot id 0 and *global* pose indices, as
    ``read_g2o_file`` does for the reference (``PGODatasetPublisherNode.cpp:80``).
    Precisions follow the SE-Sync rule the dpgo loader uses:
    ``tau = 3 / tr(I_t^-1)``, ``kappa = 3 / (2 tr(I_R^-1))``.
    """
    src, dst, tt, qq, infos = [], [], [], [], []
    with open(path, "r") as fh:
        for line in fh:
            if not line.startswith("EDGE_SE3:QUAT"):
                continue
            tok = line.split()
            src.append(int(tok[1]))
            dst.append(int(tok[2]))
            v = np.array(tok[3:3 + 7 + 21], dtype=np.float64)
            tt.append(v[0:3])
            qq.append(v[3:7])
            infos.append(v[7:28])
    m = len(src)
    src = np.array(src, dtype=np.int64)
    dst = np.array(dst, dtype=np.int64)
    tt = np.array(tt).reshape(m, 3)
    qq = np.array(qq).reshape(m, 4)
    infos = np.array(infos).reshape(m, 21)
    # upper-triangular 6x6, row-major: (0,0..5),(1,1..5),...
    full = np.zeros((m, 6, 6))
    iu = np.triu_indices(6)
    full[:, iu[0], iu[1]] = infos
    full[:, iu[1], iu[0]] = infos
    tran_cov = np.linalg.inv(full[:, 0:3, 0:3])
    rot_cov = np.linalg.inv(full[:, 3:6, 3:6])
    tau = 3.0 / np.trace(tran_cov, axis1=1, axis2=2)
    kappa = 3.0 / (2.0 * np.trace(rot_cov, axis1=1, axis2=2))
    num_poses = int(max(src.max(), dst.max()) + 1) if m else 0
    meas = Measurements(
        r1=np.zeros(m, np.int32), p1=src.astype(np.int32), r2=np.zeros(m, np.int32), p2=dst.astype(np.int32),
        R=quat_to_rot(qq), t=tt, kappa=kappa, tau=tau, weight=np.ones(m), fixed=np.zeros(m, np.uint8))
    return meas, num_poses


def write_g2o(path: str, meas: Measurements, num_poses: int) -> None:
    """Write a g2o file (isotropic information matrices reproducing kappa, tau)."""
    from_rot = rot_to_quat(meas.R)
    with open(path, "w") as fh:
        for i in range(num_poses):
            fh.write(f"VERTEX_SE3:QUAT {i} 0 0 0 0 0 0 1\n")
        for e in range(len(meas)):
            it = meas.tau[e]
            ir = 2.0 * meas.kappa[e]
            info = [it, 0, 0, 0, 0, 0, it, 0, 0, 0, 0, it, 0, 0, 0, ir, 0, 0, ir, 0, ir]
            vals = list(meas.t[e]) + list(from_rot[e]) + info
            fh.write("EDGE_SE3:QUAT %d %d " % (meas.p1[e], meas.p2[e]) + " ".join(repr(float(v)) for v in vals) + "\n")


def rot_to_quat(R: np.ndarray) -> np.ndarray:
    """Rotation matrices [m,3,3] -> quaternions (x,y,z,w) [m,4]."""
    R = np.asarray(R)
    m = R.shape[0]
    q = np.empty((m, 4))
    for e in range(m):
        M = R[e]
        tr = M[0, 0] + M[1, 1] + M[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            q[e] = [(M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s, 0.25 * s]
        elif M[0, 0] > M[1, 1] and M[0, 0] > M[2, 2]:
            s = np.sqrt(1.0 + M[0, 0] - M[1, 1] - M[2, 2]) * 2
            q[e] = [0.25 * s, (M[0, 1] + M[1, 0]) / s, (M[0, 2] + M[2, 0]) / s, (M[2, 1] - M[1, 2]) / s]
        elif M[1, 1] > M[2, 2]:
            s = np.sqrt(1.0 + M[1, 1] - M[0, 0] - M[2, 2]) * 2
            q[e] = [(M[0, 1] + M[1, 0]) / s, 0.25 * s, (M[1, 2] + M[2, 1]) / s, (M[0, 2] - M[2, 0]) / s]
        else:
            s = np.sqrt(1.0 + M[2, 2] - M[0, 0] - M[1, 1]) * 2
            q[e] = [(M[0, 2] + M[2, 0]) / s, (M[1, 2] + M[2, 1]) / s, 0.25 * s, (M[1, 0] - M[0, 1]) / s]
    return q


def partition_contiguous(meas: Measurements, num_poses: int, num_robots: int) -> Tuple[Measurements, np.ndarray]:
    """Map global pose ids to (robot, local id) in contiguous blocks.

    Rule of ``src/PGODatasetPublisherNode.cpp:84-103``: ``n // num_robots`` poses
    per robot, the last robot takes the remainder.  Returns the relabelled
    measurements and ``start[robot]`` (length num_robots+1).  Odometry edges
    (same robot, ``p1 + 1 == p2``) are marked ``fixed`` as the message path does
    (``src/utils.cpp:147-149``).
    """
    per = num_poses // num_robots
    if per <= 0:
        raise ValueError("Number of robots must be smaller than total number of poses")
    start = np.array([r * per for r in range(num_robots)] + [num_poses], dtype=np.int64)
    robot_of = np.minimum(np.arange(num_poses) // per, num_robots - 1).astype(np.int32)
    local_of = (np.arange(num_poses) - start[robot_of]).astype(np.int32)
    out = meas.take(np.arange(len(meas)))
    out.r1 = robot_of[meas.p1]
    out.r2 = robot_of[meas.p2]
    out.p1 = local_of[meas.p1]
    out.p2 = local_of[meas.p2]
    odo = (out.r1 == out.r2) & (out.p1 + 1 == out.p2)
    out.fixed = odo.astype(np.uint8)
    return out, start


def measurements_of_robot(meas: Measurements, robot: int) -> Measurements:
    """Every measurement that involves `robot`, in the order the reference
    inserts them: odometry, private loop closures, then shared loop closures
    (``src/PGODatasetPublisherNode.cpp:137-158``).  Shared loop closures are
    given to *both* end robots, which is the state the reference reaches after
    ``publishPublicMeasurements`` (``src/PGOAgentROS.cpp:692-719,1286-1313``).
    Duplicate (src, dst) pairs are dropped, as ``hasMeasurement`` guards the
    insert at ``src/PGOAgentROS.cpp:276``.
    """
    same = (meas.r1 == robot) & (meas.r2 == robot)
    odo = same & (meas.p1 + 1 == meas.p2)
    plc = same & ~odo
    shared = ((meas.r1 == robot) | (meas.r2 == robot)) & ~same
    order = np.concatenate([np.nonzero(odo)[0], np.nonzero(plc)[0], np.nonzero(shared)[0]])
    seen = set()
    keep = []
    for e in order:
        key = (int(meas.r1[e]), int(meas.p1[e]), int(meas.r2[e]), int(meas.p2[e]))
        if key in seen:
            continue
        seen.add(key)
        keep.append(e)
    return meas.take(keep)


def read_measurements_csv(path: str, load_weight: bool = False) -> Measurements:
    """Per-robot CSV of the tunnels dataset (``PGOLogger::loadMeasurements``).

    Header: robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,
    is_known_inlier,weight.  The reference loads with ``load_weight=False``
    (``src/PGODatasetPublisherNode.cpp:168-169``) so weights start at 1.
    """
    a = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    m = a.shape[0]
    meas = Measurements(
        r1=a[:, 0].astype(np.int32), p1=a[:, 1].astype(np.int32), r2=a[:, 2].astype(np.int32),
        p2=a[:, 3].astype(np.int32), R=quat_to_rot(a[:, 4:8]), t=a[:, 8:11].copy(), kappa=a[:, 11].copy(),
        tau=a[:, 12].copy(), weight=(a[:, 14].copy() if load_weight else np.ones(m)),
        fixed=a[:, 13].astype(np.uint8))
    return meas


def load_tunnels(ros_message_path: bool = True) -> Tuple[Measurements, List[int]]:
    """The 8-robot tunnels dataset as one global measurement set (deduplicated).

    With ``ros_message_path`` the precisions and inlier flags are what an agent
    sees after the wire round trip: kappa=10000, tau=100 and only odometry is
    fixed (``src/utils.cpp:141-149``) -- identical to the CSV values here.
    """
    parts = []
    for rid in range(8):
        parts.append(read_measurements_csv(os.path.join(DATA_DIR, "tunnels", f"robot{rid}", "measurements.csv")))
    allm = Measurements.concat(parts)
    seen: Dict[Tuple[int, int, int, int], int] = {}
    keep = []
    for e in range(len(allm)):
        key = (int(allm.r1[e]), int(allm.p1[e]), int(allm.r2[e]), int(allm.p2[e]))
        if key in seen:
            continue
        seen[key] = e
        keep.append(e)
    allm = allm.take(keep)
    if ros_message_path:
        allm.kappa[:] = 10000.0
        allm.tau[:] = 100.0
        odo = (allm.r1 == allm.r2) & (allm.p1 + 1 == allm.p2)
        allm.fixed = odo.astype(np.uint8)
    n_per_robot = []
    for rid in range(8):
        mx = -1
        s1 = allm.r1 == rid
        s2 = allm.r2 == rid
        if s1.any():
            mx = max(mx, int(allm.p1[s1].max()))
        if s2.any():
            mx = max(mx, int(allm.p2[s2].max()))
        n_per_robot.append(mx + 1)
    return allm, n_per_robot


# ----------------------------------------------------------------------------
# SE(3) helpers + initial guesses (host side; same arrays go to oracle and GPU)
# ----------------------------------------------------------------------------

def se3_compose(Ra, ta, Rb, tb):
    return Ra @ Rb, Ra @ tb + ta


def se3_inverse(R, t):
    return R.T, -R.T @ t


def odometry_chain(meas: Measurements, robot: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Chain this robot's odometry from pose 0 = identity (robot-local frame).

    Mirrors the Odometry local initialisation selected by
    ``local_initialization_method`` (``src/PGOAgentROSNode.cpp:106-108``).
    """
    R = np.tile(np.eye(3), (n, 1, 1))
    t = np.zeros((n, 3))
    sel = np.nonzero((meas.r1 == robot) & (meas.r2 == robot) & (meas.p1 + 1 == meas.p2))[0]
    by_src = {int(meas.p1[e]): e for e in sel}
    for i in range(n - 1):
        e = by_src.get(i)
        if e is None:
            raise ValueError(f"robot {robot}: missing odometry edge {i}->{i + 1}")
        R[i + 1], t[i + 1] = se3_compose(R[i], t[i], meas.R[e], meas.t[e])
    return R, t


def project_to_so3(M: np.ndarray) -> np.ndarray:
    U, _, Vt = np.linalg.svd(M)
    D = np.diag([1.0, 1.0, np.linalg.det(U @ Vt)])
    return U @ D @ Vt


def robust_frame_alignment(meas: Measurements, traj: Dict[int, Tuple[np.ndarray, np.ndarray]], num_robots: int,
                           rot_tol: float = 0.2, tran_tol: float = 1.0) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
    """World-frame transform of every robot's local frame, robot 0 = identity.

    Breadth-first over the robot graph; for each (known robot a, unknown robot
    b) every shared loop closure votes for T_world_b, and the candidate with the
    largest consensus set wins, then the consensus set is chordal-averaged.
    This stands in for dpgo's robust multi-robot initialisation (SURVEY §8f,
    rank 1 -- outside the hot path); it is deterministic and is applied
    identically to oracle and GPU runs.
    """
    frames: Dict[int, Tuple[np.ndarray, np.ndarray]] = {0: (np.eye(3), np.zeros(3))}
    pending = [r for r in range(1, num_robots)]
    shared = np.nonzero(meas.r1 != meas.r2)[0]
    progress = True
    while pending and progress:
        progress = False
        for b in list(pending):
            cands = []
            for e in shared:
                r1, r2 = int(meas.r1[e]), int(meas.r2[e])
                if r2 == b and r1 in frames:
                    a = r1
                    Ra, ta = traj[a][0][meas.p1[e]], traj[a][1][meas.p1[e]]
                    Rwa, twa = frames[a]
                    Rwi, twi = se3_compose(Rwa, twa, Ra, ta)
                    Rwj, twj = se3_compose(Rwi, twi, meas.R[e], meas.t[e])
                    Rb, tb = traj[b][0][meas.p2[e]], traj[b][1][meas.p2[e]]
                    Rbi, tbi = se3_inverse(Rb, tb)
                    cands.append(se3_compose(Rwj, twj, Rbi, tbi))
                elif r1 == b and r2 in frames:
                    a = r2
                    Ra, ta = traj[a][0][meas.p2[e]], traj[a][1][meas.p2[e]]
                    Rwa, twa = frames[a]
                    Rwj, twj = se3_compose(Rwa, twa, Ra, ta)
                    Rmi, tmi = se3_inverse(meas.R[e], meas.t[e])
                    Rwi, twi = se3_compose(Rwj, twj, Rmi, tmi)
                    Rb, tb = traj[b][0][meas.p1[e]], traj[b][1][meas.p1[e]]
                    Rbi, tbi = se3_inverse(Rb, tb)
                    cands.append(se3_compose(Rwi, twi, Rbi, tbi))
            if not cands:
                continue
            Rs = np.stack([c[0] for c in cands])
            ts = np.stack([c[1] for c in cands])
            best, best_set = -1, None
            for c in range(len(cands)):
                dR = np.linalg.norm(Rs - Rs[c], axis=(1, 2))
                dt = np.linalg.norm(ts - ts[c], axis=1)
                inl = np.nonzero((dR < rot_tol) & (dt < tran_tol))[0]
                if len(inl) > best:
                    best, best_set = len(inl), inl
            Rm = project_to_so3(Rs[best_set].mean(axis=0))
            tm = ts[best_set].mean(axis=0)
            frames[b] = (Rm, tm)
            pending.remove(b)
            progress = True
    for b in pending:  # disconnected robots: leave at identity
        frames[b] = (np.eye(3), np.zeros(3))
    return frames


def with_local_initialization(problem: "Problem", local_trajectory) -> "Problem":
    """`problem` with a new initial guess: `local_trajectory(rid)` -> [n, 3, 4] poses in the robot's own frame (e.g. the
    Chordal initialisation of the oracle or of the GPU agent), placed in the global frame by robust_frame_alignment
    over the shared loop closures (robot 0 = identity) -- the INITIALIZE round of the wrapper
    (src/PGOAgentROS.cpp:322-366, 1091-1159) without its messaging."""
    traj = {}
    for rid in range(problem.num_robots):
        T = np.asarray(local_trajectory(rid))
        traj[rid] = (T[:, :, :3].copy(), T[:, :, 3].copy())
    frames = robust_frame_alignment(problem.meas, traj, problem.num_robots)
    T_init = []
    for rid in range(problem.num_robots):
        Rw, tw = frames[rid]
        T = np.zeros((problem.n[rid], 3, 4))
        T[:, :, :3] = np.einsum("ij,njk->nik", Rw, traj[rid][0])
        T[:, :, 3] = traj[rid][1] @ Rw.T + tw
        T_init.append(T)
    return Problem(name=problem.name + "+init", num_robots=problem.num_robots, meas=problem.meas, n=list(problem.n),
                   T_init=T_init)


def fixed_lifting_matrix(r: int, d: int = 3) -> np.ndarray:
    """A fixed YLift in St(d, r), identical for oracle and GPU runs.

    Deterministic: QR of a closed-form full-rank r x d matrix (no RNG), sign
    fixed so diag(R) > 0.
    """
    i = np.arange(r)[:, None].astype(np.float64)
    j = np.arange(d)[None, :].astype(np.float64)
    A = np.cos(0.7 * (i + 1) * (j + 1)) + 0.3 * np.sin(1.3 * i - 0.4 * j) + (i == j)
    Q, Rr = np.linalg.qr(A)
    Q = Q * np.sign(np.diag(Rr))[None, :]
    return np.ascontiguousarray(Q)


@dataclasses.dataclass
class Problem:
    """A partitioned multi-robot problem ready to hand to oracle / GPU agents."""

    name: str
    num_robots: int
    meas: Measurements  # global set, robot-labelled
    n: List[int]  # poses per robot
    T_init: List[np.ndarray]  # per robot: [n_i, 3, 4] initial poses in the GLOBAL frame

    def robot_measurements(self, rid: int) -> Measurements:
        return measurements_of_robot(self.meas, rid)


def spanning_tree_init(meas: Measurements, num_poses: int) -> Tuple[np.ndarray, np.ndarray]:
    """Initial guess for graphs whose odometry chain is broken (cubicle.g2o, rim.g2o: SURVEY App. C): compose the
    measurements along a breadth-first spanning tree rooted at pose 0, edges taken in file order, either direction."""
    from collections import deque

    adj: List[List[Tuple[int, int, bool]]] = [[] for _ in range(num_poses)]
    for e in range(len(meas)):
        i, j = int(meas.p1[e]), int(meas.p2[e])
        adj[i].append((j, e, True))
        adj[j].append((i, e, False))
    R = np.tile(np.eye(3), (num_poses, 1, 1))
    t = np.zeros((num_poses, 3))
    seen = np.zeros(num_poses, dtype=bool)
    seen[0] = True
    queue = deque([0])
    while queue:
        i = queue.popleft()
        for j, e, forward in adj[i]:
            if seen[j]:
                continue
            if forward:
                R[j], t[j] = se3_compose(R[i], t[i], meas.R[e], meas.t[e])
            else:
                Rinv, tinv = se3_inverse(meas.R[e], meas.t[e])
                R[j], t[j] = se3_compose(R[i], t[i], Rinv, tinv)
            seen[j] = True
            queue.append(j)
    if not seen.all():
        raise ValueError(f"pose graph is disconnected: {int((~seen).sum())} poses unreachable from pose 0")
    return R, t


def _global_odometry_init(meas: Measurements, num_poses: int) -> Tuple[np.ndarray, np.ndarray]:
    try:
        return odometry_chain(meas, 0, num_poses)
    except ValueError:
        return spanning_tree_init(meas, num_poses)


def load_g2o_problem(name: str, num_robots: int, path: str | None = None) -> Problem:
    """g2o dataset split over `num_robots`, odometry initial guess in the global frame."""
    path = path or os.path.join(DATA_DIR, name + ".g2o")
    meas, num_poses = read_g2o(path)
    # de-duplicate (src,dst) pairs the way hasMeasurement does (keeps the first)
    seen, keep = set(), []
    for e in range(len(meas)):
        key = (int(meas.p1[e]), int(meas.p2[e]))
        if key in seen:
            continue
        seen.add(key)
        keep.append(e)
    meas = meas.take(keep)
    Rg, tg = _global_odometry_init(meas, num_poses)
    part, start = partition_contiguous(meas, num_poses, num_robots)
    n = [int(start[r + 1] - start[r]) for r in range(num_robots)]
    T_init = []
    for r in range(num_robots):
        T = np.zeros((n[r], 3, 4))
        T[:, :, :3] = Rg[start[r]:start[r + 1]]
        T[:, :, 3] = tg[start[r]:start[r + 1]]
        T_init.append(T)
    return Problem(name=f"{name}/{num_robots}", num_robots=num_robots, meas=part, n=n, T_init=T_init)


def load_tunnels_problem() -> Problem:
    meas, n = load_tunnels()
    traj = {r: odometry_chain(meas, r, n[r]) for r in range(8)}
    frames = robust_frame_alignment(meas, traj, 8)
    T_init = []
    for r in range(8):
        Rw, tw = frames[r]
        T = np.zeros((n[r], 3, 4))
        for i in range(n[r]):
            Ri, ti = se3_compose(Rw, tw, traj[r][0][i], traj[r][1][i])
            T[i, :, :3] = Ri
            T[i, :, 3] = ti
        T_init.append(T)
    return Problem(name="tunnels/8", num_robots=8, meas=meas, n=n, T_init=T_init)


def _snake_lattice(num_poses: int):
    """Lattice coordinates of a 3-D boustrophedon path: consecutive poses are always lattice neighbours."""
    side = int(np.ceil(num_poses ** (1.0 / 3.0)))
    nx = ny = side
    nz = int(np.ceil(num_poses / (nx * ny)))
    idx = np.arange(nx * ny * nz)
    z = idx // (nx * ny)
    rem = idx % (nx * ny)
    yy = rem // nx
    y = np.where(z % 2 == 0, yy, ny - 1 - yy)          # alternate the row order per layer
    row = z * ny + yy                                  # global row counter along the path
    xx = rem % nx
    x = np.where(row % 2 == 0, xx, nx - 1 - xx)        # alternate the direction per row
    coords = np.stack([x, y, z], axis=1)[:num_poses]
    lut = -np.ones((nx, ny, nz), dtype=np.int64)
    lut[coords[:, 0], coords[:, 1], coords[:, 2]] = np.arange(num_poses)
    return coords, lut, (nx, ny, nz)


def make_synthetic_problem(num_poses: int, num_edges: int, num_robots: int, seed: int = 0,
                           kappa: float = 200.0, tau: float = 100.0, init: str = "near_truth") -> Problem:
    """Seeded synthetic SE(3) graph for BASELINE config 5 (100k poses / 1M edges / 8 agents at full size).

    The trajectory snakes through a unit 3-D lattice (like the grid3D / torus3D benchmark graphs): pose i sits
    on lattice site i of a boustrophedon path, its orientation follows a random walk exp(N(0, 0.2^2 I)) per
    step.  Edges: the `num_poses - 1` odometry edges plus loop closures drawn without replacement (seeded) from
    all non-consecutive pose pairs whose sites are within the 26-neighbourhood (relative translations <= sqrt 3,
    so the graph is as well conditioned as the real benchmarks).  Noise: rotation exp(N(0, 0.05^2 I)),
    translation N(0, 0.1^2 I); constant kappa / tau.  Contiguous split over `num_robots` (the dataset
    publisher's rule, src/PGODatasetPublisherNode.cpp:84-103).

    init: "near_truth" perturbs the ground truth by exp(N(0, 0.05^2 I)) / N(0, 0.1^2 I) and stands in for the
    Chordal initialisation of the async demo (launch/asapp_demo.launch:10; initialisation is SURVEY 8f rank 1,
    outside the hot path); "odometry" chains the odometry edges from pose 0.
    """
    rng = np.random.default_rng(seed)

    def expm_so3(w):
        th = np.linalg.norm(w, axis=-1, keepdims=True)
        th = np.maximum(th, 1e-12)
        k = w / th
        K = np.zeros(w.shape[:-1] + (3, 3))
        K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
        K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
        K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
        s = np.sin(th)[..., None]
        c = np.cos(th)[..., None]
        return np.eye(3) + s * K + (1 - c) * (K @ K)

    coords, lut, dims = _snake_lattice(num_poses)
    tgt = coords.astype(np.float64)
    dR = expm_so3(rng.normal(0, 0.2, size=(num_poses - 1, 3)))
    Rgt = np.empty((num_poses, 3, 3))
    Rgt[0] = np.eye(3)
    for i in range(num_poses - 1):
        Rgt[i + 1] = Rgt[i] @ dR[i]
    # candidate loop closures: half of the 26-neighbourhood (each unordered pair once), minus the odometry pairs
    cs, cd = [], []
    offs = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dz, dy, dx) > (0, 0, 0)]
    for dx, dy, dz in offs:
        q = coords + np.array([dx, dy, dz])
        ok = ((q >= 0) & (q < np.array(dims))).all(axis=1)
        a = np.nonzero(ok)[0]
        b = lut[q[a, 0], q[a, 1], q[a, 2]]
        keep = (b >= 0) & (np.abs(b - a) > 1)
        a, b = a[keep], b[keep]
        cs.append(np.minimum(a, b))
        cd.append(np.maximum(a, b))
    cs, cd = np.concatenate(cs), np.concatenate(cd)
    n_lc = num_edges - (num_poses - 1)
    if n_lc > len(cs):
        raise ValueError(f"only {len(cs)} lattice loop closures available for {num_poses} poses, asked for {n_lc}")
    pick = rng.choice(len(cs), size=n_lc, replace=False)
    pick.sort()
    src = np.concatenate([np.arange(num_poses - 1, dtype=np.int64), cs[pick]])
    dst = np.concatenate([np.arange(1, num_poses, dtype=np.int64), cd[pick]])
    m = src.shape[0]
    Rn = expm_so3(rng.normal(0, 0.05, size=(m, 3)))
    tn = rng.normal(0, 0.1, size=(m, 3))
    Rrel = np.einsum("mji,mjk->mik", Rgt[src], Rgt[dst]) @ Rn
    trel = np.einsum("mji,mj->mi", Rgt[src], tgt[dst] - tgt[src]) + tn
    meas = Measurements(r1=np.zeros(m, np.int32), p1=src.astype(np.int32), r2=np.zeros(m, np.int32),
                        p2=dst.astype(np.int32), R=Rrel, t=trel, kappa=np.full(m, kappa), tau=np.full(m, tau),
                        weight=np.ones(m), fixed=np.zeros(m, np.uint8))
    if init == "near_truth":
        Rg = Rgt @ expm_so3(rng.normal(0, 0.05, size=(num_poses, 3)))
        tg = tgt + rng.normal(0, 0.1, size=(num_poses, 3))
    elif init == "odometry":
        Rg, tg = _global_odometry_init(meas, num_poses)
    else:
        raise ValueError(init)
    part, start = partition_contiguous(meas, num_poses, num_robots)
    n = [int(start[r + 1] - start[r]) for r in range(num_robots)]
    T_init = []
    for r in range(num_robots):
        T = np.zeros((n[r], 3, 4))
        T[:, :, :3] = Rg[start[r]:start[r + 1]]
        T[:, :, 3] = tg[start[r]:start[r + 1]]
        T_init.append(T)
    return Problem(name=f"synthetic{num_poses}/{num_robots}", num_robots=num_robots, meas=part, n=n, T_init=T_init)


def make_random_walk_problem(num_poses: int = 100000, num_edges: int = 1000000, num_robots: int = 8, seed: int = 0,
                             kappa: float = 200.0, tau: float = 100.0, window: int = 2000, far_fraction: float = 0.1,
                             g2o_path: str | None = None) -> Problem:
    """BASELINE config 5 exactly as SURVEY 8(d) specifies it (the lattice graph of make_synthetic_problem is the
    better-conditioned stand-in round 1 benchmarked; both are kept and the bench line says which one ran):

      * poses on a seeded 3-D random walk, numpy default_rng(seed): step i -> i+1 = translation U([0.5, 1.5]) e_x in the
        body frame, then a rotation exp(w^), w ~ N(0, 0.2^2 I);
      * num_poses - 1 odometry edges + (num_edges - num_poses + 1) loop closures (i, j), i < j, no duplicates, none with
        j = i + 1: 90 % with |i - j| <= `window` (2000), 10 % uniform over all pairs;
      * measurement noise: rotation exp(N(0, 0.05^2 I)), translation N(0, 0.1^2 I); kappa = 200, tau = 100 on every edge;
      * initial guess: the odometry chain from pose 0 (global, then cut per robot);
      * contiguous split over `num_robots` (src/PGODatasetPublisherNode.cpp:84-103).

    g2o_path: also write the graph as EDGE_SE3:QUAT with the matching information matrices (tau I_3 for the
    translation, 2 kappa I_3 for the rotation: the SE-Sync rule of SURVEY App. B gives kappa = 3 / (2 tr(I_R^-1)),
    tau = 3 / tr(I_t^-1) back)."""
    rng = np.random.default_rng(seed)

    def expm_so3(w):
        th = np.maximum(np.linalg.norm(w, axis=-1, keepdims=True), 1e-12)
        k = w / th
        K = np.zeros(w.shape[:-1] + (3, 3))
        K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
        K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
        K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
        return np.eye(3) + np.sin(th)[..., None] * K + (1 - np.cos(th))[..., None] * (K @ K)

    step_len = rng.uniform(0.5, 1.5, size=num_poses - 1)
    dR = expm_so3(rng.normal(0, 0.2, size=(num_poses - 1, 3)))
    Rgt = np.empty((num_poses, 3, 3))
    tgt = np.zeros((num_poses, 3))
    Rgt[0] = np.eye(3)
    for i in range(num_poses - 1):
        tgt[i + 1] = tgt[i] + Rgt[i][:, 0] * step_len[i]
        Rgt[i + 1] = Rgt[i] @ dR[i]
    n_lc = num_edges - (num_poses - 1)
    if n_lc < 0:
        raise ValueError("num_edges must be at least num_poses - 1")
    n_far = int(round(far_fraction * n_lc))
    n_near = n_lc - n_far
    have = set()
    src_l, dst_l = [], []

    def draw(count, near):
        got = 0
        while got < count:
            k = max(1024, 2 * (count - got))
            a = rng.integers(0, num_poses, size=k)
            if near:
                d = rng.integers(2, window + 1, size=k)
                b = a + d
            else:
                b = rng.integers(0, num_poses, size=k)
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            ok = (hi < num_poses) & (hi - lo >= 2)
            for x, y in zip(lo[ok].tolist(), hi[ok].tolist()):
                key = x * num_poses + y
                if key in have:
                    continue
                have.add(key)
                src_l.append(x)
                dst_l.append(y)
                got += 1
                if got == count:
                    break

    draw(n_near, True)
    draw(n_far, False)
    order = np.lexsort((np.array(dst_l, dtype=np.int64), np.array(src_l, dtype=np.int64)))
    src = np.concatenate([np.arange(num_poses - 1, dtype=np.int64), np.array(src_l, dtype=np.int64)[order]])
    dst = np.concatenate([np.arange(1, num_poses, dtype=np.int64), np.array(dst_l, dtype=np.int64)[order]])
    m = src.shape[0]
    Rn = expm_so3(rng.normal(0, 0.05, size=(m, 3)))
    tn = rng.normal(0, 0.1, size=(m, 3))
    Rrel = np.einsum("mji,mjk->mik", Rgt[src], Rgt[dst]) @ Rn
    trel = np.einsum("mji,mj->mi", Rgt[src], tgt[dst] - tgt[src]) + tn
    meas = Measurements(r1=np.zeros(m, np.int32), p1=src.astype(np.int32), r2=np.zeros(m, np.int32),
                        p2=dst.astype(np.int32), R=Rrel, t=trel, kappa=np.full(m, kappa), tau=np.full(m, tau),
                        weight=np.ones(m), fixed=np.zeros(m, np.uint8))
    if g2o_path:
        write_g2o(g2o_path, meas, num_poses)
    Rg, tg = _global_odometry_init(meas, num_poses)
    part, start = partition_contiguous(meas, num_poses, num_robots)
    n = [int(start[r + 1] - start[r]) for r in range(num_robots)]
    T_init = []
    for r in range(num_robots):
        T = np.zeros((n[r], 3, 4))
        T[:, :, :3] = Rg[start[r]:start[r + 1]]
        T[:, :, 3] = tg[start[r]:start[r + 1]]
        T_init.append(T)
    return Problem(name=f"randomwalk{num_poses}/{num_robots}", num_robots=num_robots, meas=part, n=n, T_init=T_init)

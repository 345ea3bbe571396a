#!/usr/bin/env python
"""bench.py -- RBCD iterations/s on sphere2500 split over 8 agents (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle on the host cores

One *step* = one global RBCD iteration of the synchronous schedule: the robot
holding the UPDATE token runs iterate(true), every other robot iterate(false),
public (+ auxiliary) poses are exchanged (src/PGOAgentROS.cpp:1161-1189,
109-113).  Workload = BASELINE config 2: sphere2500.g2o, 8 agents, r = 5, RGD
(stepsize 0.2, preconditioner on: launch/asapp_demo.launch:7-8) + Nesterov
acceleration (restart 50), RoundRobin, odometry initial guess, fixed YLift.

`value`  : all 8 agents resident on the GPU(s), the persistent kernel runs the K
           steps with device-side exchange (N > 1: one persistent kernel per GPU,
           public poses stored into peer memory over NVLink -- the fabric); timed
           with CUDA events inside the library around the launches (on the
           launching stream), max over ranks.
`e2e`    : the same K steps driven through the per-robot C ABI that PGOAgentROS
           would call (iterate / getSharedPoseDictWithNeighbor / updateNeighborPoses)
           with HOST buffers, one OS thread per robot: every step's public poses
           cross PCIe both ways (N > 1: and a shared-memory segment between the
           per-GPU processes).
Secondary objects on the same line: `reference_wrapper` (the reference's unmodified
PGOAgentROS running its demo launch file on both back ends), `async_mode` (the reference's asynchronous
demo configuration as parallel ticks) and, at N = 1, `hbm_bound_regime` (one rank
of BASELINE config 5 at the named size -- the HBM-bound regime of this path).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIG2 = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50,
               rel_change_tol=0.0, max_num_iters=10 ** 9)  # tolerance 0: the bench never stops early
WORKLOAD = "sphere2500.g2o / 8 agents / r=5 / RGD(step 0.2, precond) + Nesterov(restart 50) / RoundRobin"
# the reference's asynchronous demo (launch/asapp_demo.launch:2-10: sphere2500, RGD 0.2 + preconditioner, no
# acceleration) as its equal-rate / unit-delay schedule: every robot steps every tick (secondary figure, not `value`)
ASYNC_CONFIG = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=0, rel_change_tol=0.0,
                    max_num_iters=10 ** 9)


def async_cpu_reference(ticks=300, threads=None):
    from dpgo_ros_b200 import datasets
    from oracle import binding as orc
    threads = threads or min(8, os.cpu_count() or 1)
    pb = datasets.load_g2o_problem("sphere2500", 8)
    team = orc.OracleTeam(pb, **ASYNC_CONFIG)
    team.run_parallel(20, threads=threads)
    res = team.run_parallel(ticks, threads=threads)
    return res.iterations / res.wall_seconds


def hbm_regime_single_rank(device, iters=6, cpu_beside=True, generator="lattice"):
    """Secondary figure: one rank of BASELINE config 5 at the named size (synthetic 100k poses / 1M edges / 8 agents,
    robot 0: 12 500 poses) -- one iterate(true) streams the 20 GB dense preconditioner once, the regime in which the
    HBM roofline is the physical bound.  Same measurement as tools/bench_config5.py."""
    import time as _t
    from dpgo_ros_b200 import agent as gpu, datasets
    # generator "lattice": datasets.make_synthetic_problem (banded graph: the CPU oracle's sparse Cholesky can run beside
    # it); "random_walk": datasets.make_random_walk_problem, the generator SURVEY 8(d) specifies (seeded random walk, 90 %
    # of the loop closures within 2000 poses + 10 % uniform, odometry guess) -- its far loop closures fill the oracle's
    # sparse factor in (no result after 6 minutes on one core), so the CPU runs beside the lattice instance only
    if generator == "random_walk":
        pb = datasets.make_random_walk_problem(100000, 1000000, 8, seed=0)
        cpu_beside = False
    else:
        pb = datasets.make_synthetic_problem(100000, 1000000, 8, seed=0)
    P = gpu.make_params(num_robots=8, **ASYNC_CONFIG)
    yl = datasets.fixed_lifting_matrix(P.r)
    eye = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)
    ag = gpu.PGOAgent(0, P, device)
    m = pb.robot_measurements(0)
    ag.addMeasurements(m)
    ag.setLiftingMatrix(yl)
    ag.initialize(pb.T_init[0])
    ag.initializeInGlobalFrame(eye)
    need = {}
    for e in np.nonzero(m.r1 != m.r2)[0]:
        o, f = (int(m.r2[e]), int(m.p2[e])) if int(m.r1[e]) == 0 else (int(m.r1[e]), int(m.p1[e]))
        need.setdefault(o, set()).add(f)
    for o, frames in need.items():
        fr = np.array(sorted(frames), dtype=np.int32)
        ag.updateNeighborPoses(o, fr, np.ascontiguousarray(np.einsum("ak,nkc->nca", yl, pb.T_init[o][fr])), False)
    t_build = _t.perf_counter()
    ag.iterate(True)   # builds Q and the 50 016^2 dense inverse
    t_build = _t.perf_counter() - t_build
    ts = []
    for _ in range(iters):
        t0 = _t.perf_counter()
        ag.iterate(True)
        ts.append(_t.perf_counter() - t0)
    n = pb.n[0]
    # the Riemannian-gradient kernel of this regime on its own (k_edge_grad), cold L2 like inside a step, device time
    # between the first CTA's start and the last CTA's end: the "per-iteration Riemannian-gradient kernel" of the north star
    shared = m.r1 != m.r2
    nbr_pub = sum(len(v) for v in need.values())
    b_grad = len(m) * 128 + 2 * n * P.r * 4 * 8 + nbr_pub * P.r * 4 * 8
    grad = {}
    try:
        kns, ens = [], []
        for _ in range(9):
            _, _, k, e = ag.edgeGrad(None, flush_l2=True)
            kns.append(k)
            ens.append(e)
        peak, _ = load_peaks()
        us = float(np.median(kns)) * 1e-3
        grad = {"kernel": "k_edge_grad<5> (dpgo_ros_b200/csrc/edge_grad.cu)", "us": us, "us_cuda_events": float(np.median(ens)) * 1e-3,
                "b_grad_bytes": b_grad, "achieved_GBps": b_grad / (us * 1e-6) / 1e9, "peak_GBps": peak,
                "frac": b_grad / (us * 1e-6) / 1e9 / peak,
                "how": "median of 9 launches, each after rewriting 512 MB (cold L2, as inside a step where the 20 GB "
                       "preconditioner has just streamed through); device time = last CTA end - first CTA start "
                       "(globaltimer); bytes = SURVEY 8(d) B_grad of this agent"}
    except Exception as e:  # noqa: BLE001
        grad = {"error": str(e)[:200]}
    parity = None
    if generator == "random_walk":
        try:   # f and the Riemannian gradient at the current iterate against the oracle's (no factorisation involved)
            from oracle import binding as orc
            oteam = orc.OracleTeam(pb, **ASYNC_CONFIG)
            oteam.exchange_all()
            X = ag.getX()
            f, rg, _, _ = ag.edgeGrad(X)
            fo, _, rgo = oteam.eval(0, X)
            parity = {"what": "k_edge_grad vs oracle at the iterate after %d iterate(true)" % (iters + 1),
                      "f_rel_diff": abs(f - fo) / abs(fo), "rgrad_rel_diff": float(np.linalg.norm(rg - rgo) / np.linalg.norm(rgo))}
        except Exception as e:  # noqa: BLE001
            parity = {"error": str(e)[:200]}
    ag.close()
    npad = (4 * n + 31) // 32 * 32
    nbytes = npad * npad * 8 + 2 * n * P.r * 4 * 8 + len(m) * 128 + 2 * n * P.r * 4 * 8
    ms = float(np.median(ts)) * 1e3
    peak, _ = load_peaks()
    gen_note = ("SURVEY 8(d) generator datasets.make_random_walk_problem: seeded random walk, 90 % of the loop closures within "
                "2000 poses + 10 % uniform, odometry guess" if generator == "random_walk" else
                "lattice generator datasets.make_synthetic_problem (the CPU oracle can factor it; the SURVEY 8(d) random-walk "
                "generator runs in hbm_bound_regime_8d_generator)")
    out = {"workload": "config 5 at the named size, one of its 8 ranks: robot 0 of the synthetic 100k-pose / 1M-edge graph "
                       f"(n={n}, {len(m)} edges; {gen_note}), RGD 0.2 + dense preconditioner, per-robot C ABI iterate(true)",
           "ms_per_iterate": ms, "preconditioner_build_s": t_build, "algorithmic_bytes": nbytes,
           "achieved_GBps": nbytes / (ms * 1e-3) / 1e9, "peak_GBps": peak, "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
           "frac_note": "implementation bytes (the 20 GB dense inverse streamed once per step), NOT SURVEY 8(d) bytes",
           "riemannian_gradient_kernel": grad}
    if parity is not None:
        out["parity_vs_oracle"] = parity
        out["cpu_beside"] = {"note": "the oracle's sparse Cholesky of this robot's Q + 0.1 I did not finish in 6 minutes on one core "
                                     "(fill-in from the far loop closures); the GPU's dense inverse: preconditioner_build_s"}
    if cpu_beside:
        try:   # the same robot's iterate(true) on the CPU oracle (sparse Cholesky preconditioner), one core
            from oracle import binding as orc
            oteam = orc.OracleTeam(pb, **ASYNC_CONFIG)
            t0 = _t.perf_counter()
            oteam.iterate(0, True)
            t_fact = _t.perf_counter() - t0
            cs = []
            for _ in range(5):
                t0 = _t.perf_counter()
                oteam.iterate(0, True)
                cs.append(_t.perf_counter() - t0)
            out["cpu_beside"] = {"ms_per_iterate": float(np.median(cs)) * 1e3, "first_iterate_s": t_fact, "cores": 1,
                                 "kind": "port", "speedup": float(np.median(cs)) * 1e3 / ms}
        except Exception as e:  # noqa: BLE001
            out["cpu_beside"] = {"error": str(e)[:200]}
    return out


def async_mode_single_gpu(pb, device, ticks=2000):
    """Secondary figure: the asynchronous mode (all 8 robots step every tick) on one GPU."""
    from dpgo_ros_b200 import agent as gpu
    team, agents = gpu.make_team(pb, device=device, **ASYNC_CONFIG)
    team.set_schedule(1)
    team.run(200, stop_on_terminate=False)
    res = team.run(ticks, stop_on_terminate=False)
    cost = team.global_cost()
    team.close()
    for a in agents:
        a.close()
    tps = ticks / (res.device_ms * 1e-3)
    return {"workload": "sphere2500.g2o / 8 agents / RGD(step 0.2, precond), no acceleration / every robot steps "
                        "every tick (asynchronous mode, equal-rate unit-delay schedule)",
            "ticks_per_s": tps, "robot_updates_per_s": tps * pb.num_robots, "us_per_tick": 1e6 / tps,
            "final_cost_2f": cost}


WRAPPER_RUNS = (
    # (key, BASELINE config, command-line of oracle/_ref/dpgo_ros_inproc_*, what it is)
    ("sphere2500_5_odometry", None,
     ["--robots", "5", "--g2o", "data/sphere2500.g2o", "--preset", "dpgo_demo", "--param", "local_initialization_method=Odometry"],
     "launch/dpgo_demo.launch, README.md:32 command: sphere2500.g2o / 5 robots / RTR 3x50 / Odometry guess"),
    ("config3_torus3D_4_r6", 3,
     ["--robots", "4", "--g2o", "data/torus3D.g2o", "--preset", "dpgo_demo", "--param", "relaxation_rank=6"],
     "BASELINE config 3: launch/dpgo_demo.launch on torus3D.g2o / 4 robots / r = 6 / RTR 3x50 / Chordal guess"),
    ("config4_tunnels_8_gnc", 4,
     ["--robots", "8", "--measurements", "data/tunnels", "--preset", "gnc_demo"],
     "BASELINE config 4: launch/dpgo_gnc_demo.launch on the tunnels dataset / 8 robots / GNC_TLS, 3 weight updates"),
)


def reference_wrapper_e2e(timeout_s=60, arms=("b200", "oracle"), repeats=2):
    """Secondary figures: the reference's OWN wrapper (unmodified sources built by oracle/Makefile.ref into oracle/_ref,
    DESIGN.md 5.1 / 6.1) running its demo launch files in one process, once on libdpgo_b200.so and once on the CPU
    oracle -- the only numbers that run the reference's real host code, and the driver-run numbers of BASELINE configs 3
    and 4 (RTR 3x50 is what the ROS node forces in synchronous mode, src/PGOAgentROSNode.cpp:82-87).  Per run:
    wall-clock seconds between the first UPDATE command and TERMINATE (best of `repeats`), and the part of it spent
    inside the library (every dpgo_b200_* entry point; the rest is the wrapper's message handling + the ROS stand-in,
    identical on both arms).  Never fails the bench line."""
    import subprocess
    import tempfile

    out = {"what": "the unmodified PGOAgentROS on both back ends (kappa = 10000, tau = 100 on the message path); "
                   "wall_seconds: first UPDATE -> TERMINATE; library_seconds: inside dpgo_b200_* calls during that window"}
    for key, _cfg, argv, what in WRAPPER_RUNS:
        entry = {"workload": what}
        for arm in arms:
            exe = os.path.join(ROOT, "oracle", "_ref", "dpgo_ros_inproc_" + arm)
            if not os.path.exists(exe):
                entry[arm] = {"unavailable": "oracle/_ref is not built (needs the reference sources at build time)"}
                continue
            best = None
            try:
                for _ in range(repeats):
                    with tempfile.TemporaryDirectory() as tmp:
                        res = os.path.join(tmp, "r.json")
                        cmd = [exe] + [os.path.join(ROOT, a) if a.startswith("data/") else a for a in argv] + ["--out", res, "--log", "0"]
                        p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=timeout_s)
                        if p.returncode != 0:
                            entry[arm] = {"error": (p.stderr or "")[-200:]}
                            break
                        d = json.load(open(res))
                    it, wall = d["round_iterations"][0], d["round_wall_seconds"][0]
                    lib = (d.get("round_library_seconds") or [None])[0]
                    if best is None or wall < best["wall_seconds"]:
                        best = {"backend": d["backend"], "iterations": it, "wall_seconds": wall, "library_seconds": lib,
                                "iters_per_s": it / wall, "gpu_kernel_launches": d["kernel_launches"]}
                if best is not None:
                    entry[arm] = best
            except Exception as e:  # noqa: BLE001
                entry[arm] = {"error": str(e)[:200]}
        if all(isinstance(entry.get(a), dict) and "wall_seconds" in entry[a] for a in ("b200", "oracle") if a in arms) and len(arms) == 2:
            entry["speedup_wall"] = entry["oracle"]["wall_seconds"] / entry["b200"]["wall_seconds"]
            if entry["b200"].get("library_seconds") and entry["oracle"].get("library_seconds"):
                entry["speedup_library"] = entry["oracle"]["library_seconds"] / entry["b200"]["library_seconds"]
        out[key] = entry
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the bench runs."""

    def __init__(self, index=0):
        self.rows = []
        self.stop = False
        self.index = index
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(problem, r):
    """Per-step algorithmic bytes, averaged over the RoundRobin cycle.

    B_grad  = edges*128 + 2*n*r*4*8 + n_nbr_pub*r*4*8     per Riemannian-gradient evaluation
    one RGD step of the selected agent = 2 B_grad (gradient at Y, statistics at X+) + B_precond
    B_precond            = (4n)^2 * 8 + 2*n*r*4*8                      dense (r x 4n)(4n x 4n) product
    Nesterov bookkeeping of ALL agents = sum_a 4 * n_a*r*4*8           read X, V; write Y, X (or V)
    """
    per_agent = []
    for rid in range(problem.num_robots):
        m = problem.robot_measurements(rid)
        n = problem.n[rid]
        shared = m.r1 != m.r2
        nbr_pub = len({(int(a), int(b)) for a, b in zip(np.where(m.r1[shared] == rid, m.r2[shared], m.r1[shared]),
                                                        np.where(m.r1[shared] == rid, m.p2[shared], m.p1[shared]))})
        b_grad = len(m) * 128 + 2 * n * r * 4 * 8 + nbr_pub * r * 4 * 8
        b_pre = (4 * n) ** 2 * 8 + 2 * n * r * 4 * 8
        per_agent.append((b_grad, b_pre, n))
    nest = sum(4 * n * r * 4 * 8 for _, _, n in per_agent)
    step = [2 * bg + bp + nest for bg, bp, _ in per_agent]
    return float(np.mean(step)), float(np.mean([bg for bg, _, _ in per_agent]))


def cpu_reference(steps, warmup, threads=None, sample_note=True):
    """The CPU arm: oracle/ (a port -- the reference's own arithmetic is not vendored) on the host cores."""
    from dpgo_ros_b200 import datasets
    from oracle import binding as orc
    native = orc.use_native()   # -O3 -march=native on this very machine, like the reference's CMakeLists.txt:9
    cores = os.cpu_count() or 1
    threads = threads or min(8, cores)
    pb = datasets.load_g2o_problem("sphere2500", 8)
    team = orc.OracleTeam(pb, **CONFIG2)
    # the CPU arm is timed warm: thread pool up, factorisations done, caches hot (round 1's 20 cold steps read 730 it/s
    # where the steady state is ~900)
    warmup = max(warmup, 300)
    team.run(warmup, threads=threads, stop_on_terminate=False)
    res = team.run(steps, threads=threads, stop_on_terminate=False)
    out = dict(value=res.iterations / res.wall_seconds, seconds=res.wall_seconds, steps=res.iterations,
               threads=threads, cores=cores, cost=team.global_cost(),
               build="-O3 -march=native, built on this machine" if native else "-O3 -march=x86-64-v3 (prebuilt; native build unavailable)")
    # "also report per-iterate(true) median us" -- one robot's local solve on one core, neighbours' poses fresh
    try:
        samples = []
        for k in range(64):
            rid = k % pb.num_robots
            t0 = time.perf_counter()
            team.iterate(rid, True)
            samples.append((time.perf_counter() - t0) * 1e6)
            team.exchange_all()
        out["iterate_true_median_us"] = float(np.median(samples))
    except Exception:  # noqa: BLE001
        out["iterate_true_median_us"] = None
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = args.steps
    t0 = time.time()
    r = cpu_reference(steps, args.warmup)
    line = {
        "impl": "reference", "metric": "rbcd_iters_per_sec", "value": r["value"], "unit": "iters/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": 1e3 / r["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "sphere2500.g2o",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": r["value"], "unit": "iters/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{r['steps']} steps of the same workload after 300 warm-up steps, one OS thread per "
                                   f"agent ({r['threads']} threads on {r['cores']} host cores); {r['build']}",
                         "iterate_true_median_us": r.get("iterate_true_median_us")},
        "e2e": {"value": r["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_cost_2f": r["cost"], "wall_s": time.time() - t0,
        # the reference's own wrapper (unmodified sources, oracle/_ref) on the CPU oracle back end -- no CUDA on this arm
        "reference_wrapper": reference_wrapper_e2e(arms=("oracle",)),
    }
    print(json.dumps(line))
    return 0


def e2e_host_exchange(problem, steps, warmup, device):
    """K steps through the per-robot C ABI with host-buffer exchange (the PGOAgentROS call sequence,
    replayed natively by dpgo_b200_sync_driver_run with one OS thread per robot)."""
    from dpgo_ros_b200 import agent as gpu
    _, agents = gpu.make_team(problem, device=device, colocate=False, **CONFIG2)
    gpu.exchange_host(agents, accel=True)
    gpu.sync_driver_run(agents, warmup, True)
    sec, _ = gpu.sync_driver_run(agents, steps, True)
    payload = gpu.exchange_payload_bytes(agents, True)
    X = [a.getX() for a in agents]
    for a in agents:
        a.close()
    return steps / sec, float(payload), float(payload), X


def e2e_step_count(args):
    """Steps the e2e arm is timed over -- the same rule at every N (round 1 used 4000 at N = 1 and 20 at N > 1)."""
    return max(args.steps, args.e2e_steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=3000, help="bounded CPU-baseline sample (steps)")
    ap.add_argument("--e2e-steps", type=int, default=2000, help="lower bound of the steps the e2e arm is timed over")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary objects (async_mode, hbm_bound_regime, reference_wrapper): profiler runs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from dpgo_ros_b200 import dist
        return dist.bench_multi_gpu(args, CONFIG2, WORKLOAD)

    from dpgo_ros_b200 import agent as gpu
    from dpgo_ros_b200 import capi, datasets

    L = capi.lib()
    if L.dpgo_b200_device_count() < 1:
        print(json.dumps({"error": "no CUDA device: the RBCD path has no CPU fallback"}))
        return 1
    pb = datasets.load_g2o_problem("sphere2500", 8)
    launches0 = L.dpgo_b200_kernel_launch_count()
    team, agents = gpu.make_team(pb, device=local_rank, **CONFIG2)
    with ClockSampler(local_rank) as clk:
        # warm-up (also keeps the clocks up for the sampler): W steps, then ~1 s of back-to-back steps
        team.run(args.warmup, stop_on_terminate=False)
        t_end = time.time() + 1.0
        while time.time() < t_end:
            team.run(2000, stop_on_terminate=False)
        launches_before = L.dpgo_b200_kernel_launch_count()
        res = team.run(args.steps, stop_on_terminate=False)   # <- the timed region (CUDA events inside)
        launches_timed = L.dpgo_b200_kernel_launch_count() - launches_before
        time.sleep(0.25)
    assert res.iterations == args.steps
    ms_per_step = res.device_ms / args.steps
    value = 1e3 / ms_per_step
    cost = team.global_cost()
    # cold-L2 variant: one step per launch with a >L2 buffer rewritten between launches
    cold_ms = None
    try:
        import torch
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local_rank}")
        tt = []
        for _ in range(24):
            flush.fill_(1)
            torch.cuda.synchronize()
            r1 = team.run(1, stop_on_terminate=False)
            tt.append(r1.device_ms)
        cold_ms = float(np.median(tt))
        del flush
    except Exception:
        pass
    team.close()
    for a in agents:
        a.close()

    e2e_val, h2d, d2h, _ = e2e_host_exchange(pb, e2e_step_count(args), max(3, min(args.warmup, 20)), local_rank)

    peak, peak_src = load_peaks()
    step_bytes, grad_bytes = algorithmic_bytes(pb, CONFIG2["r"])
    # achieved = B_grad x Riemannian-gradient evaluations / kernel time.  One step evaluates the gradient of
    # the selected agent twice (the step's own gradient at Y, and f / |grad| at X+ for mLocalOptResult).
    grad_evals_per_step = 2
    achieved = grad_evals_per_step * grad_bytes * args.steps / (res.device_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for tname in ("traffic_r2.json", "traffic_r1.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):  # dram__bytes_read+write of one `ncu --set full` capture of this kernel, per step
            traffic = json.load(open(tpath))["traffic_bytes_per_step"] * args.steps
            traffic_src = f"profiles/{tname} (ncu --set full capture of this kernel, bytes per step x steps; not measured in this run)"
            break
    model_achieved = step_bytes * args.steps / (res.device_ms * 1e-3) / 1e9
    cpu = cpu_reference(args.cpu_steps, 50)
    if args.no_secondary:
        async_mode = hbm_regime = hbm_regime_8d = wrapper = {"skipped": "--no-secondary"}
    else:
        async_mode = async_mode_single_gpu(pb, local_rank)
        async_mode["cpu_ticks_per_s"] = async_cpu_reference()
        try:
            hbm_regime = hbm_regime_single_rank(local_rank)
        except Exception as e:  # noqa: BLE001  (secondary figure: never fail the bench line over it)
            hbm_regime = {"error": str(e)[:200]}
        try:
            hbm_regime_8d = hbm_regime_single_rank(local_rank, generator="random_walk")
        except Exception as e:  # noqa: BLE001
            hbm_regime_8d = {"error": str(e)[:200]}
        wrapper = reference_wrapper_e2e()
    line = {
        "metric": "rbcd_iters_per_sec", "value": value, "unit": "iters/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "sphere2500.g2o (reference data/, odometry initial guess)",
        "config": {"workload": WORKLOAD, "agents_per_gpu": 8,
                   "l2": "steady state: one persistent launch runs all K steps and re-reads the same ~105 MB "
                         "(8 dense preconditioners + graph) every RoundRobin cycle, as the solver does; no flush "
                         "inside the launch. cold_l2_ms_per_step = one step per launch after rewriting 512 MB"},
        "cold_l2_ms_per_step": cold_ms,
        "final_cost_2f": cost,
        "gpu_launches": int(launches_timed),
        "kernel_launches_total": int(L.dpgo_b200_kernel_launch_count() - launches0),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_val, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "per-robot C ABI (iterate / getSharedPoseDict / updateNeighborPoses), host buffers, 8 agents"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "k_team_run<5,1> (persistent: all phases of all K steps in one launch)",
                     "accounting": "SURVEY 8(d): B_grad = edges*128 + 2*n*r*4*8 + n_nbr_pub*r*4*8 per Riemannian-gradient "
                                   "evaluation, 2 evaluations per step, / CUDA-event time of the launch",
                     "b_grad_per_agent": grad_bytes, "grad_evals_per_step": grad_evals_per_step,
                     "note": "config 2 is latency-bound, not bandwidth-bound: one agent's gradient is ~200 KB and the "
                             "whole step is L2-resident, so the HBM fraction of the 8(d) bytes is ~0.003-0.005 by "
                             "construction (a step is a chain of ~10 dependent L2 round trips and 2 grid barriers). "
                             "FP64 throughout: tcgen05 has no FP64 path and DMMA m8n8k4 measures the DFMA rate on "
                             "B200 (profiles/microbench_r1.txt), so the dense (r x 4n)(4n x 4n) product stays on the "
                             "FP64 pipe -- tensor cores declined with evidence. The HBM-bound regime of this path is "
                             "config 5: see hbm_bound_regime.",
                     "effective_step_model": {
                         "what": "NOT the 8(d) figure: bytes the implementation's own step touches (2 B_grad + the "
                                 "12.5 MB dense preconditioner of the selected agent + Nesterov bookkeeping of all "
                                 "agents), kept for continuity with round 1",
                         "bytes_per_step": step_bytes, "achieved": model_achieved, "frac": model_achieved / peak}},
        "cpu_baseline": {"value": cpu["value"], "unit": "iters/s", "cores": cpu["threads"], "kind": "port",
                         "sample": f"{cpu['steps']} steps of the same workload on the oracle, one OS thread per "
                                   f"agent ({cpu['threads']} threads, {cpu['cores']} host cores); {cpu['build']}",
                         "iterate_true_median_us": cpu.get("iterate_true_median_us")},
        "async_mode": async_mode,
        "hbm_bound_regime": hbm_regime,
        "hbm_bound_regime_8d_generator": hbm_regime_8d,
        "reference_wrapper": wrapper,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

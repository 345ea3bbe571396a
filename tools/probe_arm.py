"""One robot alone on the GPU (its neighbour lives in the CPU oracle): iterate(true) latency with and without armed
launches (Agent::maybe_arm).  usage: python tools/probe_arm.py"""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import ctypes as C
    import numpy as np
    from dpgo_ros_b200 import agent as gpu, datasets, capi
    from oracle import binding as orc
    kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50, rel_change_tol=0.0, max_num_iters=10 ** 9)
    pb = datasets.load_g2o_problem("sphere2500", 8)
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    me = agents[3]
    for a in agents:
        if a is not me:
            a.close()
    oteam = orc.OracleTeam(pb, **kw)
    L = capi.lib()
    L.dpgo_b200_debug_host_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    out = (C.c_double * 4)()
    yl = datasets.fixed_lifting_matrix(5)

    def feed():   # neighbours' public poses (regular + auxiliary) from the oracle
        for nb in me.getNeighbors():
            frames = np.array(sorted({int(f) for f in me_frames[nb]}), dtype=np.int32)
            X, Y = oteam.get_x(nb), oteam.get_x(nb, 1)
            me.updateNeighborPoses(nb, frames, np.ascontiguousarray(np.stack([X[:, 4 * f:4 * f + 4].T for f in frames])), False)
            me.updateNeighborPoses(nb, frames, np.ascontiguousarray(np.stack([Y[:, 4 * f:4 * f + 4].T for f in frames])), True)
    m = pb.robot_measurements(3)
    me_frames = {}
    for e in np.nonzero(m.r1 != m.r2)[0]:
        o, f = (int(m.r2[e]), int(m.p2[e])) if int(m.r1[e]) == 3 else (int(m.r1[e]), int(m.p1[e]))
        me_frames.setdefault(o, set()).add(f)
    tt = []
    for it in range(8 * 60):
        sel = it % 8
        if sel == 3:
            feed()
            t0 = time.perf_counter()
            ok = me.iterate(True)
            tt.append(time.perf_counter() - t0)
            assert ok
        else:
            me.iterate(False)
            for nb in me.getNeighbors():
                me.getSharedPoseDictWithNeighbor(nb, False)
        oteam.run(1, stop_on_terminate=False)
    print("iterate(true) median us", float(np.median(tt[5:])) * 1e6, "min", min(tt[5:]) * 1e6, flush=True)
    err = np.linalg.norm(me.getX() - oteam.get_x(3)) / np.linalg.norm(oteam.get_x(3))
    print("rel. difference of robot 3 to the oracle team", err)
else:
    for env in ({"DPGO_B200_NO_ARM": "1"}, {}, {"DPGO_B200_NO_ARM": "1"}, {}):
        e = dict(os.environ); e.update(env)
        print("==", env or "armed launches", flush=True)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=200)
        print(r.stdout[-400:], r.stderr[-600:], flush=True)

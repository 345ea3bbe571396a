"""RTR through the per-robot C ABI on small agents (tunnels) and mid-size ones (torus3D): where one iterate(true) spends its
time -- launch call, launch -> result, kernel (CUDA events), and the same per tCG iteration."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DPGO_B200_TIME_LAUNCHES"] = "1"
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets, capi

L = capi.lib()
L.dpgo_b200_debug_host_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
for name, pb, kw in (("tunnels/8", datasets.load_tunnels_problem(), dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2, cost_type=5, gnc_barc=3.0)),
                     ("torus3D/4 r=6", datasets.load_g2o_problem("torus3D", 4), dict(r=6, method=0, gradnorm_tol=0.5, rel_change_tol=0.2)),
                     ("sphere2500/5", datasets.load_g2o_problem("sphere2500", 5), dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.2))):
    t0 = time.perf_counter()
    _, agents = gpu.make_team(pb, colocate=False, **kw)
    gpu.exchange_host(agents, accel=False)
    t1 = time.perf_counter()
    agents[0].iterate(True)   # builds everything for robot 0
    t2 = time.perf_counter()
    for a in agents[1:]:
        a.iterate(True)
    gpu.exchange_host(agents, accel=False)
    print(f"== {name}: setup {t1 - t0:.3f} s, first iterate(true) of robot 0 {1e3 * (t2 - t1):.1f} ms (n = {pb.n[0]})")
    out = (C.c_double * 4)()
    tot = dict(calls=0, wall=0.0, launch=0.0, result=0.0, kern=0.0, tcg=0, outer=0)
    for sweep in range(6):
        for a in agents:
            L.dpgo_b200_debug_host_profile(a.h, out, 1)
            t = time.perf_counter()
            a.iterate(True)
            dt = time.perf_counter() - t
            L.dpgo_b200_debug_host_profile(a.h, out, 1)
            o = a.localOptResult()
            tot["calls"] += 1; tot["wall"] += dt; tot["launch"] += out[0]; tot["result"] += out[1]; tot["kern"] += out[3]
            if o is not None:
                tot["tcg"] += o.tcg_iters; tot["outer"] += o.rtr_outer_iters
            gpu.exchange_host(agents, accel=False, only=[a.id])
    c = tot["calls"]
    print(f"   per iterate(true): wall {1e6 * tot['wall'] / c:.1f} us | launch call {1e6 * tot['launch'] / c:.1f} | launch->result {1e6 * tot['result'] / c:.1f} | kernel (events) {1e6 * tot['kern'] / c:.1f}"
          f" | tCG iterations {tot['tcg'] / c:.1f}, outer {tot['outer'] / c:.1f}"
          + (f" | kernel us per tCG iteration {1e6 * tot['kern'] / max(1, tot['tcg']):.2f}" if tot["tcg"] else ""))
    for a in agents:
        a.close()

"""Phase timeline of the persistent kernel + barrier micro-benchmark (GPU diagnostics)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets, capi

L = capi.lib()
L.dpgo_b200_debug_barrier_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
L.dpgo_b200_debug_team_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
for grid in (148, 64, 16, 1):
    for mode, nm in ((0, "barrier"), (1, "reduce2")):
        ms = C.c_float()
        rc = L.dpgo_b200_debug_barrier_bench(0, grid, 2000, mode, C.byref(ms))
        print(f"grid {grid:4d} {nm}: {ms.value*1e3/2000:.2f} us each (rc {rc})")

pb = datasets.load_g2o_problem("sphere2500", 8)
kw = dict(r=5, method=1, rgd_stepsize=0.2, rgd_use_preconditioner=1, acceleration=1, restart_interval=50, rel_change_tol=0.0, max_num_iters=10**9)
team, agents = gpu.make_team(pb, **kw)
team.run(3000, stop_on_terminate=False)
names = ["start", "nesterov", "barrier", "grad", "reduce", "rgdstep", "reduce", "grad2", "reduce"]
for cta in (0, 73, 147):
    iters = 12
    buf = (C.c_longlong * (iters * 16 + 32))()
    rc = L.dpgo_b200_debug_team_profile(team.h, iters, cta, buf)
    dbg = np.array(buf[iters * 16:])
    print('   grad marks 0..9 (cycles from mark 0):', [int(x - dbg[0]) for x in dbg[:10]])
    print('   warp-1 nesterov marks 24..27 (from dense mark 19):', [int(dbg[k] - dbg[19]) for k in (24, 25, 26, 27)])
    print('   dense marks 20,16,17,18,19,21,22,23 (from 20):', [int(dbg[k] - dbg[20]) for k in (20, 16, 17, 18, 19, 21, 22, 23)])
    a = np.array(buf[:iters * 16]).reshape(iters, 16)
    d = np.diff(a[:, :9], axis=1)
    print(f"cta {cta}: cycles per segment (median over {iters} iters), rc={rc}")
    for k in range(8):
        print(f"   {names[k]:9s}-> {names[k+1]:9s} {np.median(d[:, k]):9.0f}")
    print("   iteration total", np.median(a[1:, 0] - a[:-1, 0]))

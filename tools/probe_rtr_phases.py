"""clock64 timeline of one RTR solve inside the persistent kernel (marks of rtr_solve, team_run.cuh; CTA 0):
entry | grad | reduce | per outer iteration: precond | reduce | per tCG iteration: hess_dir | reduce | [precond_cg | reduce] |
then: (mark) | optional barrier | grad(cand) | hess(eta) | dot | reduce."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dpgo_ros_b200 import agent as gpu, datasets, capi

L = capi.lib()
L.dpgo_b200_debug_team_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
L.dpgo_b200_debug_team_marks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
cases = (("tunnels/8", datasets.load_tunnels_problem(), dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.0, max_num_iters=10**9)),
         ("sphere2500/5", datasets.load_g2o_problem("sphere2500", 5), dict(r=5, method=0, gradnorm_tol=0.5, rel_change_tol=0.0, max_num_iters=10**9)),
         ("torus3D/4 r=6", datasets.load_g2o_problem("torus3D", 4), dict(r=6, method=0, gradnorm_tol=0.5, rel_change_tol=0.0, max_num_iters=10**9)))
for name, pb, kw in cases:
    team, agents = gpu.make_team(pb, **kw)
    N = len(agents)
    team.run(2 * N, stop_on_terminate=False)       # warm: every robot has solved twice; next selected robot is 0
    buf = (C.c_longlong * (16 + 32))()
    L.dpgo_b200_debug_team_profile(team.h, 1, 0, buf)
    marks = (C.c_longlong * 160)()
    L.dpgo_b200_debug_team_marks(team.h, marks)
    m = np.array(marks[:])
    k = int(np.argmax(m == 0)) if (m == 0).any() else len(m)
    m = m[:k]
    d = np.diff(m) / 1965.0
    print(f"== {name}: robot 0, n = {pb.n[0]}, {k} marks, {d.sum():.1f} us in rtr_solve")
    print("   us between marks:", " ".join(f"{x:.1f}" for x in d))
    for a in agents:
        a.close()

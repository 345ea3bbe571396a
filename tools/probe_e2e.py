import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DPGO_B200_DRIVER_PROFILE"] = "1"
from dpgo_ros_b200 import agent as gpu, datasets
import bench
pb = datasets.load_g2o_problem("sphere2500", 8)
_, agents = gpu.make_team(pb, colocate=False, **bench.CONFIG2)
gpu.exchange_host(agents, accel=True)
gpu.sync_driver_run(agents, 50, True)
sec, _ = gpu.sync_driver_run(agents, 2000, True)
print("e2e it/s", 2000 / sec)
# single-agent call latencies
a = agents[3]
for name, fn in (("iterate(false)", lambda: a.iterate(False)), ("iterate(true)", lambda: a.iterate(True))):
    t = time.perf_counter()
    for _ in range(200): fn()
    print(name, (time.perf_counter() - t) / 200 * 1e6, "us")
import ctypes as C
L = a.L
L.dpgo_b200_debug_host_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
out = (C.c_double * 4)()
for name, flag in (("iterate(false)", False), ("iterate(true)", True)):
    L.dpgo_b200_debug_host_profile(a.h, out, 1)
    t = time.perf_counter()
    for _ in range(300): a.iterate(flag)
    dt = (time.perf_counter() - t) / 300 * 1e6
    L.dpgo_b200_debug_host_profile(a.h, out, 1)
    print(f"{name}: total {dt:.1f} us | launch call {out[0]/out[2]*1e6:.1f} | until result {out[1]/out[2]*1e6:.1f} | other host {dt - (out[0]+out[1])/out[2]*1e6:.1f} | kernel (events) {out[3]/out[2]*1e6:.1f}")
# kernel-only time of a 1-agent team iteration
from dpgo_ros_b200 import capi

"""e2e through the per-robot C ABI with host buffers (dpgo_b200_sync_driver_run), armed launches on / off."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    os.environ["DPGO_B200_DRIVER_PROFILE"] = "1"
    import numpy as np
    from dpgo_ros_b200 import agent as gpu, datasets
    from oracle import binding as orc
    import bench
    pb = datasets.load_g2o_problem("sphere2500", 8)
    _, agents = gpu.make_team(pb, colocate=False, **bench.CONFIG2)
    gpu.exchange_host(agents, accel=True)
    gpu.sync_driver_run(agents, 50, True)
    sec, _ = gpu.sync_driver_run(agents, 4000, True)
    print("e2e it/s", 4000 / sec, "us/iter", sec / 4000 * 1e6, flush=True)
    # parity of the whole driven run against the oracle
    oteam = orc.OracleTeam(pb, **bench.CONFIG2)
    oteam.run(4050, threads=8, stop_on_terminate=False)
    err = max(np.linalg.norm(a.getX() - oteam.get_x(a.id)) / np.linalg.norm(oteam.get_x(a.id)) for a in agents)
    print("max rel. difference to the oracle after 4050 iterations:", err, flush=True)
    for a in agents:
        a.close()
else:
    for env in ({"DPGO_B200_NO_ARM": "1"}, {}, {"DPGO_B200_NO_ARM": "1"}, {}):
        e = dict(os.environ); e.update(env)
        print("==", env or "armed launches", flush=True)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=300)
        print(r.stdout[-600:], r.stderr[-700:], flush=True)

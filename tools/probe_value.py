"""us per global iteration of the persistent kernel on config 2 (GPU diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpgo_ros_b200 import agent as gpu, datasets
import bench
pb = datasets.load_g2o_problem("sphere2500", 8)
team, agents = gpu.make_team(pb, **bench.CONFIG2)
team.run(3000, stop_on_terminate=False)
best = []
for _ in range(5):
    res = team.run(4000, stop_on_terminate=False)
    best.append(res.device_ms * 1e3 / 4000)
print("us/iter:", " ".join(f"{b:.2f}" for b in best))

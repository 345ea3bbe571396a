"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`: the CPU oracle on the host
cores) prints ONE JSON line with the agreed keys, and the algorithmic-bytes accounting matches DESIGN.md 3.4."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "40",
                          "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rbcd_iters_per_sec" and d["unit"] == "iters/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 40
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "sphere2500" in d["config"]["workload"]


def test_reference_arm_on_other_ranks_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_algorithmic_bytes_accounting():
    sys.path.insert(0, ROOT)
    import bench
    from dpgo_ros_b200 import datasets
    pb = datasets.load_g2o_problem("sphere2500", 8)
    step, grad = bench.algorithmic_bytes(pb, 5)
    # B_grad: ~199 KB per agent; one step = 2 B_grad + the 12.5 MB dense preconditioner + Nesterov of all
    assert 190e3 < grad < 210e3
    n = pb.n[0]
    assert abs(step - (2 * grad + (4 * n) ** 2 * 8 + 1.6e6)) < 0.6e6
    assert 14.0e6 < step < 15.2e6

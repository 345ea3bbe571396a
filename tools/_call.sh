mkdir -p gpurun_out/c44
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c44/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c44/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/c44/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c44/bench_n1_s20.json 2> gpurun_out/c44/bench_n1_s20.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c44/bench_ref.json 2> gpurun_out/c44/bench_ref.err; echo "ref rc=$?"
kill $SMI
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c44/bench_n1_s20.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"], "cpu", d["cpu_baseline"]["value"])
rw = d["reference_wrapper"]
for k, v in rw.items():
    if isinstance(v, dict): print(k, v["b200"]["wall_seconds"], v["oracle"]["wall_seconds"], v.get("speedup_wall"), v.get("speedup_library"))
r = json.loads(open("gpurun_out/c44/bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"))
PY

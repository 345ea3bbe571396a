mkdir -p gpurun_out/c52
SECONDS=0
timeout 400 python bench.py > gpurun_out/c52/bench_default.json 2> gpurun_out/c52/bench_default.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c52/bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
PY

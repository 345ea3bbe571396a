mkdir -p gpurun_out/c33
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/c33/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c33/bench_n1_s20.json 2> gpurun_out/c33/bench_n1_s20.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c33/bench_ref.json 2> gpurun_out/c33/bench_ref.err; echo "ref rc=$?"
kill $SMI
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c33/bench_n1_s20.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "cold_l2_ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "roofline", d["roofline"]["frac"])
print(json.dumps(d["hbm_bound_regime"])[:1800])
print(json.dumps(d["reference_wrapper"])[:3000])
r = json.loads(open("gpurun_out/c33/bench_ref.json").read().strip().splitlines()[-1])
print({k: r.get(k) for k in ("impl", "value", "ms_per_step")}, r.get("e2e"))
PY

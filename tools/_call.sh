mkdir -p gpurun_out/c39
timeout 300 python -m pytest tests/test_gpu_parallel.py -m gpu -q -k "config5_full_size or edge_record" 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c39/bench_n1_s20.json 2> gpurun_out/c39/bench_n1_s20.err; echo "bench rc=$?"; tail -3 gpurun_out/c39/bench_n1_s20.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c39/bench_n1_s20.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
print(json.dumps(d["hbm_bound_regime_8d_generator"])[:2500])
h = d["hbm_bound_regime"]; print(h["ms_per_iterate"], h["preconditioner_build_s"], h["frac"], h.get("cpu_beside"))
PY

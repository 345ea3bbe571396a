mkdir -p gpurun_out/c51
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c51/bench_n1_s20.json 2> gpurun_out/c51/bench_n1_s20.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c51/bench_n1_s20.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
rw = d["reference_wrapper"]
for k, v in rw.items():
    if isinstance(v, dict): print(k, v["b200"]["wall_seconds"], v["oracle"]["wall_seconds"], round(v.get("speedup_wall"), 2), round(v.get("speedup_library"), 2))
print(d["hbm_bound_regime"]["ms_per_iterate"], d["hbm_bound_regime"]["preconditioner_build_s"], d["hbm_bound_regime_8d_generator"]["ms_per_iterate"])
PY

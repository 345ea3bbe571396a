mkdir -p gpurun_out/c21
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/c21/clocks.csv &
SMI=$!
# 1. the bench line itself (full), for profiles/bench_r2.json
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c21/bench_n1_s20.json 2> gpurun_out/c21/bench_n1_s20.err; echo "bench rc=$?"
# 2. launch list of the same command (secondary objects skipped: they launch thousands of set-up kernels)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c21/launches_r2.csv python bench.py --steps 20 --warmup 5 --no-secondary --e2e-steps 40 --cpu-steps 40 > gpurun_out/c21/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
# 3. ncu --set full: the persistent kernel (one launch of 400 iterations) and the edge-record gradient
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_team_run -s 2 -c 1 -o gpurun_out/c21/team_run_r2 python tools/ncu_target.py 400 rgd > gpurun_out/c21/ncu_target.log 2>&1; echo "ncu team_run rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edge_grad -s 2 -c 1 -o gpurun_out/c21/edge_grad_r2 python tools/ncu_edge_target.py > gpurun_out/c21/ncu_edge.log 2>&1; echo "ncu edge_grad rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:k_team_run -s 3 -c 1 -o gpurun_out/c21/team_run_rtr_r2 python tools/ncu_target.py 24 rtr > gpurun_out/c21/ncu_target_rtr.log 2>&1; echo "ncu rtr rc=$?"
kill $SMI
ls -la gpurun_out/c21 | head -20
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c21/bench_n1_s20.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "cold_l2_ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "roofline", d["roofline"]["frac"])
print(json.dumps(d["hbm_bound_regime"])[:1500])
print(json.dumps(d["reference_wrapper"])[:2500])
PY

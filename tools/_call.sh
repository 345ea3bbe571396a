mkdir -p gpurun_out/c48
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c48/bench_n4_s20.json 2> gpurun_out/c48/bench_n4_s20.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c48/bench_n4_s20.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "bit_identical_to_single_team")}, "e2e", d["e2e"]["value"], d["e2e"].get("final_cost_2f"), "async", d.get("async_mode", {}).get("ticks_per_s"))
PY

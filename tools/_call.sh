mkdir -p gpurun_out/c41
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c41/bench_n8_s20.json 2> gpurun_out/c41/bench_n8_s20.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c41/bench_n8_s20.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "bit_identical_to_single_team", "gpu_launches")}, "e2e", d["e2e"]["value"], "async", d.get("async_mode", {}).get("ticks_per_s"))
PY

mkdir -p gpurun_out/c46
timeout 300 python -m pytest tests/test_gpu_fabric_multi.py tests/test_gpu_fabric.py -m gpu -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sync_driver or armed or standalone" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c46/bench_n2_s20.json 2> gpurun_out/c46/bench_n2_s20.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c46/bench_n2_s20.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "bit_identical_to_single_team")}, "e2e", d["e2e"]["value"], d["e2e"].get("final_cost_2f"), "async", d.get("async_mode", {}).get("ticks_per_s"))
PY

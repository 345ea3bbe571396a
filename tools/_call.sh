mkdir -p gpurun_out/c43
timeout 400 python tools/probe_e2e.py > gpurun_out/c43/probe_e2e.txt 2>&1; cat gpurun_out/c43/probe_e2e.txt | tail -24
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sync_driver or armed or standalone" 2>&1 | tail -3

mkdir -p gpurun_out/c36
timeout 400 python tools/probe_e2e.py > gpurun_out/c36/probe_e2e.txt 2>&1; cat gpurun_out/c36/probe_e2e.txt | tail -24

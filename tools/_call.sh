mkdir -p gpurun_out/c37
timeout 120 python tools/probe_rtr_phases.py > gpurun_out/c37/rtr_phases.txt 2>&1; tail -6 gpurun_out/c37/rtr_phases.txt
timeout 120 python tools/probe_rtr.py > gpurun_out/c37/probe_rtr.txt 2>&1; grep "per iterate" gpurun_out/c37/probe_rtr.txt
timeout 300 python tools/bench_config5.py > gpurun_out/c37/config5.json 2> gpurun_out/c37/config5.err; tail -c 900 gpurun_out/c37/config5.json
timeout 800 python -m pytest tests -m gpu -q -x > gpurun_out/c37/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c37/pytest.log

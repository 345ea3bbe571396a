mkdir -p gpurun_out/c17
timeout 400 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/c17/pytest.log 2>&1; tail -6 gpurun_out/c17/pytest.log | cut -c1-400
timeout 300 python tools/bench_config5.py --iters 6 > gpurun_out/c17/config5.json 2> gpurun_out/c17/config5.err; tail -c 1800 gpurun_out/c17/config5.json; tail -3 gpurun_out/c17/config5.err

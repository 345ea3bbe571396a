mkdir -p gpurun_out/c11
timeout 600 python tools/probe_e2e.py 2>&1 | tail -40
python -m pytest tests/test_gpu_parity.py tests/test_shim.py tests/test_gpu_edge.py -m gpu -q --timeout 900 -x > gpurun_out/c11/pytest.log 2>&1; tail -6 gpurun_out/c11/pytest.log | cut -c1-300

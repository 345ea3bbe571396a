mkdir -p gpurun_out/c12
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 > gpurun_out/c12/pytest.log 2>&1; tail -8 gpurun_out/c12/pytest.log | cut -c1-400

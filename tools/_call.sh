mkdir -p gpurun_out/c2
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2/pytest.log
tail -25 gpurun_out/c2/pytest.log

mkdir -p gpurun_out/c40
REPS=1 timeout 300 compute-sanitizer --tool memcheck ./tools/check_dense_inverse 32 96 448 1248 > gpurun_out/c40/memcheck_inverse.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/c40/memcheck_inverse.txt
REPS=1 timeout 400 compute-sanitizer --tool racecheck ./tools/check_dense_inverse 96 448 > gpurun_out/c40/racecheck_inverse.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/c40/racecheck_inverse.txt
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rtr_team or rtr_single or rtr_accelerated_matches" > gpurun_out/c40/memcheck_rtr.txt 2>&1; echo "memcheck rtr rc=$?"; tail -5 gpurun_out/c40/memcheck_rtr.txt
REPS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_leaf -s 12 -c 1 -o gpurun_out/c40/leaf2 ./tools/check_dense_inverse 1248 > /dev/null 2>&1; echo "ncu leaf rc=$?"

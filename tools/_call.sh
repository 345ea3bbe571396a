mkdir -p gpurun_out/c47
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm --launch-skip 612 --launch-count 14 -o gpurun_out/c47/gemm10016 ./tools/check_dense_inverse 10016 > gpurun_out/c47/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c47/ncu.log

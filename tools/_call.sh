mkdir -p gpurun_out/c49
for t in 74 148 296 592 100000; do echo "== big tiles from $t"; DPGO_B200_GEMM_BIG_TILES=$t ./tools/check_dense_inverse 2016 2528 5024 10016 | grep "N="; done 2>&1 | tee gpurun_out/c49/tile_policy.txt

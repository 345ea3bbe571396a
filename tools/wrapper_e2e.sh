#!/bin/bash
# e2e through the reference's OWN wrapper (oracle/_ref, built by oracle/Makefile.ref): the demo launch files run in one
# process on the ROS stand-in, once on libdpgo_b200.so and once on the CPU oracle, on this box.  Wall-clock time between
# the first UPDATE command and TERMINATE is in round_wall_seconds of every JSON.  usage: tools/wrapper_e2e.sh OUTDIR
out=${1:-gpurun_out/wrapper_e2e}; mkdir -p "$out"; cd "$(dirname "$0")/.."
run() { tag=$1; shift; for be in b200 oracle; do timeout 25 ./oracle/_ref/dpgo_ros_inproc_$be "$@" --out "$out/${tag}_$be.json" --log 0 2>"$out/${tag}_$be.err" || echo "$tag $be rc=$?"; done; }
run torus3D_4_r6        --robots 4 --g2o data/torus3D.g2o    --preset dpgo_demo --param relaxation_rank=6
run sphere2500_8        --robots 8 --g2o data/sphere2500.g2o --preset dpgo_demo
run sphere2500_5_odom   --robots 5 --g2o data/sphere2500.g2o --preset dpgo_demo --param local_initialization_method=Odometry
run sphere2500_5_odom_acc --robots 5 --g2o data/sphere2500.g2o --preset dpgo_demo --param local_initialization_method=Odometry --param acceleration=true
run tunnels_8_gnc       --robots 8 --measurements data/tunnels --preset gnc_demo
python - "$out" <<'PY'
import json, glob, os, sys
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.json"))):
    d = json.load(open(f))
    it, w = d["round_iterations"], d["round_wall_seconds"]
    print(os.path.basename(f), d["backend"], it, w, "%.0f iters/s" % (it[0] / w[0]) if it and w else "-", "launches", d["kernel_launches"])
PY
nproc

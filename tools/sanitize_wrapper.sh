#!/bin/bash
# Builds the reference wrapper + DPGO:: shim + ROS stand-in + CPU oracle back end under AddressSanitizer/UBSan and under
# ThreadSanitizer (reference sources read from $REF, objects under /tmp) and runs the wrapper scenarios: two rounds, GNC with
# persisting rejections, RECOVER, robust initialisation, the asynchronous demo (the shim's optimisation threads next to the
# wrapper's callbacks).  Prints one line per run: sanitizer reports found / exit code.  usage: tools/sanitize_wrapper.sh
set -u
R="$(cd "$(dirname "$0")/.." && pwd)"; REF=${REF:-/root/reference}
make -s -f "$R/oracle/Makefile.ref" "$R/oracle/_ref/build/gen/.stamp" || exit 1
INC="-I$R/include -I$R/tests/cpp/ros_stub/include -I$R/oracle/_ref/build/gen -I$REF/include"
build() {  # $1 = tag, $2 = sanitizer flags
  d=/tmp/dpgo_san_$1; mkdir -p $d; F="-std=c++17 -O1 -g $2 -fno-omit-frame-pointer -pthread -w"
  ( g++ $F $INC -c $REF/src/utils.cpp -o $d/utils.o & g++ $F $INC -c $REF/src/PGOAgentROS.cpp -o $d/a.o &
    g++ $F $INC -Dmain=dpgo_ros_agent_main -c $REF/src/PGOAgentROSNode.cpp -o $d/n.o &
    g++ $F $INC -Dmain=dpgo_ros_dataset_publisher_main -c $REF/src/PGODatasetPublisherNode.cpp -o $d/p.o &
    g++ $F $INC -c $R/tests/cpp/ros_stub/inproc_launch.cpp -o $d/l.o & g++ $F -c $R/oracle/abi_on_oracle.cpp -o $d/abi.o &
    g++ $F -c $R/oracle/dpgo_oracle.cpp -o $d/o.o & wait )
  g++ $F -o $d/inproc $d/*.o || exit 1
}
run() {  # $1 = tag, rest = arguments
  tag=$1; shift
  out=$(timeout 1200 /tmp/dpgo_san_$tag/inproc "$@" --out /tmp/dpgo_san_$tag/r.json --log 0 2>&1); rc=$?
  n=$(echo "$out" | grep -cE "ERROR: AddressSanitizer|runtime error:|WARNING: ThreadSanitizer|LeakSanitizer")
  echo "$tag reports=$n rc=$rc :: $*" | sed "s#$R/##g"
}
D=$R/data
for tag in asan tsan; do
  [ $tag = asan ] && build asan "-fsanitize=address,undefined" || build tsan "-fsanitize=thread"
  run $tag --robots 2 --g2o $D/smallGrid3D.g2o --preset dpgo_demo --rounds 2
  run $tag --robots 8 --measurements $D/tunnels --preset gnc_demo --param weight_convergence_threshold=0.5 --rounds 2 --max-sim-seconds 400
  run $tag --robots 5 --g2o $D/sphere2500.g2o --preset dpgo_demo --param local_initialization_method=Odometry --param enable_recovery=true --disconnect 3@36.2
  run $tag --robots 8 --measurements $D/tunnels --preset gnc_demo --param local_initialization_method=GNC_TLS
  run $tag --robots 5 --g2o $D/sphere2500.g2o --preset asapp_demo --realtime 5 --run-sim-seconds 45
done

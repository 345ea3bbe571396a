mkdir -p gpurun_out/c27
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/c27/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c27/pytest.log
for i in 1 2 3; do
timeout 25 ./oracle/_ref/dpgo_ros_inproc_b200 --robots 5 --g2o data/sphere2500.g2o --preset dpgo_demo --param local_initialization_method=Odometry --out gpurun_out/c27/sphere5_$i.json --log 0 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c27/sphere5_*.json")):
    d = json.load(open(f))
    p = d.get("round_library_profile", [{}])[0]
    print(f.split("/")[-1], d["round_iterations"], "wall %.4f lib %.4f" % (d["round_wall_seconds"][0], d["round_library_seconds"][0]),
          {k: round(v[0], 4) for k, v in p.items() if k.startswith(".")})
PY

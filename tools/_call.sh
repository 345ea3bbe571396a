mkdir -p gpurun_out/c35
timeout 800 python -m pytest tests -m gpu -q -x > gpurun_out/c35/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c35/pytest.log
timeout 200 tools/wrapper_e2e.sh gpurun_out/c35/wrapper_e2e > gpurun_out/c35/wrapper_e2e.txt 2>&1; cat gpurun_out/c35/wrapper_e2e.txt
python - <<'PY'
import json
for n in ("tunnels_8_gnc_b200", "torus3D_4_r6_b200", "sphere2500_5_odom_b200"):
    d = json.load(open("gpurun_out/c35/wrapper_e2e/%s.json" % n))
    print(n, d.get("round_wall_seconds"), d.get("round_library_seconds"))
    for p in d.get("round_library_profile", []):
        print("   ", json.dumps({k: round(v[0], 4) for k, v in p.items() if v[0] > 1e-3}))
PY

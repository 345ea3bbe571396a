#!/bin/bash
# Regenerates the iteration counts of tests/golden/reference_wrapper.json: every run listed there through the reference's
# unmodified wrapper on the CPU oracle back end (oracle/_ref, built by `make -f oracle/Makefile.ref` where /root/reference
# exists).  Prints "name iterations"; edit the JSON by hand if a number legitimately changes.
cd "$(dirname "$0")/../.." || exit 1
python3 - <<'PY'
import json, os, subprocess, tempfile
g = json.load(open("tests/golden/reference_wrapper.json"))
for run in g["runs"]:
    args = [os.path.join("data", a) if a.endswith(".g2o") or a == "tunnels" else a for a in run["args"]]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "r.json")
        subprocess.run(["oracle/_ref/dpgo_ros_inproc_oracle", *args, "--out", out, "--log", "0"], check=True)
        print(run["name"], json.load(open(out))["round_iterations"][0])
PY

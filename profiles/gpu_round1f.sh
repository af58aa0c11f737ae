#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_joint.py -x -q -k "not full_size" > gpurun_out/pytest_pool.log 2>&1; echo "pytest rc=$?" > gpurun_out/status_f.txt
for i in 1 2 3; do timeout 40 python bench.py --no-e2e --no-cpu-baseline --steps 30 > gpurun_out/pool_$i.json 2>> gpurun_out/pool.err; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/pool_*.json")):
    d = json.load(open(f)); print(f, round(d["value"], 1), d["clocks"]["samples"])
PY
cat gpurun_out/status_f.txt; tail -2 gpurun_out/pytest_pool.log

#!/bin/bash
# NOTE: measured an EXPERIMENT build that is not in the tree any more (knobs / variants removed after the
# measurement; results in profiles/r02/direct_index_ab.jsonl or tune_gather.jsonl, discussion in DESIGN.md 4.3).
# Source-only list words (4 bytes) for the first 16 slots, weights derived by the gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/src32_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/src32_pytest.log
tail -3 gpurun_out/src32_pytest.log
B="--no-cpu-baseline --no-e2e --steps 20"
for i in 1 2; do timeout 120 python bench.py $B > gpurun_out/src32_$i.json 2>> gpurun_out/src32.err; done
timeout 120 python bench.py $B --motion B > gpurun_out/src32_motionB.json 2>> gpurun_out/src32.err
timeout 120 python bench.py $B --motion C > gpurun_out/src32_motionC.json 2>> gpurun_out/src32.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/src32_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/src32.err

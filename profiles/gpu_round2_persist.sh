#!/bin/bash
# NOTE: measured an EXPERIMENT build that is not in the tree any more (knobs / variants removed after the
# measurement; results in profiles/r02/direct_index_ab.jsonl or tune_gather.jsonl, discussion in DESIGN.md 4.3).
# Persistent gather (fixed grid of k CTAs per SM): does leaving a quarter of the register file to the side stream pay?
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 20"
for k in 0 4 3; do
  SLR_GATHER_PERSIST=$k timeout 120 python bench.py $B > gpurun_out/persist_$k.json 2>> gpurun_out/persist.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/persist_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/persist.err

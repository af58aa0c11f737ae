#!/bin/bash
# NOTE: measured an EXPERIMENT build that is not in the tree any more (knobs / variants removed after the
# measurement; results in profiles/r02/direct_index_ab.jsonl or tune_gather.jsonl, discussion in DESIGN.md 4.3).
# Does leaving room beside the gather (3 CTAs/SM) let insert_kernel of the next batch overlap it?
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 20"
for room in 0 60000 74000; do
  SLR_GATHER_ROOM=$room timeout 120 python bench.py $B > gpurun_out/room_$room.json 2>> gpurun_out/room.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/room_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/room.err

#!/bin/bash
mkdir -p gpurun_out
for b in 8 12 16 20 24; do
  timeout 120 python bench.py --no-cpu-baseline --no-e2e --steps 20 --batch $b > gpurun_out/batch_$b.json 2>> gpurun_out/batch.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/batch_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/batch.err

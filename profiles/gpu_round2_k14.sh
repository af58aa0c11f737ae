#!/bin/bash
# NOTE: measured an EXPERIMENT build (-DSLR_GATHER_K14=1: a 14-slot variant of the gather between 12 and 16).
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 20"
for i in 1 2; do
  timeout 120 python bench.py $B > gpurun_out/k14_base_$i.json 2>> gpurun_out/k14.err
  SLR_LIB=gpurun_variants/libslr_splat_k14.so timeout 120 python profiles/bench_with_lib.py $B > gpurun_out/k14_k14_$i.json 2>> gpurun_out/k14.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/k14_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/k14.err

#!/bin/bash
# A/B of compile-time variants of the library (prebuilt by profiles/build_variants.py) against the product build:
#   bash profiles/gpu_round2_tune.sh <variant> ...      -> gpurun_out/tune_*.json + a summary
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 30"
for i in 1 2; do
  timeout 90 python bench.py $B > gpurun_out/tune_base_$i.json 2>> gpurun_out/tune.err
  for v in "$@"; do
    SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 90 python profiles/bench_with_lib.py $B > gpurun_out/tune_${v}_$i.json 2>> gpurun_out/tune.err
  done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/tune_*_[0-9].json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/tune.err

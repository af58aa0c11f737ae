#!/bin/bash
# NOTE: measured EXPERIMENT builds (-DSLR_GATHER_K10=1: a 10-slot gather variant; -DSLR_INSERT_THREADS=128 / 64).
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 20"
timeout 120 python bench.py $B > gpurun_out/micro_base.json 2>> gpurun_out/micro.err
for v in k10 ins128 ins64; do
  SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 120 python profiles/bench_with_lib.py $B > gpurun_out/micro_$v.json 2>> gpurun_out/micro.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/micro_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/micro.err

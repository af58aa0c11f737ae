#!/bin/bash
# (historical record: the variant libraries were built locally into gpurun_variants/ with
#  SLR_EXPAND_SMEM_SLOTS / SLR_EXPAND_MINBLOCKS / SLR_GATHER_FINE_BUCKETS defines that have since been removed)
# A/B of compile-time variants prebuilt into gpurun_variants/ (no nvcc on the GPU box), alternating runs.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 30"
for i in 1 2; do
  timeout 60 python bench.py $B > gpurun_out/v_base_$i.json 2>> gpurun_out/v.err
  for v in e13x8 e12x8 e13x7 buckets; do
    SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 60 python profiles/bench_with_lib.py $B > gpurun_out/v_${v}_$i.json 2>> gpurun_out/v.err
  done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/v_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), {k: round(v, 4) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# NOTE: measured an EXPERIMENT build that is not in the tree any more (knobs / variants removed after the
# measurement; results in profiles/r02/direct_index_ab.jsonl or tune_gather.jsonl, discussion in DESIGN.md 4.3).
# Gather knobs again on the direct index: loads in flight, CTA shape
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 20"
timeout 120 python bench.py $B > gpurun_out/tune2_base.json 2>> gpurun_out/tune2.err
SLR_GATHER_SHAPE=1x4 timeout 120 python bench.py $B > gpurun_out/tune2_shape1x4.json 2>> gpurun_out/tune2.err
for v in loads6 loads10 loads12; do
  SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 120 python profiles/bench_with_lib.py $B > gpurun_out/tune2_$v.json 2>> gpurun_out/tune2.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/tune2_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/tune2.err

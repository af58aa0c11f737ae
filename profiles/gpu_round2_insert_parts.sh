#!/bin/bash
# NOTE: measured an EXPERIMENT build that is not in the tree any more (knobs / variants removed after the
# measurement; results in profiles/r02/direct_index_ab.jsonl or tune_gather.jsonl, discussion in DESIGN.md 4.3).
# Component isolation of insert_kernel (timing only, results invalid): x1 = no stores, x2 = no atomics, x3 = loads + arithmetic only
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 10"
for v in x2d16 x4 x2cg; do
  SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 120 python profiles/bench_with_lib.py $B > gpurun_out/parts_$v.json 2>> gpurun_out/parts.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/parts_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/parts.err

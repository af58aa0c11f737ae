#!/bin/bash
# Rebuild the library with different gather tuning knobs and bench each (run under gpurun).
mkdir -p gpurun_out
for cfg in "16 2" "16 3" "12 3" "12 2"; do
  set -- $cfg
  SLR_DEFINES="-DSLR_GATHER_DEPTH=$1 -DSLR_GATHER_MINBLOCKS=$2" python slr-sfs_b200/csrc/build.py --force > /dev/null
  python bench.py --steps 3 --warmup 3 --batch 12 --no-e2e --no-cpu-baseline > gpurun_out/sweep_d$1_b$2.json 2> gpurun_out/sweep_d$1_b$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_d$1_b$2.json"))
print("depth $1 minblocks $2:", round(d["value"],1), "frames/s; per-frame ms", {k: round(v,4) for k,v in d["roofline"]["all_kernels_ms_per_frame"].items()})
PY
done
python slr-sfs_b200/csrc/build.py --force > /dev/null

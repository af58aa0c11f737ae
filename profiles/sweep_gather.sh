#!/bin/bash
# Rebuild the library with different -D tuning knobs (see csrc/clip_gather.cu: SLR_GATHER_MINBLOCKS,
# SLR_EXPAND_MINBLOCKS) and bench each; run under gpurun.
# usage: sweep_gather.sh "<defines 1>" "<defines 2>" ...
mkdir -p gpurun_out
i=0
for defs in "$@"; do
  i=$((i+1))
  SLR_DEFINES="$defs" python slr-sfs_b200/csrc/build.py --force > /dev/null
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$i.json"))
print("[$defs]:", round(d["value"],1), "frames/s; per-frame ms", {k: round(v,4) for k,v in d["roofline"]["all_kernels_ms_per_frame"].items()})
PY
done
python slr-sfs_b200/csrc/build.py --force > /dev/null

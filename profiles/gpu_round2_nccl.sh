#!/bin/bash
# Does a narrower NCCL broadcast (fewer channels = fewer CTAs beside the gather) expose less of the exchange?
N=${1:-8}
mkdir -p gpurun_out
for ch in 2 6; do
  NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/nccl_ch${ch}_${N}gpu.json 2> gpurun_out/nccl_ch${ch}_${N}gpu.err
  echo "ch=$ch rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/nccl_ch*_${N}gpu.json")):
    try:
        d = json.load(open(f)); m = d["multi_gpu"]
        print(f, round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "exposed", round(m["exposed_broadcast_us_per_scene"], 1), "with", round(m["ms_per_step_with_broadcast"], 3), "without", round(m["ms_per_step_without_broadcast"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY

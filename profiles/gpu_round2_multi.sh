#!/bin/bash
# N-GPU bench (the driver's launch line), exchange inside the timed region, --check on
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --check --no-cpu-baseline > gpurun_out/multi_${N}gpu.json 2> gpurun_out/multi_${N}gpu.err
echo rc=$?
python - <<PY
import json
d = json.load(open("gpurun_out/multi_${N}gpu.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["all_kernels_ms_per_frame"], d.get("multi_gpu"), d.get("check"), d.get("e2e"))
PY
tail -3 gpurun_out/multi_${N}gpu.err

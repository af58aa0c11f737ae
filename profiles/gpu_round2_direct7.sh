#!/bin/bash
# Round 2, direct index v7 (moving blocks only)
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 30"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/direct7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/direct7_pytest.log
tail -3 gpurun_out/direct7_pytest.log
for i in 1 2; do
  timeout 120 python bench.py $B > gpurun_out/direct7_ldg_$i.json 2>> gpurun_out/direct7.err
done
timeout 120 python bench.py $B --motion B > gpurun_out/direct7_ldg_motionB.json 2>> gpurun_out/direct7.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 3 -c 1 -f -o gpurun_out/ncu_insert_kernel_v7 \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/direct7_ncu_full.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/direct7_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/direct7.err

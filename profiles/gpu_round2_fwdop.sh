#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fwdop_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fwdop_pytest.log
tail -5 gpurun_out/fwdop_pytest.log
timeout 300 python profiles/bench_forward_op.py > gpurun_out/forward_op.json 2> gpurun_out/forward_op.err
cat gpurun_out/forward_op.json; tail -3 gpurun_out/forward_op.err
for i in 1 2; do timeout 300 python bench.py --no-e2e --steps 20 > gpurun_out/fwdop_bench_$i.json 2>> gpurun_out/fwdop.err; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/fwdop_bench_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()}, "level0", d["level0"]["value"], "refgpu", d["reference_gpu"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/fwdop.err

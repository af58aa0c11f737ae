#!/bin/bash
# Round 2, direct index (insert_kernel) against the bin pipeline (SLR_GATHER_MODE=bins):
# GPU parity suite, alternating bench runs, ncu launch list of one step (single stream).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/direct_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/direct_pytest.log
tail -5 gpurun_out/direct_pytest.log
B="--no-cpu-baseline --no-e2e --steps 30"
for i in 1 2; do
  timeout 120 python bench.py $B > gpurun_out/direct_ldg_$i.json 2>> gpurun_out/direct.err
  SLR_GATHER_MODE=bins timeout 120 python bench.py $B > gpurun_out/direct_bins_$i.json 2>> gpurun_out/direct.err
done
timeout 120 python bench.py $B --motion B > gpurun_out/direct_ldg_motionB.json 2>> gpurun_out/direct.err
timeout 120 python bench.py $B --motion C > gpurun_out/direct_ldg_motionC.json 2>> gpurun_out/direct.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_direct_single_stream.csv \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/direct_ncu.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/direct_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/direct.err

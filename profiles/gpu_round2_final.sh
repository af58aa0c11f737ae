#!/bin/bash
# Round 2 final state on 1 GPU: parity suite, both bench arms (driver defaults), motions B / C, sweep, launch list,
# ncu --set full of the two hot kernels, SASS listings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 600 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
for m in B C; do
  timeout 120 python bench.py --no-cpu-baseline --no-e2e --steps 20 --motion $m > gpurun_out/final_bench_motion$m.json 2>> gpurun_out/final.err
done
SLR_GATHER_MODE=bins timeout 120 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/final_bench_bins.json 2>> gpurun_out/final.err
timeout 900 python profiles/sweep_configs.py > gpurun_out/final_sweep_configs.json 2>> gpurun_out/final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final_direct_single_stream.csv \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/final_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowgather_kernel -s 3 -c 1 -f -o gpurun_out/ncu_rowgather_final \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/final_ncu_full1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 3 -c 1 -f -o gpurun_out/ncu_insert_final \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/final_ncu_full2.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try:
        d = json.load(open(f)); r = d.get("roofline") or {}
        print(f, round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "live", r.get("frac"), "single", (r.get("single_stream") or {}).get("frac"),
              {k: round(v * 1000, 1) for k, v in (r.get("all_kernels_ms_per_frame") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/final_sweep_configs.json | cut -c1-400
tail -3 gpurun_out/final.err

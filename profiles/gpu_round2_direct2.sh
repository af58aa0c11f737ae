#!/bin/bash
# Round 2, direct index v2 (claims issued together): bench, list-depth variants, launch list, ncu --set full of insert_kernel.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 30"
timeout 600 python -m pytest tests/test_gpu_joint.py tests/test_gpu_bench_path.py tests/test_gpu_zz_clip_table.py -m gpu -x -q > gpurun_out/direct2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/direct2_pytest.log
tail -3 gpurun_out/direct2_pytest.log
for i in 1 2; do
  timeout 120 python bench.py $B > gpurun_out/direct2_ldg_$i.json 2>> gpurun_out/direct2.err
  for v in d32 d48 d64; do
    SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 120 python profiles/bench_with_lib.py $B > gpurun_out/direct2_${v}_$i.json 2>> gpurun_out/direct2.err
  done
done
timeout 120 python bench.py $B --motion B > gpurun_out/direct2_ldg_motionB.json 2>> gpurun_out/direct2.err
SLR_LIB=gpurun_variants/libslr_splat_d32.so timeout 120 python profiles/bench_with_lib.py $B --motion B > gpurun_out/direct2_d32_motionB.json 2>> gpurun_out/direct2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_direct2_single_stream.csv \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/direct2_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 3 -c 1 -f -o gpurun_out/ncu_insert_kernel \
  python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 1 --warmup 1 > gpurun_out/direct2_ncu_full.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/direct2_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v * 1000, 1) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/direct2.err

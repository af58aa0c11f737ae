#!/bin/bash
# On the GPU box: for the product build and every named variant prebuilt by profiles/build_variants.py,
# the clip-pipeline parity tests and two bench runs (alternating), summarised at the end.
#   gpurun --timeout 500 -- 'bash profiles/run_variants.sh static shift both'
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 30"
for v in "$@"; do
  timeout 120 python -m pytest tests/test_gpu_joint.py tests/test_gpu_zz_clip_table.py -x -q --slr-lib gpurun_variants/libslr_splat_$v.so \
      > gpurun_out/variant_${v}_pytest.log 2>&1; echo "$v pytest rc=$?" >> gpurun_out/variants_status.txt
done
for i in 1 2; do
  timeout 60 python bench.py $B > gpurun_out/variant_base_$i.json 2>> gpurun_out/variants.err
  for v in "$@"; do
    SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 60 python profiles/bench_with_lib.py $B > gpurun_out/variant_${v}_$i.json 2>> gpurun_out/variants.err
  done
done
cat gpurun_out/variants_status.txt
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/variant_*_[0-9].json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "gather single-stream frac", round(r["single_stream"]["frac"], 4),
              {k: round(v, 4) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY

mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --steps 30"
for i in 1 2; do
  timeout 90 python bench.py $B > gpurun_out/tune_base_$i.json 2>> gpurun_out/tune.err
  SLR_GATHER_SHAPE=1x4 timeout 90 python bench.py $B > gpurun_out/tune_shape1x4_$i.json 2>> gpurun_out/tune.err
  for v in loads10 loads12 loads16; do
    SLR_LIB=gpurun_variants/libslr_splat_$v.so timeout 90 python profiles/bench_with_lib.py $B > gpurun_out/tune_${v}_$i.json 2>> gpurun_out/tune.err
  done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/tune_*_[0-9].json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, round(d["value"], 1), "live", round(r["frac"], 4), "single", round(r["single_stream"]["frac"], 4),
              {k: round(v, 4) for k, v in r["all_kernels_ms_per_frame"].items() if k in ("slr_clip_gather", "slr_clip_expand", "slr_scene_prep")})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/tune.err

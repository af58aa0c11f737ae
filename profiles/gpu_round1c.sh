#!/bin/bash
# (historical record: run at commit "bench: live timing brackets the dominant kernel only ...")
# Second GPU call of this session: A/B of the rowgather CTA shape through bench.py itself (alternating runs),
# and the main-stream-priority stage of the knob sweep.
mkdir -p gpurun_out
S=gpurun_out/status_c.txt
: > $S
t0=$(date +%s)
stamp() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" >> $S; }
B="python bench.py --no-cpu-baseline --no-e2e --steps 30"
for i in 1 2; do
  timeout 90 $B > gpurun_out/ab_default_$i.json 2> gpurun_out/ab.err; stamp ab_default_$i $?
  SLR_GATHER_SHAPE=2x2 timeout 90 $B > gpurun_out/ab_2x2_$i.json 2>> gpurun_out/ab.err; stamp ab_2x2_$i $?
  SLR_GATHER_SHAPE=4x1 timeout 90 $B > gpurun_out/ab_4x1_$i.json 2>> gpurun_out/ab.err; stamp ab_4x1_$i $?
done
SLR_GATHER_SHAPE=2x2 SLR_BATCH=16 timeout 90 $B > gpurun_out/ab_2x2_b16.json 2>> gpurun_out/ab.err; stamp ab_2x2_b16 $?
SWEEP_STAGES=main timeout 120 python profiles/sweep_variants.py > gpurun_out/sweep_main.jsonl 2> gpurun_out/sweep_main.err; stamp sweep_main $?
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print(f, round(d["value"], 1), "live frac", round(r["frac"], 4), "single", round(r.get("single_stream", {}).get("frac", 0), 4),
              {k: round(v, 4) for k, v in r["all_kernels_ms_per_frame"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
cat $S

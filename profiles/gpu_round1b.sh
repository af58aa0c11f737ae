#!/bin/bash
# (historical record: run at commit "gather pad-smem knob; sweep stage for side-stream room")
# ONE gpurun call (round 1, second session): parity suite, bench, knob sweep, parity + bench under the
# winning knobs, ncu launch list + one full capture of the hot kernel.  Ordered by priority; every step has
# its own timeout and writes to gpurun_out/ as it goes.
#   gpurun --timeout 560 -- 'bash profiles/gpu_round1b.sh'
mkdir -p gpurun_out
S=gpurun_out/status.txt
: > $S
t0=$(date +%s)
stamp() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" >> $S; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1

timeout 300 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; stamp pytest_gpu $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; stamp smoke $?
timeout 150 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; stamp bench_default $?
timeout 150 python profiles/sweep_variants.py > gpurun_out/sweep_variants.jsonl 2> gpurun_out/sweep_variants.err; stamp sweep $?

# the winner's knobs as environment for the next steps
python - > gpurun_out/best_env.sh <<'PY'
import json
best = None
try:
    for line in open("gpurun_out/sweep_variants.jsonl"):
        row = json.loads(line)
        if "best" in row:
            best = row["best"]
except Exception:
    pass
if best:
    print("export SLR_GATHER_SHAPE=%s SLR_EXPAND_CLAIM=%s SLR_SMEM_CARVEOUT=%s SLR_SIDE_PRIORITY=%s SLR_BATCH=%s SLR_GATHER_PAD_SMEM=%s"
          % (best["shape"], best["claim"], best["carveout"], best["priority"], best["batch"], best["pad"]))
PY
cat gpurun_out/best_env.sh >> $S
source gpurun_out/best_env.sh
timeout 150 python -m pytest tests/test_gpu_joint.py tests/test_gpu_vs_reference_kernel.py -x -q > gpurun_out/pytest_best.log 2>&1; stamp pytest_best $?
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err; stamp bench_best $?

# ncu: one full capture of the hot kernel, then the launch list of one step (shares only)
timeout 150 ncu --set full --clock-control none --import-source on -k regex:rowgather -s 6 -c 1 -o gpurun_out/rowgather_r01b \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-pipeline > gpurun_out/ncu_full.log 2>&1; stamp ncu_full $?
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 45 --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-pipeline > gpurun_out/ncu_list.log 2>&1; stamp ncu_list $?
timeout 120 python profiles/sweep_configs.py > gpurun_out/sweep_configs_r01b.json 2> gpurun_out/sweep_configs.err; stamp sweep_configs $?
cat $S

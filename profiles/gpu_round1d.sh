#!/bin/bash
# Final GPU call of round 1 (second session): parity suite, smoke, bench (both arms), other motions,
# ncu captures of the final state, config sweep, compute-sanitizer memcheck on small cases.
mkdir -p gpurun_out
S=gpurun_out/status_d.txt
: > $S
t0=$(date +%s)
stamp() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" >> $S; }
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; stamp pytest_gpu $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; stamp smoke $?
timeout 150 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; stamp bench $?
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_ref.err; stamp bench_ref $?
timeout 60 python bench.py --motion B --no-e2e --no-cpu-baseline > gpurun_out/bench_final_motionB.json 2>> gpurun_out/bench_final.err; stamp motionB $?
timeout 60 python bench.py --motion C --no-e2e --no-cpu-baseline > gpurun_out/bench_final_motionC.json 2>> gpurun_out/bench_final.err; stamp motionC $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:rowgather -s 6 -c 1 -o gpurun_out/rowgather_final \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-pipeline > gpurun_out/ncu_full.log 2>&1; stamp ncu_full $?
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 45 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-pipeline > gpurun_out/ncu_list.log 2>&1; stamp ncu_list $?
timeout 100 python profiles/sweep_configs.py > gpurun_out/sweep_configs_final.json 2> gpurun_out/sweep_configs.err; stamp sweep_configs $?
for i in 1 2 3; do timeout 60 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_repeat_$i.json 2>> gpurun_out/bench_final.err; stamp repeat_$i $?; done
cat $S

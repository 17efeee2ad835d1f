#!/bin/bash
# Round 2, call O: fresh evidence on the committed state: bench line, launch list, full capture of one gap's pass launches
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_raw.csv \
    python bench.py --steps 1 --warmup 3 --profiles 8 --no-cpu-baseline --no-c5 --no-extra > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> $O/summary.txt
# second L2 call of trace_run: phase-1 launches 30..59; gap index 4 = launches 42,43,44 ; first seed = 30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1_ --launch-skip 42 --launch-count 3 \
    -o $O/pass_g4 -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_pass.log 2>&1
echo "ncu pass rc=$?" >> $O/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1_ --launch-skip 30 --launch-count 1 \
    -o $O/pass_seed0 -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_seed0.log 2>&1
echo "ncu seed0 rc=$?" >> $O/summary.txt
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 3 > $O/pass_timing.log 2>&1
grep "pass timing" $O/pass_timing.log | tail -1
cat $O/summary.txt
ls -la $O

#!/bin/bash
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 2 > $O/timing_c4.log 2>&1
echo "timing rc=$?" >> $O/summary.txt
grep "resident launch" $O/timing_c4.log | tail -24
timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 1 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:resident_kernel -c 7 -o $O/res_c4 -f python tools/trace_run.py --workload c4 --profiles 0 --reps 1 > $O/ncu.log 2>&1
echo "ncu rc=$?" >> $O/summary.txt
ls -la $O

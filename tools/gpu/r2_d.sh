#!/bin/bash
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 2 > $O/timing_c4.log 2>&1
echo "timing rc=$?" >> $O/summary.txt
IADH_RES_WIDE=400:128 IADH_PASS_MIN=100000000 IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 2 > $O/timing_c4_wide128.log 2>&1
grep -A1 "resident launch" $O/timing_c4.log | tail -48

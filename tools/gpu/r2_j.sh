#!/bin/bash
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
IADH_TRACE=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 4 > $O/trace.log 2>&1
grep "iadh\]\|L2 call wall" $O/trace.log | tail -24 | cut -c1-120
